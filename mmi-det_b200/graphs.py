"""CUDA-graph replay of launch-bound inference: at batch 1 a fusion block is a dozen short kernels (token gather, RMSNorm, four
projections, conv prologue, scan, token scatter ...) whose launch overhead exceeds their run time; capturing the block once
per input shape and replaying the graph removes it.  Inference only (no autograd through a replay); the module's parameters
are read at replay time, so loading new weights in place needs no recapture, while new shapes / dtypes capture a new graph.

    fus = MambaFusion(256).cuda().half().eval()
    fast = Graphed(fus)
    rgb_out, ir_out = fast([rgb, ir])          # first call per shape captures, later calls replay
"""
from __future__ import annotations

import torch


def _flatten(x):
    if isinstance(x, torch.Tensor):
        return [x]
    out = []
    for v in x:
        out += _flatten(v)
    return out


def _rebuild(x, it):
    if isinstance(x, torch.Tensor):
        return next(it)
    return type(x)(_rebuild(v, it) for v in x)


class Graphed(torch.nn.Module):
    """Wrap a module whose forward takes tensors / nested lists of tensors and returns the same kind of structure."""

    def __init__(self, module: torch.nn.Module, warmup: int = 2):
        super().__init__()
        self.module = module
        self.warmup = warmup
        self._graphs = {}

    def _key(self, flat):
        return tuple((tuple(t.shape), t.dtype, t.device) for t in flat)

    @torch.no_grad()
    def forward(self, x):
        flat = _flatten(x)
        if not all(t.is_cuda for t in flat):
            raise RuntimeError("Graphed: CUDA tensors required")
        key = self._key(flat)
        entry = self._graphs.get(key)
        if entry is None:
            static_in = [torch.empty_like(t) for t in flat]
            for s, t in zip(static_in, flat):
                s.copy_(t)
            sx = _rebuild(x, iter(static_in))
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # warm-up outside the capture: lazy library loads, tensor-map cache, allocator
                for _ in range(self.warmup):
                    self.module(sx)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self.module(sx)
            entry = (graph, static_in, static_out)
            self._graphs[key] = entry
        graph, static_in, static_out = entry
        for s, t in zip(static_in, flat):
            s.copy_(t)
        graph.replay()
        return _rebuild(static_out, iter([t.clone() for t in _flatten(static_out)]))
