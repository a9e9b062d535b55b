"""CPU: the C-ABI library builds, loads and exports every symbol include/mmidet_b200.h declares; argument errors
are reported through return codes + mmi_last_error (no compute calls here -- there is no GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from mmidet_b200 import _lib
    if _lib.needs_build():
        _lib.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "mmidet_b200.h")).read()
    declared = set(re.findall(r"\b(mmi_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    from mmidet_b200 import _lib
    assert declared == set(_lib.exported_symbols())
    for s in declared:
        assert hasattr(lib, s), s


def test_version_and_chunk(lib):
    assert lib.mmi_version() >= 100
    assert lib.mmi_selscan_chunk() in (8, 16, 32, 64)
    assert lib.mmi_selscan_bwd_ws_bytes(2, 100, 64, 16) > 0
    assert lib.mmi_pscan_ws_bytes(2, 100, 64, 16) > 0


def test_argument_errors_are_reported(lib):
    null = ctypes.c_void_p(None)
    rc = lib.mmi_selscan_fwd(null, null, null, null, null, null, null, null, null, null, null, null, 1, 8, 16, 16, 16, 16,
                             16, 16, 16, 0, 0, null)
    assert rc != 0 and b"null" in lib.mmi_last_error()
    one = ctypes.c_void_p(256)  # never dereferenced: validation fails first
    rc = lib.mmi_selscan_fwd(one, one, null, one, one, one, one, null, one, null, null, null, 1, 8, 12, 16, 12, 12, 0, 12, 16,
                             0, 0, null)
    assert rc != 0 and b"multiple of 8" in lib.mmi_last_error()
    rc = lib.mmi_selscan_fwd(one, one, null, one, one, one, one, null, one, null, null, null, 1, 8, 16, 8, 16, 16, 0, 16, 16,
                             0, 0, null)
    assert rc != 0 and b"d_state" in lib.mmi_last_error()


def test_ffm_kept_range_matches_oracle(lib):
    """host helper vs the oracle's restatement of the reference's slice semantics (common.py:44-56)."""
    import numpy as np
    from oracle import oracle as O
    for H, W in [(8, 8), (16, 16), (20, 20), (7, 7), (8, 12), (5, 64), (64, 6), (3, 3), (1, 1), (80, 80), (160, 160)]:
        v = [ctypes.c_int() for _ in range(4)]
        lib.mmi_ffm_kept_range(H, W, *[ctypes.byref(i) for i in v])
        r0, r1, c0, c1 = [i.value for i in v]
        m = np.zeros((H, W), bool)
        m[r0:r1, c0:c1] = True
        keep_high, keep_low = O.ffm_masks(H, W)
        assert np.array_equal(m, keep_low), (H, W)
        assert np.array_equal(~m, keep_high), (H, W)


def test_no_cpu_fallback():
    """CPU tensors must raise, never silently compute."""
    import torch
    from mmidet_b200 import ops
    from mmidet_b200.ffm import extract_frequency2
    from mmidet_b200.pscan import pscan
    with pytest.raises(RuntimeError):
        ops.selective_scan(torch.randn(1, 8, 16), torch.randn(1, 8, 16), torch.randn(16, 16), torch.randn(1, 8, 16),
                           torch.randn(1, 8, 16), torch.randn(16))
    with pytest.raises(RuntimeError):
        pscan(torch.rand(1, 4, 2, 16), torch.rand(1, 4, 2, 16))
    with pytest.raises(RuntimeError):
        extract_frequency2(torch.randn(1, 2, 8, 8))
