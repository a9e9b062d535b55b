"""ctypes loader and in-tree builder for libmmidet_b200.so (the C ABI of include/mmidet_b200.h).

There is deliberately no CPU / eager fallback: if the library is missing, or no sm_100 device is present,
every operator raises RuntimeError."""
from __future__ import annotations

import ctypes
import glob
import os
import subprocess

_DIR = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_DIR)
SO_PATH = os.path.join(_DIR, "libmmidet_b200.so")
CSRC = os.path.join(_DIR, "csrc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fvisibility=default"]

MMI_F32, MMI_BF16, MMI_F16 = 0, 1, 2
FLAG_NO_GEOM = 1
FLAG_DELTA_SOFTPLUS = 2
FLAG_CFG_SHIFT = 4
FLAG_NSEG_SHIFT = 8

_lib = None


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


HASH_PATH = SO_PATH + ".srchash"


def _dep_files():
    return sources() + sorted(glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                              glob.glob(os.path.join(_ROOT, "include", "*.h")))


def source_hash() -> str:
    """sha256 over every source / header the library is built from (names + contents) and the compiler flags."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in _dep_files():
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def needs_build() -> bool:
    """Stale = the sources' content hash differs from the one recorded next to the library at build time.  (File times are
    useless here: the repository travels to the GPU box as a snapshot whose mtimes are in copy order.)"""
    if not os.path.exists(SO_PATH) or not os.path.exists(HASH_PATH):
        return True
    try:
        return open(HASH_PATH).read().strip() != source_hash()
    except OSError:
        return True


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> mmi-det_b200/libmmidet_b200.so (in-tree).
    Every csrc/*.cu is compiled to an object in parallel (the scan kernels dominate), then linked.  Serialised across
    processes by a file lock and published atomically, so that the ranks of a torchrun job can all call load() at once."""
    import fcntl
    lock_path = SO_PATH + ".lock"
    with open(lock_path, "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                return SO_PATH  # another process built it while this one waited for the lock
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> str:
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = os.path.join(_DIR, "build")
    os.makedirs(objdir, exist_ok=True)
    hdrs = [f for f in _dep_files() if not f.endswith(".cu")]
    import hashlib
    hdr_hash = hashlib.sha256(b"".join(open(f, "rb").read() for f in hdrs) + " ".join(NVCC_FLAGS).encode()).hexdigest()
    log = []

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        tag = obj + ".srchash"
        want = hashlib.sha256(open(src, "rb").read() + hdr_hash.encode()).hexdigest()
        if not force and os.path.exists(obj) and os.path.exists(tag) and open(tag).read().strip() == want:
            return obj
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + r.stdout + r.stderr)
        with open(tag, "w") as f:
            f.write(want)
        log.append(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, sources()))
    tmp = SO_PATH + f".tmp{os.getpid()}"
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs,
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + r.stdout + r.stderr)
    os.replace(tmp, SO_PATH)  # atomic: a concurrent dlopen sees the old or the new file, never a partial one
    with open(HASH_PATH, "w") as f:
        f.write(source_hash())
    if verbose:
        print("".join(log))
    return SO_PATH


_c = ctypes
_vp, _i, _i64, _fp = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_void_p

_SIGNATURES = {
    "mmi_last_error": (_c.c_char_p, []),
    "mmi_version": (_i, []),
    "mmi_device_info": (_i, [_c.POINTER(_i)] * 3),
    "mmi_selscan_chunk": (_i, []),
    "mmi_selscan_fwd_ws_bytes": (_i64, [_i] * 4),
    "mmi_selscan_fwd": (_i, [_vp] * 12 + [_i] * 4 + [_i64] * 4 + [_i] * 3 + [_vp]),
    "mmi_selscan_bwd_ws_bytes": (_i64, [_i] * 4),
    "mmi_selscan_bwd": (_i, [_vp] * 17 + [_i] * 4 + [_i64] * 4 + [_i] * 3 + [_vp]),
    "mmi_pscan_ws_bytes": (_i64, [_i] * 4),
    "mmi_pscan_fwd": (_i, [_vp] * 4 + [_i] * 4 + [_vp]),
    "mmi_pscan_bwd": (_i, [_vp] * 6 + [_i] * 4 + [_vp]),
    "mmi_pscan_fwd_f64": (_i, [_vp] * 4 + [_i] * 4 + [_vp]),
    "mmi_pscan_bwd_f64": (_i, [_vp] * 6 + [_i] * 4 + [_vp]),
    "mmi_ffm_kept_range": (None, [_i, _i] + [_c.POINTER(_i)] * 4),
    "mmi_ffm_extract": (_i, [_vp] * 4 + [_i] * 4 + [_vp]),
    "mmi_separation_loss": (_i, [_vp] * 2 + [_i] * 2 + [_vp]),
    "mmi_ffm_pattern_ws_bytes": (_i64, [_i, _i, _i]),
    "mmi_avgpool_fwd": (_i, [_vp] * 2 + [_i] * 6 + [_vp]),
    "mmi_avgpool_bwd": (_i, [_vp] * 2 + [_i] * 6 + [_vp]),
    "mmi_upsample_bilinear_fwd": (_i, [_vp] * 2 + [_i] * 6 + [_vp]),
    "mmi_upsample_bilinear_bwd": (_i, [_vp] * 2 + [_i] * 6 + [_vp]),
    "mmi_ffm_pattern_fwd": (_i, [_vp] * 8 + [_i] * 5 + [_vp]),
    "mmi_ffm_pattern_bwd": (_i, [_vp] * 11 + [_i] * 4 + [_vp]),
    "mmi_causal_conv1d_fwd": (_i, [_vp] * 4 + [_i] * 4 + [_i64] * 2 + [_i] * 2 + [_vp]),
    "mmi_causal_conv1d_bwd": (_i, [_vp] * 7 + [_i] * 4 + [_i64] * 3 + [_i] * 2 + [_vp]),
    "mmi_rmsnorm_fwd": (_i, [_vp] * 3 + [_i64, _i, _i64, _i64, _c.c_float, _i, _i, _vp]),
    "mmi_rmsnorm_bwd": (_i, [_vp] * 5 + [_i64, _i, _i64, _i64, _i64, _c.c_float, _i, _i, _vp]),
    "mmi_tokens_gather": (_i, [_vp] * 3 + [_i] * 4 + [_vp]),
    "mmi_tokens_scatter": (_i, [_vp] * 3 + [_i] * 4 + [_vp]),
    "mmi_u8_split_normalize": (_i, [_vp] * 3 + [_i] * 4 + [_vp]),
    "mmi_detect_decode": (_i, [_vp] * 3 + [_i] * 5 + [_c.c_float, _vp, _i64, _i64, _i, _vp]),
    "mmi_nms_candidates": (_i, [_vp] * 4 + [_i, _i64, _i, _c.c_float, _i, _vp]),
    "mmi_nms_suppress": (_i, [_vp] * 7 + [_i] * 4 + [_c.c_float, _c.c_float, _vp]),
    "mmi_selscan_fwd_bwd_host": (_i, [_vp] * 16 + [_i] * 6),
    "mmi_host_workspace_free": (None, []),
}


def exported_symbols():
    return sorted(_SIGNATURES)


def load() -> ctypes.CDLL:
    """Load the extension (building it first if sources are newer and nvcc is present). Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    # MMIDET_SO=<path>: load exactly this build of the library (kernel tuning / ablation builds, profiles/r02_scan_generations.txt).
    # A missing or incomplete file raises -- there is no fallback behind it either.
    override = os.environ.get("MMIDET_SO")
    if override:
        lib = ctypes.CDLL(override)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib
    if needs_build():
        # a source newer than the library must compile: never run yesterday's binary against today's sources.  The one
        # exception is a box without nvcc (the library was built elsewhere and shipped): then the shipped file is loaded.
        try:
            build()
        except FileNotFoundError as e:
            if not os.path.exists(SO_PATH):
                raise RuntimeError(f"libmmidet_b200.so is not built and nvcc is not available here: {e}") from e
        except RuntimeError as e:
            raise RuntimeError(f"libmmidet_b200.so is older than its sources and the rebuild failed: {e}") from e
    lib = ctypes.CDLL(SO_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError => header/library mismatch, fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int, what: str = "libmmidet_b200") -> None:
    if code != 0:
        msg = load().mmi_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {code}): {msg}")
