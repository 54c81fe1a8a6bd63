"""The synthetic-stream generator (include/mvgen_core.h) through libmvgen.so — the host generator alone, for
processes that must not map the product library (bench.py --impl reference times the reference's CPU code; the
only thing it needs from this repo is the input stream). Same bytes as motionscan.synth_host."""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace

import numpy as np

from . import MV_DTYPE, PKG_DIR, MvgenSpec

LIB_PATH = PKG_DIR / "libmvgen.so"
_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FileNotFoundError(f"{LIB_PATH} is missing — build it with `python {PKG_DIR.name}/build.py`")
        L = C.CDLL(str(LIB_PATH))
        L.mscan_synth_preset.argtypes = [C.POINTER(MvgenSpec), C.c_int, C.c_uint64]
        L.mscan_synth_host_counts.argtypes = [C.POINTER(MvgenSpec), C.c_uint64, C.c_uint32, C.c_void_p, C.c_int]
        L.mscan_synth_host_fill.argtypes = [C.POINTER(MvgenSpec), C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib = L
    return _lib


def synth_preset(config: int, seed: int) -> MvgenSpec:
    s = MvgenSpec()
    lib().mscan_synth_preset(C.byref(s), config, seed)
    return s


def synth_host(spec: MvgenSpec, frame0: int, n_frames: int, n_threads: int = 8):
    L = lib()
    cnt = np.zeros(n_frames, dtype=np.uint32)
    L.mscan_synth_host_counts(C.byref(spec), frame0, n_frames, cnt.ctypes.data, n_threads)
    off = np.zeros(n_frames + 1, dtype=np.uint64)
    np.cumsum(cnt, out=off[1:])
    recs = np.zeros(int(off[-1]), dtype=MV_DTYPE)
    pts = np.zeros(n_frames, dtype=np.float64)
    L.mscan_synth_host_fill(C.byref(spec), frame0, n_frames, off.ctypes.data, recs.ctypes.data, pts.ctypes.data, n_threads)
    return cnt, off, recs, pts


def shipped_env_params():
    """config/motion_trim.env of the reference as plain attributes."""
    return SimpleNamespace(mv_threshold_sq=4.0, block_size=16, block_shift=4, vectors_needed=4, clusters_needed=2,
                           vertical_mask=0.05, adjacency=4, max_gap_sec=5.0, padding_sec=0.5, min_savings_pct=5.0)


def code_default_params():
    """include/motion_trim/config.hpp:57-123 code defaults."""
    return SimpleNamespace(mv_threshold_sq=16.0, block_size=16, block_shift=4, vectors_needed=2, clusters_needed=2,
                           vertical_mask=0.05, adjacency=4, max_gap_sec=5.0, padding_sec=0.5, min_savings_pct=5.0)
