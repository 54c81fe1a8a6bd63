"""ctypes binding of oracle/libmscan_oracle.so — the CPU oracle (test infrastructure only)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "oracle" / "libmscan_oracle.so"


class OrcCfg(C.Structure):
    _fields_ = [
        ("mv_threshold_sq", C.c_double),
        ("block_shift", C.c_int32),
        ("clusters_needed", C.c_int32),
        ("vertical_margin", C.c_int32),
        ("grid_w", C.c_int32),
        ("grid_h", C.c_int32),
        ("vectors_needed", C.c_uint8),
    ]


class OrcResult(C.Structure):
    _fields_ = [
        ("decision", C.c_int32),
        ("n_motion_frames", C.c_uint32),
        ("n_segments", C.c_uint32),
        ("reserved", C.c_uint32),
        ("out_dur", C.c_double),
        ("time_removed", C.c_double),
        ("saved_pct", C.c_double),
    ]


SEG_DTYPE = np.dtype([("start", "<f8"), ("end", "<f8")])
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(str(LIB))
        vp, u32, i = C.c_void_p, C.c_uint32, C.c_int
        L.orc_geometry.argtypes = [i, i, i, i, C.c_float, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.orc_geometry.restype = None
        L.orc_check_frame.argtypes = [C.POINTER(OrcCfg), vp, C.c_int64, vp]
        L.orc_check_frame.restype = i
        L.orc_full_count.argtypes = [C.POINTER(OrcCfg), vp, C.c_int64, vp]
        L.orc_full_count.restype = u32
        L.orc_scan_frames.argtypes = [C.POINTER(OrcCfg), vp, vp, u32, vp, vp, i]
        L.orc_scan_frames.restype = None
        L.orc_scan_frames_mt.argtypes = [C.POINTER(OrcCfg), vp, vp, u32, vp, vp, i, i]
        L.orc_scan_frames_mt.restype = None
        L.orc_select_range.argtypes = [vp, vp, u32, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, vp]
        L.orc_select_range.restype = u32
        L.orc_select_pipeline.argtypes = [vp, vp, u32, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, vp]
        L.orc_select_pipeline.restype = u32
        L.orc_merge_timestamps.argtypes = [vp, u32]
        L.orc_merge_timestamps.restype = u32
        L.orc_build_segments.argtypes = [vp, u32, C.c_double, C.c_double, vp]
        L.orc_build_segments.restype = u32
        L.orc_savings.argtypes = [vp, u32, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_savings.restype = None
        L.orc_video_tail.argtypes = [vp, vp, u32, C.c_double, C.c_double, C.c_double, C.c_double, vp, C.POINTER(OrcResult)]
        L.orc_video_tail.restype = None
        _lib = L
    return _lib


def geometry(width, height, block_size=16, block_shift=4, vertical_mask=0.05):
    gw, gh, m = C.c_int32(), C.c_int32(), C.c_int32()
    lib().orc_geometry(width, height, block_size, block_shift, vertical_mask, C.byref(gw), C.byref(gh), C.byref(m))
    return gw.value, gh.value, m.value


def make_cfg(params, gw, gh, margin) -> OrcCfg:
    """params: motionscan.Params (or anything with the same attribute names)."""
    c = OrcCfg()
    c.mv_threshold_sq = params.mv_threshold_sq
    c.block_shift = params.block_shift
    c.clusters_needed = params.clusters_needed
    c.vertical_margin = margin
    c.grid_w = gw
    c.grid_h = gh
    c.vectors_needed = params.vectors_needed & 0xFF  # config.hpp:75 static_cast<uint8_t>
    return c


def check_frame(cfg: OrcCfg, recs: np.ndarray | None) -> int:
    grid = np.zeros(max(cfg.grid_w * cfg.grid_h, 1), dtype=np.uint8)
    if recs is None:
        return lib().orc_check_frame(C.byref(cfg), None, 0, grid.ctypes.data)
    recs = np.ascontiguousarray(recs)
    assert recs.dtype.itemsize == 40, "records must keep the 40-byte AVMotionVector layout"
    return lib().orc_check_frame(C.byref(cfg), recs.ctypes.data, recs.nbytes, grid.ctypes.data)


def full_count(cfg: OrcCfg, recs: np.ndarray | None) -> int:
    grid = np.zeros(max(cfg.grid_w * cfg.grid_h, 1), dtype=np.uint8)
    if recs is None:
        return lib().orc_full_count(C.byref(cfg), None, 0, grid.ctypes.data)
    recs = np.ascontiguousarray(recs)
    assert recs.dtype.itemsize == 40, "records must keep the 40-byte AVMotionVector layout"
    return lib().orc_full_count(C.byref(cfg), recs.ctypes.data, recs.nbytes, grid.ctypes.data)


def scan_frames(cfg: OrcCfg, recs: np.ndarray, rec_off: np.ndarray, early_exit=False, threads=1):
    n = len(rec_off) - 1
    flags = np.zeros(n, dtype=np.uint8)
    counts = np.zeros(n, dtype=np.uint32)
    rec_off = np.ascontiguousarray(rec_off, dtype=np.uint64)
    recs = np.ascontiguousarray(recs)
    assert recs.dtype.itemsize == 40, "records must keep the 40-byte AVMotionVector layout"
    rp = recs.ctypes.data if recs.size else None
    if threads > 1:
        lib().orc_scan_frames_mt(C.byref(cfg), rp, rec_off.ctypes.data, n, flags.ctypes.data, counts.ctypes.data, int(early_exit), threads)
    else:
        lib().orc_scan_frames(C.byref(cfg), rp, rec_off.ctypes.data, n, flags.ctypes.data, counts.ctypes.data, int(early_exit))
    return flags, counts


def video_tail(pts, flags, duration, max_gap, padding, min_savings_pct):
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    flags = np.ascontiguousarray(flags, dtype=np.uint8)
    n = len(pts)
    segs = np.zeros(max(n, 1), dtype=SEG_DTYPE)
    res = OrcResult()
    lib().orc_video_tail(pts.ctypes.data, flags.ctypes.data, n, duration, max_gap, padding, min_savings_pct, segs.ctypes.data, C.byref(res))
    return segs[: res.n_segments].copy(), res


def select_pipeline(pts_ticks, is_key, time_base, video_fps, target_fps, duration, chunk_sec):
    """Frame indices the reference's pipeline hands to check_frame (chunk order)."""
    pts_ticks = np.ascontiguousarray(pts_ticks, dtype=np.int64)
    is_key = np.ascontiguousarray(is_key, dtype=np.uint8)
    n = len(pts_ticks)
    n_chunks = int(np.ceil(duration / chunk_sec)) + 2
    out = np.zeros(max(n, 1) * 1 + 8, dtype=np.uint32)
    k = lib().orc_select_pipeline(pts_ticks.ctypes.data, is_key.ctypes.data, n, time_base, video_fps, target_fps, duration,
                                  chunk_sec, out.ctypes.data)
    assert k <= n
    return out[:k].copy()
