"""ctypes binding of libmotionscan.so — the C ABI in include/motionscan.h.

This module is plumbing for tests/ and bench.py; the product is the shared library and the C++ host
above it. It never computes anything itself and has no CPU fallback: if the library is missing the
import of :func:`lib` raises, and without a GPU ``Context()`` raises ``MscanError(MSCAN_ERR_CUDA)``.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent.parent
import os

# MSCAN_LIB: load another build of the same library (kernel tuning experiments, tools/ka_sweep.py)
LIB_PATH = Path(os.environ["MSCAN_LIB"]) if os.environ.get("MSCAN_LIB") else PKG_DIR / "libmotionscan.so"

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_CAPACITY, ERR_UNSUPPORTED = range(6)
NO_MOTION, CUT, FULL_COPY = 0, 1, 2

# FFmpeg's AVMotionVector, 40 bytes (reference src/motion_scanner.cpp:224-226)
MV_DTYPE = np.dtype(
    {
        "names": ["source", "w", "h", "src_x", "src_y", "dst_x", "dst_y", "flags", "motion_x", "motion_y", "motion_scale"],
        "formats": ["<i4", "u1", "u1", "<i2", "<i2", "<i2", "<i2", "<u8", "<i4", "<i4", "<u2"],
        "offsets": [0, 4, 5, 6, 8, 10, 12, 16, 24, 28, 32],
        "itemsize": 40,
    }
)
# mscan_mv8: bytes 6..13 of the native record
MV8_DTYPE = np.dtype([("src_x", "<i2"), ("src_y", "<i2"), ("dst_x", "<i2"), ("dst_y", "<i2")])
SEG_DTYPE = np.dtype([("start", "<f8"), ("end", "<f8")])
STAGING_AUTO, STAGING_PACK, STAGING_NATIVE, STAGING_ELIDE, STAGING_COMPACT = 0, 1, 2, 3, 4


class Params(C.Structure):
    _fields_ = [
        ("mv_threshold_sq", C.c_double),
        ("block_size", C.c_int32),
        ("block_shift", C.c_int32),
        ("vectors_needed", C.c_int32),
        ("clusters_needed", C.c_int32),
        ("vertical_mask", C.c_float),
        ("adjacency", C.c_int32),
        ("max_gap_sec", C.c_double),
        ("padding_sec", C.c_double),
        ("min_savings_pct", C.c_double),
    ]


class Geometry(C.Structure):
    _fields_ = [("grid_w", C.c_int32), ("grid_h", C.c_int32), ("vertical_margin", C.c_int32), ("reserved", C.c_int32)]


class VideoResult(C.Structure):
    _fields_ = [
        ("decision", C.c_int32),
        ("n_motion_frames", C.c_uint32),
        ("n_segments", C.c_uint32),
        ("reserved", C.c_uint32),
        ("out_dur", C.c_double),
        ("time_removed", C.c_double),
        ("saved_pct", C.c_double),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("scan_launches", C.c_uint64),
        ("segment_launches", C.c_uint64),
        ("aux_launches", C.c_uint64),
        ("frames_scanned", C.c_uint64),
        ("records_scanned", C.c_uint64),
        ("h2d_bytes", C.c_uint64),
        ("d2h_bytes", C.c_uint64),
        ("scan_ms", C.c_double),
        ("segment_ms", C.c_double),
        ("records_projected", C.c_uint64),
        ("project_ms", C.c_double),
        ("peer_bytes", C.c_uint64),
        ("records_elided", C.c_uint64),
        ("elided_bytes", C.c_uint64),
    ]


class MvgenSpec(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64),
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("gop", C.c_int32),
        ("window", C.c_int32),
        ("frames_per_video", C.c_int32),
        ("p_window_active", C.c_uint32),
        ("max_blobs", C.c_int32),
        ("p_split2", C.c_uint32),
        ("p_noise", C.c_uint32),
        ("p_single", C.c_uint32),
        ("p_oob", C.c_uint32),
        ("dense", C.c_int32),
        ("p_dense_move", C.c_uint32),
        ("static_a0", C.c_int32),
        ("static_a1", C.c_int32),
        ("static_b0", C.c_int32),
        ("static_b1", C.c_int32),
        ("fps", C.c_double),
        ("scatter", C.c_int32),
        ("p_rec_move", C.c_uint32),
    ]


RESULT_DTYPE = np.dtype(
    [
        ("decision", "<i4"),
        ("n_motion_frames", "<u4"),
        ("n_segments", "<u4"),
        ("reserved", "<u4"),
        ("out_dur", "<f8"),
        ("time_removed", "<f8"),
        ("saved_pct", "<f8"),
    ]
)

# every symbol include/motionscan.h declares: (restype, argtypes)
_vp, _u32, _u64, _i, _d = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_double
_P = C.POINTER
SYMBOLS = {
    "mscan_abi_version": (_i, []),
    "mscan_device_count": (_i, [_P(_i)]),
    "mscan_status_string": (C.c_char_p, [_i]),
    "mscan_params_default": (_i, [_P(Params)]),
    "mscan_params_from_env": (_i, [_P(Params)]),
    "mscan_geometry_from_dims": (_i, [_P(Params), _i, _i, _P(Geometry)]),
    "mscan_create": (_i, [_i, _P(Params), _u64, _u64, _P(_vp)]),
    "mscan_destroy": (_i, [_vp]),
    "mscan_last_error": (C.c_char_p, [_vp]),
    "mscan_get_params": (_i, [_vp, _P(Params)]),
    "mscan_sync": (_i, [_vp]),
    "mscan_get_stats": (_i, [_vp, _P(Stats)]),
    "mscan_reset_stats": (_i, [_vp]),
    "mscan_set_profiling": (_i, [_vp, _i]),
    "mscan_video_open": (_i, [_vp, _u32, _i, _i]),
    "mscan_video_open_geometry": (_i, [_vp, _u32, _P(Geometry)]),
    "mscan_submit": (_i, [_vp, _u32, _u32, _vp, _vp, _vp, _P(_u64)]),
    "mscan_submit_packed": (_i, [_vp, _u32, _u32, _vp, _vp, _vp, _P(_u64)]),
    "mscan_submit_device": (_i, [_vp, _u32, _u32, _vp, _vp, _vp, _i, _vp, _P(_u64)]),
    "mscan_device_pci_bus_id": (_i, [_i, C.c_char_p, _i]),
    "mscan_pack_records": (_i, [_vp, _u64, _vp]),
    "mscan_compact_records": (_i, [_vp, _u64, _vp, _P(_u64)]),
    "mscan_elide_records": (_i, [_vp, _u32, _vp, C.c_size_t, _vp, _u32, _P(C.c_size_t)]),
    "mscan_elide_bound": (C.c_size_t, [_u32]),
    "mscan_submit_elided": (_i, [_vp, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _P(_u64)]),
    "mscan_set_staging_mode": (_i, [_vp, _i]),
    "mscan_set_pack_threads": (_i, [_vp, _i]),
    "mscan_reserve_staging": (_i, [_vp]),
    "mscan_collect_range": (_i, [_vp, _u32, _u64, _u32, _vp, _vp]),
    "mscan_flush": (_i, [_vp]),
    "mscan_collect": (_i, [_vp, _u32, _vp, _vp, _u32, _P(_u32)]),
    "mscan_segments": (_i, [_vp, _u32, _d, _vp, _u32, _P(_u32), _P(VideoResult)]),
    "mscan_motion_segments": (_i, [_vp, _u32, _d, _vp, _u32, _P(_u32), _P(VideoResult)]),
    "mscan_segments_batch": (_i, [_vp, _u32, _vp, _vp, _vp, _u64, _vp, _vp]),
    "mscan_video_close": (_i, [_vp, _u32]),
    "mscan_video_append_from": (_i, [_vp, _u32, _vp, _u32]),
    "mscan_host_alloc": (_i, [_vp, C.c_size_t, _P(_vp)]),
    "mscan_host_free": (_i, [_vp, _vp]),
    "mscan_host_register": (_i, [_vp, _vp, C.c_size_t, _i]),
    "mscan_host_unregister": (_i, [_vp, _vp]),
    "mscan_host_fence": (_i, [_vp]),
    "mscan_dev_alloc": (_i, [_vp, C.c_size_t, _P(_vp)]),
    "mscan_dev_free": (_i, [_vp, _vp]),
    "mscan_memcpy_h2d": (_i, [_vp, _vp, _vp, C.c_size_t]),
    "mscan_memcpy_d2h": (_i, [_vp, _vp, _vp, C.c_size_t]),
    "mscan_offsets_from_counts": (_i, [_vp, _vp, _u32, _vp, _vp]),
    "mscan_scan_device": (_i, [_vp, _vp, _vp, _vp, _P(Geometry), _u32, _u32, _vp, _vp, _vp]),
    "mscan_scan_device_packed": (_i, [_vp, _vp, _vp, _vp, _P(Geometry), _u32, _u32, _vp, _vp, _vp]),
    "mscan_pack_records_device": (_i, [_vp, _vp, _u64, _vp, _vp]),
    "mscan_segments_device": (_i, [_vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mscan_synth_preset": (_i, [_P(MvgenSpec), _i, _u64]),
    "mscan_synth_host_counts": (_i, [_P(MvgenSpec), _u64, _u32, _vp, _i]),
    "mscan_synth_host_fill": (_i, [_P(MvgenSpec), _u64, _u32, _vp, _vp, _vp, _i]),
    "mscan_synth_counts": (_i, [_vp, _P(MvgenSpec), _u64, _u32, _vp, _vp]),
    "mscan_synth_fill": (_i, [_vp, _P(MvgenSpec), _u64, _u32, _vp, _vp, _vp, _vp]),
}

_lib = None


class MscanError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libmotionscan status {code}: {msg}")
        self.code = code


def lib() -> C.CDLL:
    """Loads libmotionscan.so; fails loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FileNotFoundError(
                f"{LIB_PATH} is missing — build it with `python {PKG_DIR.name}/build.py` (there is no CPU fallback)"
            )
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "array must be contiguous"
        return a.ctypes.data
    return a


def default_params() -> Params:
    p = Params()
    lib().mscan_params_default(C.byref(p))
    return p


def shipped_env_params() -> Params:
    """config/motion_trim.env of the reference: the benchmark configuration."""
    p = default_params()
    p.mv_threshold_sq = 4.0
    p.vectors_needed = 4
    p.clusters_needed = 2
    p.vertical_mask = 0.05
    p.max_gap_sec = 5.0
    p.padding_sec = 0.5
    p.min_savings_pct = 5.0
    return p


def geometry_from_dims(p: Params, width: int, height: int) -> Geometry:
    g = Geometry()
    rc = lib().mscan_geometry_from_dims(C.byref(p), width, height, C.byref(g))
    if rc:
        raise MscanError(rc, "mscan_geometry_from_dims")
    return g


def pack_records(recs: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
    """mscan_pack_records: the 8-byte projection (MV8_DTYPE) of native records (MV_DTYPE). No GPU needed."""
    assert recs.dtype == MV_DTYPE and recs.flags["C_CONTIGUOUS"]
    if out is None:
        out = np.empty(len(recs), dtype=MV8_DTYPE)
    assert out.dtype == MV8_DTYPE and len(out) >= len(recs)
    rc = lib().mscan_pack_records(_ptr(recs), len(recs), _ptr(out))
    if rc:
        raise MscanError(rc, "mscan_pack_records")
    return out


def compact_records(recs: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
    """mscan_compact_records: the 8-byte projections of the records with src != dst, in order (the wire form of
    STAGING_COMPACT). Returns the filled prefix of `out` (e.g. a pinned buffer with room for len(recs) records)."""
    assert recs.dtype == MV_DTYPE and recs.flags["C_CONTIGUOUS"]
    if out is None:
        out = np.empty(max(len(recs), 1), dtype=MV8_DTYPE)
    assert out.dtype == MV8_DTYPE and len(out) >= len(recs)
    n = _u64(0)
    rc = lib().mscan_compact_records(_ptr(recs) if len(recs) else None, len(recs), _ptr(out), C.byref(n))
    if rc:
        raise MscanError(rc, "mscan_compact_records")
    return out[: n.value]


def compact_frames(recs: np.ndarray, off: np.ndarray, out: np.ndarray | None = None):
    """compact_records on every frame of a stream: (mv8 records of the moving records back to back — the filled prefix
    of `out` if given —, per-frame moving counts as uint32)."""
    n_frames = len(off) - 1
    if out is None:
        out = np.empty(max(int(off[-1] - off[0]), 1), dtype=MV8_DTYPE)
    cnt = np.zeros(n_frames, np.uint32)
    at = 0
    L = lib()
    n = _u64(0)
    for f in range(n_frames):
        a, b = int(off[f]), int(off[f + 1])
        if b > a:
            rc = L.mscan_compact_records(recs.ctypes.data + 40 * a, b - a, out.ctypes.data + 8 * at, C.byref(n))
            if rc:
                raise MscanError(rc, "mscan_compact_records")
            cnt[f] = n.value
            at += n.value
    return out[:at], cnt


def elide_records(recs: np.ndarray):
    """mscan_elide_records on one frame: (encoded bytes as uint8 array, tile_end16 uint32 array)."""
    assert recs.dtype == MV_DTYPE and recs.flags["C_CONTIGUOUS"]
    n = len(recs)
    cap = lib().mscan_elide_bound(n)
    raw = np.zeros(cap + 16, dtype=np.uint8)
    shift = (-raw.ctypes.data) % 16
    out = raw[shift : shift + cap]
    tiles = (n + 1023) // 1024
    te = np.zeros(max(tiles, 1), dtype=np.uint32)
    nbytes = C.c_size_t()
    rc = lib().mscan_elide_records(_ptr(recs) if n else None, n, out.ctypes.data, cap, _ptr(te), tiles, C.byref(nbytes))
    if rc:
        raise MscanError(rc, "mscan_elide_records")
    return out[: nbytes.value].copy(), te[:tiles].copy()


def elide_frames(recs: np.ndarray, rec_off: np.ndarray, out: np.ndarray | None = None):
    """mscan_elide_records on every frame of a stream: (enc uint8 array — `out` if given, e.g. a pinned buffer —,
    enc_off u64[F+1], tile_end16 u32[T])."""
    n_frames = len(rec_off) - 1
    parts, ends, enc_off = [], [], np.zeros(n_frames + 1, dtype=np.uint64)
    at = 0
    for f in range(n_frames):
        a, b = int(rec_off[f]), int(rec_off[f + 1])
        if b > a:
            e, te = elide_records(np.ascontiguousarray(recs[a:b]))
            parts.append(e)
            ends.append(te)
            at += len(e)
        enc_off[f + 1] = at
    if out is None:
        raw = np.zeros(at + 16, dtype=np.uint8)
        shift = (-raw.ctypes.data) % 16
        out = raw[shift : shift + at]
    assert out.dtype == np.uint8 and len(out) >= at and out.ctypes.data % 16 == 0
    pos = 0
    for e in parts:
        out[pos : pos + len(e)] = e
        pos += len(e)
    return out[:at], enc_off, (np.concatenate(ends) if ends else np.zeros(0, np.uint32)).astype(np.uint32)


def unelide_records(enc: np.ndarray, tile_end16: np.ndarray, n: int) -> np.ndarray:
    """Reference decoder of the static-elided form (pure numpy): back to MV8_DTYPE records."""
    out = np.zeros(n, dtype=MV8_DTYPE)
    flat = out.view(np.uint32).reshape(n, 2) if n else np.zeros((0, 2), np.uint32)
    start = 0
    for t, end16 in enumerate(tile_end16):
        nt = min(1024, n - 1024 * t)
        tile = enc[start : int(end16) * 16]
        nb = (nt + 31) // 32
        hdr_b, dst_b = (8 * nb + 15) & ~15, (4 * nt + 15) & ~15
        hdr = tile[:hdr_b].view(np.uint32)[: 2 * nb].reshape(nb, 2)
        dst = tile[hdr_b : hdr_b + dst_b].view(np.uint32)[:nt]
        src = tile[hdr_b + dst_b :].view(np.uint32)
        m = 0
        for b in range(nb):
            mask, base = int(hdr[b, 0]), int(hdr[b, 1])
            assert base == m, "block base must be the running count of moving records"
            for i in range(min(32, nt - 32 * b)):
                r = 1024 * t + 32 * b + i
                flat[r, 1] = dst[32 * b + i]
                if (mask >> i) & 1:
                    flat[r, 0] = src[m]
                    m += 1
                else:
                    flat[r, 0] = dst[32 * b + i]
        assert len(tile) == hdr_b + dst_b + ((4 * m + 15) & ~15)
        start = int(end16) * 16
    assert start == len(enc)
    return out


def synth_preset(config: int, seed: int) -> MvgenSpec:
    s = MvgenSpec()
    lib().mscan_synth_preset(C.byref(s), config, seed)
    return s


def synth_host(spec: MvgenSpec, frame0: int, n_frames: int, n_threads: int = 8):
    """(rec_count u32[F], rec_off u64[F+1], recs MV_DTYPE[N], pts f64[F]) generated on the host."""
    L = lib()
    cnt = np.zeros(n_frames, dtype=np.uint32)
    L.mscan_synth_host_counts(C.byref(spec), frame0, n_frames, _ptr(cnt), n_threads)
    off = np.zeros(n_frames + 1, dtype=np.uint64)
    np.cumsum(cnt, out=off[1:])
    recs = np.zeros(int(off[-1]), dtype=MV_DTYPE)
    pts = np.zeros(n_frames, dtype=np.float64)
    L.mscan_synth_host_fill(C.byref(spec), frame0, n_frames, _ptr(off), _ptr(recs), _ptr(pts), n_threads)
    return cnt, off, recs, pts


class Context:
    """One mscan_ctx (one GPU)."""

    def __init__(self, device: int = 0, params: Params | None = None, max_log_frames: int = 0, slab_bytes: int = 0):
        self.L = lib()
        self.params = params if params is not None else default_params()
        h = C.c_void_p()
        rc = self.L.mscan_create(device, C.byref(self.params), max_log_frames, slab_bytes, C.byref(h))
        if rc:
            raise MscanError(rc, self.L.mscan_status_string(rc).decode())
        self.h = h

    # -- helpers
    def _ck(self, rc: int):
        if rc:
            raise MscanError(rc, self.L.mscan_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.mscan_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- host-fed path
    def video_open(self, vid: int, width: int, height: int):
        self._ck(self.L.mscan_video_open(self.h, vid, width, height))

    def video_open_geometry(self, vid: int, g: Geometry):
        self._ck(self.L.mscan_video_open_geometry(self.h, vid, C.byref(g)))

    def submit(self, vid: int, pts, rec_count, recs) -> int:
        """Returns the video-local index of the first submitted frame."""
        n = len(rec_count)
        first = C.c_uint64()
        self._ck(self.L.mscan_submit(self.h, vid, n, _ptr(pts), _ptr(rec_count), _ptr(recs), C.byref(first)))
        return first.value

    def submit_raw(self, vid: int, n_frames: int, pts_ptr: int, cnt_ptr: int, recs_ptr: int):
        self._ck(self.L.mscan_submit(self.h, vid, n_frames, pts_ptr, cnt_ptr, recs_ptr, None))

    def submit_packed(self, vid: int, pts, rec_count, recs8) -> int:
        """Records already projected to MV8_DTYPE (pack_records)."""
        n = len(rec_count)
        first = C.c_uint64()
        self._ck(self.L.mscan_submit_packed(self.h, vid, n, _ptr(pts), _ptr(rec_count), _ptr(recs8), C.byref(first)))
        return first.value

    def submit_device(self, vid: int, pts, rec_count, d_recs: int, packed: bool = False, ready_stream: int = 0) -> int:
        """Frames whose records already lie in this GPU's memory (device pointer d_recs): scanned in place."""
        n = len(rec_count)
        first = C.c_uint64()
        self._ck(self.L.mscan_submit_device(self.h, vid, n, _ptr(pts), _ptr(rec_count), d_recs, int(packed), ready_stream or None, C.byref(first)))
        return first.value

    def submit_elided(self, vid: int, pts, rec_count, enc, enc_off, tile_end16) -> int:
        """Frames already in the static-elided form (elide_frames): enc uint8 array or address, enc_off u64[F+1]."""
        n = len(rec_count)
        first = C.c_uint64()
        enc_off = np.ascontiguousarray(enc_off, dtype=np.uint64)
        tile_end16 = np.ascontiguousarray(tile_end16, dtype=np.uint32)
        self._ck(self.L.mscan_submit_elided(self.h, vid, n, _ptr(pts), _ptr(rec_count), _ptr(enc), _ptr(enc_off), _ptr(tile_end16), C.byref(first)))
        return first.value

    def submit_packed_raw(self, vid: int, n_frames: int, pts_ptr: int, cnt_ptr: int, recs_ptr: int):
        self._ck(self.L.mscan_submit_packed(self.h, vid, n_frames, pts_ptr, cnt_ptr, recs_ptr, None))

    def set_staging_mode(self, mode: int):
        self._ck(self.L.mscan_set_staging_mode(self.h, mode))

    def reserve_staging(self):
        self._ck(self.L.mscan_reserve_staging(self.h))

    def set_pack_threads(self, n: int):
        self._ck(self.L.mscan_set_pack_threads(self.h, n))

    def collect_range(self, vid: int, first: int, n: int):
        flags = np.zeros(n, dtype=np.uint8)
        counts = np.zeros(n, dtype=np.uint32)
        self._ck(self.L.mscan_collect_range(self.h, vid, first, n, _ptr(flags), _ptr(counts)))
        return flags, counts

    def flush(self):
        self._ck(self.L.mscan_flush(self.h))

    def sync(self):
        self._ck(self.L.mscan_sync(self.h))

    def collect(self, vid: int):
        n = C.c_uint32()
        self._ck(self.L.mscan_collect(self.h, vid, None, None, 0, C.byref(n)))
        flags = np.zeros(n.value, dtype=np.uint8)
        counts = np.zeros(n.value, dtype=np.uint32)
        self._ck(self.L.mscan_collect(self.h, vid, _ptr(flags), _ptr(counts), n.value, C.byref(n)))
        return flags, counts

    def _segs(self, fn, vid: int, duration: float, cap: int):
        out = np.zeros(max(cap, 1), dtype=SEG_DTYPE)
        n = C.c_uint32()
        res = VideoResult()
        rc = fn(self.h, vid, duration, _ptr(out), cap, C.byref(n), C.byref(res))
        if rc == ERR_CAPACITY and n.value > cap:
            return self._segs(fn, vid, duration, n.value)
        self._ck(rc)
        return out[: n.value].copy(), res

    def segments(self, vid: int, duration: float, cap: int = 256):
        """FFmpegJob segments + result (CUT: clamped segments; FULL_COPY: [{0,duration}]; NO_MOTION: none)."""
        return self._segs(self.L.mscan_segments, vid, duration, cap)

    def motion_segments(self, vid: int, duration: float, cap: int = 256):
        return self._segs(self.L.mscan_motion_segments, vid, duration, cap)

    def segments_batch(self, vids, durations, cap: int = 0):
        vids = np.ascontiguousarray(vids, dtype=np.uint32)
        durations = np.ascontiguousarray(durations, dtype=np.float64)
        n = len(vids)
        cap = cap or 64 * max(n, 1)
        out = np.zeros(cap, dtype=SEG_DTYPE)
        off = np.zeros(n + 1, dtype=np.uint64)
        res = np.zeros(n, dtype=RESULT_DTYPE)
        rc = self.L.mscan_segments_batch(self.h, n, _ptr(vids), _ptr(durations), _ptr(out), cap, _ptr(off), _ptr(res))
        if rc == ERR_CAPACITY and int(off[n]) > cap:
            return self.segments_batch(vids, durations, int(off[n]))
        self._ck(rc)
        return out[: int(off[n])].copy(), off, res

    def video_append_from(self, vid: int, src: "Context", src_vid: int):
        """Cross-GPU stitch: this context's video `vid` adopts the frame results of (src, src_vid)."""
        self._ck(self.L.mscan_video_append_from(self.h, vid, src.h, src_vid))

    def video_close(self, vid: int):
        self._ck(self.L.mscan_video_close(self.h, vid))

    def host_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._ck(self.L.mscan_host_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def host_free(self, p: int):
        self._ck(self.L.mscan_host_free(self.h, p))

    def host_register(self, a: np.ndarray, read_only: bool = False):
        self._ck(self.L.mscan_host_register(self.h, a.ctypes.data, a.nbytes, int(read_only)))

    def host_unregister(self, a: np.ndarray):
        self._ck(self.L.mscan_host_unregister(self.h, a.ctypes.data))

    def host_fence(self):
        self._ck(self.L.mscan_host_fence(self.h))

    def pinned_array(self, shape, dtype):
        """numpy view over pinned host memory (freed with host_free(arr.ctypes.data))."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = self.host_alloc(max(n, 1))
        buf = (C.c_uint8 * max(n, 1)).from_address(p)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    # -- device-resident path
    def dev_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._ck(self.L.mscan_dev_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def dev_free(self, p: int):
        self._ck(self.L.mscan_dev_free(self.h, p))

    def h2d(self, d: int, a: np.ndarray):
        self._ck(self.L.mscan_memcpy_h2d(self.h, d, _ptr(a), a.nbytes))

    def d2h(self, a: np.ndarray, d: int):
        self._ck(self.L.mscan_memcpy_d2h(self.h, _ptr(a), d, a.nbytes))

    def offsets_from_counts(self, d_cnt: int, n: int, d_off: int, stream: int = 0):
        self._ck(self.L.mscan_offsets_from_counts(self.h, d_cnt, n, d_off, stream))

    def scan_device(self, d_recs, d_off, d_frame_geom, geoms, n_frames, d_flags, d_counts, stream: int = 0):
        arr = (Geometry * len(geoms))(*geoms)
        self._ck(
            self.L.mscan_scan_device(self.h, d_recs, d_off, d_frame_geom, arr, len(geoms), n_frames, d_flags, d_counts, stream)
        )

    def scan_device_packed(self, d_recs8, d_off, d_frame_geom, geoms, n_frames, d_flags, d_counts, stream: int = 0):
        arr = (Geometry * len(geoms))(*geoms)
        self._ck(
            self.L.mscan_scan_device_packed(
                self.h, d_recs8, d_off, d_frame_geom, arr, len(geoms), n_frames, d_flags, d_counts, stream
            )
        )

    def pack_records_device(self, d_recs: int, n: int, d_out: int, stream: int = 0):
        self._ck(self.L.mscan_pack_records_device(self.h, d_recs, n, d_out, stream))

    def segments_device(self, video_off, durations, d_pts, d_flags, d_segs, d_res, stream: int = 0):
        video_off = np.ascontiguousarray(video_off, dtype=np.uint64)
        durations = np.ascontiguousarray(durations, dtype=np.float64)
        self._ck(
            self.L.mscan_segments_device(
                self.h, len(durations), _ptr(video_off), _ptr(durations), d_pts, d_flags, d_segs, d_res, stream
            )
        )

    def synth_counts(self, spec: MvgenSpec, frame0: int, n: int, d_cnt: int, stream: int = 0):
        self._ck(self.L.mscan_synth_counts(self.h, C.byref(spec), frame0, n, d_cnt, stream))

    def synth_fill(self, spec: MvgenSpec, frame0: int, n: int, d_off: int, d_recs: int, d_pts: int, stream: int = 0):
        self._ck(self.L.mscan_synth_fill(self.h, C.byref(spec), frame0, n, d_off, d_recs, d_pts, stream))

    def stats(self) -> Stats:
        s = Stats()
        self._ck(self.L.mscan_get_stats(self.h, C.byref(s)))
        return s

    def reset_stats(self):
        self._ck(self.L.mscan_reset_stats(self.h))

    def set_profiling(self, on: bool):
        self._ck(self.L.mscan_set_profiling(self.h, int(on)))


# ---- measurement harness: decode-worker stand-in threads (csrc/feed_harness.cpp → libmscan_feed.so) -------------
FEED_LIB_PATH = PKG_DIR / "libmscan_feed.so"


class FeedSpec(C.Structure):
    _fields_ = [
        ("n_threads", C.c_int32),
        ("cpus", C.c_void_p),
        ("n_videos", C.c_uint32),
        ("video_ids", C.c_void_p),
        ("video_frame_off", C.c_void_p),
        ("pts", C.c_void_p),
        ("rec_count", C.c_void_p),
        ("rec_off", C.c_void_p),
        ("source", C.c_void_p),
        ("source_kind", C.c_int32),
        ("frames_per_submit", C.c_uint32),
        ("submit_kind", C.c_int32),
        ("frame_index_out", C.c_void_p),
    ]


class FeedResult(C.Structure):
    _fields_ = [
        ("wall_s", C.c_double),
        ("hot_max_s", C.c_double),
        ("hot_sum_s", C.c_double),
        ("standin_max_s", C.c_double),
        ("standin_sum_s", C.c_double),
        ("frames", C.c_uint64),
        ("records", C.c_uint64),
        ("submits", C.c_uint64),
        ("rc", C.c_int32),
        ("avx512", C.c_int32),
    ]


_feed = None


def feed_lib() -> C.CDLL:
    global _feed
    if _feed is None:
        lib()  # libmotionscan.so first: the harness links against it
        if not FEED_LIB_PATH.exists():
            raise FileNotFoundError(f"{FEED_LIB_PATH} is missing — build it with `python {PKG_DIR.name}/build.py`")
        L = C.CDLL(str(FEED_LIB_PATH))
        L.mscan_feed_run.restype = C.c_int
        L.mscan_feed_run.argtypes = [C.c_void_p, C.POINTER(FeedSpec), C.POINTER(FeedResult)]
        L.mscan_feed_expand.restype = C.c_int
        L.mscan_feed_expand.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        _feed = L
    return _feed


def feed_expand(recs8: np.ndarray) -> np.ndarray:
    """The stand-in decoder's record writer: native records (MV_DTYPE) from MV8_DTYPE coordinates."""
    assert recs8.dtype == MV8_DTYPE and recs8.flags["C_CONTIGUOUS"]
    out = np.zeros(len(recs8), dtype=MV_DTYPE)
    rc = feed_lib().mscan_feed_expand(_ptr(recs8), len(recs8), _ptr(out))
    if rc:
        raise MscanError(rc, "mscan_feed_expand")
    return out


def feed_run(ctx: "Context", video_ids, video_frame_off, pts, rec_count, rec_off, source, *, n_threads: int, cpus=None,
             frames_per_submit: int = 1, submit_kind: int = 0, want_index: bool = True):
    """T decode-worker stand-ins feed `ctx` (see csrc/feed_harness.cpp). `source`: MV8_DTYPE (expanded to native
    records per frame by the stand-in) or MV_DTYPE (copied). Returns (FeedResult, frame_index or None)."""
    video_ids = np.ascontiguousarray(video_ids, dtype=np.uint32)
    video_frame_off = np.ascontiguousarray(video_frame_off, dtype=np.uint64)
    rec_count = np.ascontiguousarray(rec_count, dtype=np.uint32)
    rec_off = np.ascontiguousarray(rec_off, dtype=np.uint64)
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    assert source.dtype in (MV8_DTYPE, MV_DTYPE) and source.flags["C_CONTIGUOUS"]
    n_frames = int(video_frame_off[-1])
    assert len(pts) >= n_frames and len(rec_count) >= n_frames and len(rec_off) >= n_frames + 1
    cpu_arr = np.ascontiguousarray(cpus, dtype=np.int32) if cpus is not None else None
    assert cpu_arr is None or len(cpu_arr) >= n_threads
    index = np.zeros(n_frames, dtype=np.uint64) if want_index else None
    sp = FeedSpec(
        n_threads=n_threads,
        cpus=_ptr(cpu_arr),
        n_videos=len(video_ids),
        video_ids=_ptr(video_ids),
        video_frame_off=_ptr(video_frame_off),
        pts=_ptr(pts),
        rec_count=_ptr(rec_count),
        rec_off=_ptr(rec_off),
        source=_ptr(source),
        source_kind=1 if source.dtype == MV8_DTYPE else 0,
        frames_per_submit=frames_per_submit,
        submit_kind=submit_kind,
        frame_index_out=_ptr(index),
    )
    res = FeedResult()
    rc = feed_lib().mscan_feed_run(ctx.h, C.byref(sp), C.byref(res))
    if rc:
        raise MscanError(rc, ctx.L.mscan_last_error(ctx.h).decode())
    return res, index
