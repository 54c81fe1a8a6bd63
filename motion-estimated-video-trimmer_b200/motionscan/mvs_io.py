"""MVS1 motion-vector stream files (include/mvs_format.h): writer/reader used by the tests, the
reference runner and the golden-fixture generator."""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

MAGIC = b"MVSTRM01"
HDR = struct.Struct("<8siiiiiiqIIQQ")  # 64 bytes
FRAME_DTYPE = np.dtype([("pts", "<i8"), ("first_record", "<u8"), ("n_records", "<u4"), ("flags", "<u4")])
KEY, HAS_MVS = 1, 2
assert HDR.size == 64 and FRAME_DTYPE.itemsize == 24


def write_mvs(path, width, height, fps_num, fps_den, pts_ticks, rec_count, recs, has_mvs=None, key=None,
              tb_num=None, tb_den=None, duration_us=None):
    """pts_ticks in time_base units (default time_base = fps_den/fps_num, i.e. one tick per frame)."""
    n = len(rec_count)
    tb_num = fps_den if tb_num is None else tb_num
    tb_den = fps_num if tb_den is None else tb_den
    rec_count = np.asarray(rec_count, dtype=np.uint32)
    if has_mvs is None:
        has_mvs = rec_count > 0
    if key is None:
        key = np.zeros(n, dtype=bool)
        key[0:1] = True
        key |= ~np.asarray(has_mvs, dtype=bool)  # I-frames carry no vectors and are the seek points
    if duration_us is None:
        duration_us = -(-n * 1_000_000 * fps_den // fps_num)  # ceil(n / fps) in AV_TIME_BASE units
    fr = np.zeros(n, dtype=FRAME_DTYPE)
    fr["pts"] = np.asarray(pts_ticks, dtype=np.int64)
    off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(rec_count, out=off[1:])
    fr["first_record"] = off[:-1]
    fr["n_records"] = rec_count
    fr["flags"] = np.where(key, KEY, 0) | np.where(has_mvs, HAS_MVS, 0)
    rec_off = (HDR.size + fr.nbytes + 63) & ~63
    n_rec = int(off[-1])
    with open(path, "wb") as f:
        f.write(HDR.pack(MAGIC, width, height, tb_num, tb_den, fps_num, fps_den, int(duration_us), n, 0, rec_off, n_rec))
        f.write(fr.tobytes())
        f.write(b"\0" * (rec_off - HDR.size - fr.nbytes))
        if n_rec:
            recs = np.ascontiguousarray(recs)
            assert recs.dtype.itemsize == 40 and len(recs) == n_rec
            f.write(recs.tobytes())
    return duration_us / 1_000_000.0


def read_mvs(path, mv_dtype):
    raw = Path(path).read_bytes()
    magic, w, h, tbn, tbd, fn, fd, dur, n, _, rec_off, n_rec = HDR.unpack_from(raw, 0)
    assert magic == MAGIC
    fr = np.frombuffer(raw, dtype=FRAME_DTYPE, count=n, offset=HDR.size)
    recs = np.frombuffer(raw, dtype=mv_dtype, count=n_rec, offset=rec_off)
    return dict(width=w, height=h, tb=(tbn, tbd), fps=(fn, fd), duration_us=dur, frames=fr, recs=recs)
