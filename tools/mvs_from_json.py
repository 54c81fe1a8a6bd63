#!/usr/bin/env python
"""extract_mvs JSON → MVS1 stream file.

The reference's tools/extract_mvs.cpp (:96-176) dumps, per decoded frame, `frame_index`, `pts_seconds`,
`frame_type`, `num_mvs` and per vector `dst_x,dst_y,src_x,src_y (floats, recomputed),w,h,motion_x,
motion_y,motion_scale,source`, plus the stream `time_base`. It does NOT dump the int16 `src_x/src_y`
fields check_frame reads (motion_scanner.cpp:246-247) nor the picture size, so:
  * src is rebuilt the way FFmpeg's export fills it: src = dst + motion / motion_scale with C integer
    division (truncation toward zero) — the JSON's float src_x/src_y are only used as a cross-check;
  * --width/--height must be given (they size the block grid, motion_scanner.cpp:189-192).
A dump made on any machine that has FFmpeg thus becomes an input for motion_trim_b200 / the tests.

    python tools/mvs_from_json.py dump.json out.mvs --width 1920 --height 1080 [--fps 30]
"""
import argparse
import json
import sys
from fractions import Fraction
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "motion-estimated-video-trimmer_b200")]


def c_div(a: int, b: int) -> int:
    """C integer division (truncates toward zero)."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def convert(doc: dict, width: int, height: int, fps: Fraction | None = None):
    import motionscan as ms

    tb = Fraction(doc["time_base"])
    frames = doc["frames"]
    ticks, counts, has, key, recs = [], [], [], [], []
    for fr in frames:
        pts = fr.get("pts_seconds")
        ticks.append(0 if pts is None else int(round(Fraction(str(pts)) / tb)))
        mvs = fr.get("motion_vectors") or []
        assert len(mvs) == fr.get("num_mvs", len(mvs)), f"frame {fr.get('frame_index')}: num_mvs mismatch"
        counts.append(len(mvs))
        has.append(len(mvs) > 0)
        key.append(fr.get("frame_type") == "I")
        for mv in mvs:
            scale = mv["motion_scale"] or 1
            sx = mv["dst_x"] + c_div(mv["motion_x"], scale)
            sy = mv["dst_y"] + c_div(mv["motion_y"], scale)
            if abs(sx - mv["src_x"]) >= 1 or abs(sy - mv["src_y"]) >= 1:
                raise ValueError(f"frame {fr.get('frame_index')}: src rebuilt from motion_x/scale disagrees with the dump")
            recs.append((mv["source"], mv["w"], mv["h"], sx, sy, mv["dst_x"], mv["dst_y"], 0, mv["motion_x"], mv["motion_y"],
                         mv["motion_scale"]))
    arr = np.zeros(len(recs), dtype=ms.MV_DTYPE)
    for i, name in enumerate(ms.MV_DTYPE.names):
        arr[name] = [r[i] for r in recs] if recs else []
    key = np.array(key, bool)
    if len(key):
        key[0] = True
    if fps is None:  # avg frame rate from the pts span
        span = (ticks[-1] - ticks[0]) * tb if len(ticks) > 1 else Fraction(0)
        fps = Fraction(len(ticks) - 1, 1) / span if span > 0 else Fraction(25)
    fps = Fraction(fps).limit_denominator(100000)
    return dict(width=width, height=height, fps=(fps.numerator, fps.denominator), tb=(tb.numerator, tb.denominator),
                ticks=np.array(ticks, np.int64), counts=np.array(counts, np.uint32), has=np.array(has, bool), key=key, recs=arr)


def main():
    from motionscan import mvs_io

    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("json")
    ap.add_argument("out")
    ap.add_argument("--width", type=int, required=True)
    ap.add_argument("--height", type=int, required=True)
    ap.add_argument("--fps", type=str, default=None, help="avg frame rate, e.g. 30 or 30000/1001 (default: from pts)")
    a = ap.parse_args()
    d = convert(json.loads(Path(a.json).read_text()), a.width, a.height, Fraction(a.fps) if a.fps else None)
    dur_us = None
    if len(d["ticks"]):
        tb = Fraction(*d["tb"])
        dur_us = int((d["ticks"][-1] - d["ticks"][0]) * tb * 1_000_000 + Fraction(1_000_000) / Fraction(*d["fps"]))
    mvs_io.write_mvs(a.out, d["width"], d["height"], d["fps"][0], d["fps"][1], d["ticks"], d["counts"], d["recs"], has_mvs=d["has"],
                     key=d["key"], tb_num=d["tb"][0], tb_den=d["tb"][1], duration_us=dur_us)
    print(f"{a.out}: {len(d['ticks'])} frames, {len(d['recs'])} records")


if __name__ == "__main__":
    main()
