#!/usr/bin/env python
"""Generates tests/golden/ref_golden.json by running the REFERENCE's own code (oracle/_ref/ref_scan =
/root/reference/src/{motion_scanner,pipeline,memory_io,task_queue,ffmpeg_queue,logging,system}.cpp
compiled unmodified against the fake-libav shim) over the cases in tests/golden_cases.py.

    make -C oracle -f ref.mk && python tests/golden/make_golden.py

Needs /root/reference (only present in the build container). The fixture stores, per case, the
input digest and the reference's outputs: timestamps with motion (scan_range), the FFmpegJob
segments, decision, duration, time_removed and saved_pct — doubles as hex strings, exact. `ts` comes from ONE
scan_range(0, duration) call; the job comes from the chunked, multi-threaded pipeline (the two differ
when TARGET_FPS skips frames, because the skip counter restarts at every chunk's seek key frame).
"""
import json
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path[:0] = [str(ROOT / "motion-estimated-video-trimmer_b200"), str(ROOT / "tests")]

import golden_cases as gc  # noqa: E402
from motionscan import mvs_io  # noqa: E402
import ref_runner  # noqa: E402


def main():
    assert ref_runner.available(), "build oracle/_ref first: make -C oracle -f ref.mk"
    out = {"generator": "tests/golden/make_golden.py", "reference": "oracle/_ref/ref_scan (reference sources + fake-libav shim)",
           "cases": {}}
    with tempfile.TemporaryDirectory() as d:
        for c in gc.all_cases():
            path = Path(d) / (c.name + ".mvs")
            mvs_io.write_mvs(path, c.width, c.height, c.fps[0], c.fps[1], c.ticks, c.cnt, c.recs, has_mvs=c.has_mvs,
                             tb_num=c.tb[0], tb_den=c.tb[1], duration_us=c.duration_us)
            r = ref_runner.run(path, c.params, threads=c.threads, chunk_sec=c.chunk_sec, target_fps=c.target_fps)
            assert r["scan_ok"] == 1 and r["run_rc"] == 0, c.name
            assert r["duration"] == c.duration, (c.name, r["duration"], c.duration)
            out["cases"][c.name] = {
                "digest": gc.digest(c.cnt, c.recs, c.ticks),
                "n_frames": int(len(c.cnt)),
                "n_records": int(c.cnt.sum()),
                "params": gc.params_dict(c.params),
                "chunk_sec": c.chunk_sec,
                "threads": c.threads,
                "target_fps": c.target_fps,
                "duration": float(r["duration"]).hex(),
                "ts": [float(t).hex() for t in r["ts"]],
                "segments": [[float(a).hex(), float(b).hex()] for a, b in r["segs"]],
                "decision": int(r["decision"]),
                "time_removed": float(r["time_removed"]).hex(),
                "saved_pct": float(r["saved_pct"]).hex(),
            }
            path.unlink()
            print(f"{c.name:28s} frames {len(c.cnt):5d} records {int(c.cnt.sum()):9d} motion {len(r['ts']):4d} "
                  f"segs {len(r['segs']):2d} decision {r['decision']} saved {r['saved_pct']:.2f}%")
    (ROOT / "tests" / "golden" / "ref_golden.json").write_text(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
