#!/usr/bin/env python
"""BASELINE.json configs[3] through the product CLI: a batch of 64 synthetic 60 s 1080p30 clips (MVS1 files,
seeds 100..163) handed to `motion_trim_b200 <dir> <out>`, whose BatchProcessor deals the files to
PARALLEL_STREAMS stream threads per GPU (batch_processor.cpp role) — the videos are sharded over the GPUs,
nothing is exchanged between them. Reports wall time and records/s for 1, 2, 4, 8 GPUs (those the box has)
(the mapped records are projected by the library's staging pass: cudaHostRegister refuses tmpfs/page-cache
mappings on this platform, so the in-place DMA of mapped files does not apply), with the mapping lazy (default)
or populated up front like the reference's.

    python tools/batch64_cli.py [--clips 64] [--frames 1800] [--dir /dev/shm/mscan_batch64]
"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "motion-estimated-video-trimmer_b200"), str(ROOT / "tests")]
import motionscan as ms  # noqa: E402
from motionscan import mvs_io  # noqa: E402

BIN = ROOT / "motion-estimated-video-trimmer_b200" / "host" / "motion_trim_b200"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=64)
    ap.add_argument("--frames", type=int, default=0, help="frames per clip; 0 = uniform 30–120 s per clip (SURVEY §8(d) config 4)")
    ap.add_argument("--dir", default="/dev/shm/mscan_batch64")
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("--streams", default="2", help="PARALLEL_STREAMS per GPU (shipped env: 2; comma list)")
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--modes", default="default,nopin")
    ap.add_argument("--chunk", default="10", help="CHUNK_DURATION_SEC (comma list: one run per value)")
    ap.add_argument("--threads", default="0", help="THREADS_PER_STREAM (0 = auto: CPUs / streams, like the reference; comma list)")
    ap.add_argument("--trace", action="store_true", help="MSCAN_TRACE=1: print the library's per-entry-point wall times of each run")
    args = ap.parse_args()
    import ctypes as C

    n_dev = C.c_int()
    ms.lib().mscan_device_count(C.byref(n_dev))
    ind, threads = Path(args.dir) / "in", os.cpu_count() or 8
    ind.mkdir(parents=True, exist_ok=True)
    t0, n_rec, expect = time.time(), 0, {}
    sys.path.insert(0, str(ROOT))
    from bench import batch64_lengths

    lengths = [args.frames] * args.clips if args.frames else batch64_lengths(100, args.clips)
    for k in range(args.clips):
        spec = ms.synth_preset(3, 100 + k)
        spec.frames_per_video = lengths[k]
        cnt, off, recs, pts = ms.synth_host(spec, 0, lengths[k], n_threads=threads)
        mvs_io.write_mvs(ind / f"clip{k:03d}.mvs", spec.width, spec.height, int(spec.fps), 1, np.arange(lengths[k]), cnt, recs)
        n_rec += int(off[-1])
    print(f"# generated {args.clips} clips of {min(lengths)}–{max(lengths)} frames ({sum(lengths)} in all) = {n_rec} records "
          f"({n_rec * 40 / 1e9:.1f} GB) in {time.time() - t0:.1f} s", flush=True)
    baseline = None
    combos = [(g, t, c, st) for g in [int(x) for x in args.gpus.split(",") if int(x) <= max(n_dev.value, 1)] for t in args.threads.split(",")
              for c in args.chunk.split(",") for st in args.streams.split(",")]
    for g, thr, chunk, n_streams in combos:
        env0 = dict(os.environ, MV_THRESHOLD_SQ="4.0", VECTORS_NEEDED="4", CLUSTERS_NEEDED="2", VERTICAL_MASK="0.05", MAX_GAP_SEC="5",
                    PADDING_SEC="0.5", MIN_SAVINGS_PCT="5", CHUNK_DURATION_SEC=chunk, TARGET_FPS="0", THREADS_PER_STREAM=thr,
                    PARALLEL_STREAMS=n_streams)
        if args.trace:
            env0["MSCAN_TRACE"] = "1"
        # default: the mapped file is pinned and DMA'd in place (40 B/record over PCIe, no host pass); nopin: the chunk
        # workers project their chunks into the library's pinned ring (8 B/record over PCIe); populate: MAP_POPULATE
        for mode, extra in (("default", {}), ("nopin", {"MOTION_TRIM_NO_PIN": "1"}), ("populate", {"MOTION_TRIM_POPULATE": "1"}),
                            ("compact", {"MOTION_TRIM_NO_PIN": "1", "MOTION_TRIM_STAGING": "compact"})):
            if mode not in args.modes.split(","):
                continue
            outd = Path(args.dir) / f"out_{g}_{mode}"
            shutil.rmtree(outd, ignore_errors=True)
            t = time.time()
            r = subprocess.run([str(BIN), "--print-segments", str(ind), str(outd)], env=dict(env0, MOTION_TRIM_GPUS=str(g), **extra),
                               capture_output=True, text=True)
            wall = time.time() - t
            m = re.search(r"wall ([0-9.]+)s", r.stdout)
            scan_wall = float(m.group(1)) if m else float("nan")
            res = {ln.split()[1]: ln.split()[2:] for ln in r.stdout.splitlines() if ln.startswith("RESULT ")}
            dec = {k: [x for x in v if x.startswith("decision=") or x.startswith("saved_pct=")] for k, v in sorted(res.items())}
            if baseline is None:
                baseline = dec
            ph = re.search(r"phases \(sum over files, s\): (.*)", r.stdout)
            if args.trace:
                print("\n".join(ln for ln in r.stderr.splitlines() if "mscan trace" in ln), flush=True)
            print(json.dumps({"gpus": g, "feed": mode, "threads_per_stream": thr, "chunk_sec": chunk, "phases": ph.group(1) if ph else None, "streams_per_gpu": n_streams, "rc": r.returncode, "files": len(res),
                              "batch_wall_s": scan_wall, "process_wall_s": round(wall, 3), "records_per_s": n_rec / scan_wall,
                              "same_results_as_first_run": dec == baseline}), flush=True)
    if not args.keep:
        shutil.rmtree(args.dir, ignore_errors=True)


if __name__ == "__main__":
    main()
