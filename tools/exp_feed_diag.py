#!/usr/bin/env python
"""Diagnosis of the host side of the producers e2e mode (one GPU): how the decode stand-in, the projection and the
full mscan_submit scale with the number of producer threads (csrc/feed_harness.cpp submit_kind 2 / 3 / 0), and the
CPU topology the threads are pinned to."""
import glob
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "motion-estimated-video-trimmer_b200"))
import motionscan as ms  # noqa: E402

for f in sorted(glob.glob("/sys/devices/system/cpu/cpu[0-9]*/topology/thread_siblings_list"))[:40]:
    print(f.split("/")[5], open(f).read().strip(), end=" | ")
print()
p = ms.shipped_env_params()
spec = ms.synth_preset(4, 5)
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
cnt, off, recs, pts = ms.synth_host(spec, 0, n_frames, n_threads=16)
r8 = ms.pack_records(recs)
n = int(off[-1])
voff = np.array([0, n_frames], np.uint64)
cpus = sorted(os.sched_getaffinity(0))
with ms.Context(0, p, 1 << 20, 64 << 20) as ctx:
    for kind, name in ((2, "stand-in only"), (3, "stand-in + pack (private buffer)"), (0, "stand-in + mscan_submit")):
        for T in (1, 2, 4, 8, 12, 16):
            if T > len(cpus):
                continue
            best = 0
            for rep in range(3):
                ctx.video_open(1, spec.width, spec.height)
                t0 = time.perf_counter()
                res, _ = ms.feed_run(ctx, [1], voff, pts, cnt, off, r8, n_threads=T, cpus=cpus[:T], submit_kind=kind, want_index=False)
                if kind == 0:
                    ctx.collect(1)
                dt = time.perf_counter() - t0
                ctx.video_close(1)
                best = max(best, n / dt / 1e9)
            print(f"{name:34s} T={T:2d}: {best:6.2f} G rec/s total, {best / T:5.2f} per thread; thread time: stand-in {res.standin_sum_s / T * 1e3:7.2f} ms, "
                  f"call {res.hot_sum_s / T * 1e3:7.2f} ms (avx512={res.avx512})", flush=True)
