#!/usr/bin/env python
"""Diagnosis: producers feeding with MSCAN_STAGING_ELIDE vs AUTO (one GPU) — library stats per run."""
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "motion-estimated-video-trimmer_b200"))
import motionscan as ms  # noqa: E402

p = ms.shipped_env_params()
spec = ms.synth_preset(4, 5)
n_frames = 3000
cnt, off, recs, pts = ms.synth_host(spec, 0, n_frames, n_threads=16)
r8 = ms.pack_records(recs)
n = int(off[-1])
voff = np.array([0, n_frames], np.uint64)
cpus = sorted(os.sched_getaffinity(0))
with ms.Context(0, p, 1 << 20, 64 << 20) as ctx:
    ctx.set_profiling(True)
    for mode, name in ((ms.STAGING_AUTO, "auto"), (ms.STAGING_ELIDE, "elide")):
        ctx.set_staging_mode(mode)
        for T in [int(x) for x in (sys.argv[1].split(',') if len(sys.argv) > 1 else ['1', '4', '16'])]:
            if T > len(cpus):
                continue
            for rep in range(2):
                ctx.video_open(1, spec.width, spec.height)
                ctx.reset_stats()
                t0 = time.perf_counter()
                res, _ = ms.feed_run(ctx, [1], voff, pts, cnt, off, r8, n_threads=T, cpus=cpus[:T], want_index=False)
                t1 = time.perf_counter()
                ctx.collect(1)
                dt = time.perf_counter() - t0
                st = ctx.stats()
                ctx.video_close(1)
            print(f"{name:5s} T={T:2d}: {n / dt / 1e9:6.2f} G rec/s wall (feed {t1 - t0:.4f}s, tail {dt - (t1 - t0):.4f}s); thread time: stand-in {res.standin_sum_s / T * 1e3:7.2f} ms, "
                  f"submit {res.hot_sum_s / T * 1e3:7.2f} ms; K-A launches {st.scan_launches}, K-A total {st.scan_ms:.3f} ms; project cpu {st.project_ms:.1f} ms; "
                  f"h2d {st.h2d_bytes / 1e6:.1f} MB; elided {st.records_elided}", flush=True)
