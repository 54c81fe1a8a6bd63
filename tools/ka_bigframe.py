#!/usr/bin/env python
"""K-A on frames larger than 4K (8K: 480x270 = 129 600 cells, 16K-wide strips): which counter plan runs and how fast.
Usage: python tools/ka_bigframe.py [width height frames]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "motion-estimated-video-trimmer_b200")]
import numpy as np  # noqa: E402

import motionscan as ms  # noqa: E402

w, h, n = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (7680, 4320, 400)
p = ms.shipped_env_params()
ctx = ms.Context(0, p, 1 << 16, 64 << 20)
spec = ms.synth_preset(2, 3)
spec.width, spec.height = w, h
d_cnt, d_off = ctx.dev_alloc(4 * n), ctx.dev_alloc(8 * (n + 1))
ctx.synth_counts(spec, 0, n, d_cnt)
ctx.offsets_from_counts(d_cnt, n, d_off)
ctx.sync()
off = np.zeros(n + 1, np.uint64)
ctx.d2h(off, d_off)
nrec = int(off[-1])
d_recs, d_fl, d_ct = ctx.dev_alloc(40 * nrec + 256), ctx.dev_alloc(n), ctx.dev_alloc(4 * n)
ctx.synth_fill(spec, 0, n, d_off, d_recs, 0)
ctx.sync()
g = ms.geometry_from_dims(p, w, h)
for layout in ("native", "packed"):
    scan, d_in, rb = ctx.scan_device, d_recs, 40
    if layout == "packed":
        d_r8 = ctx.dev_alloc(8 * nrec + 256)
        ctx.pack_records_device(d_recs, nrec, d_r8)
        ctx.sync()
        scan, d_in, rb = ctx.scan_device_packed, d_r8, 8
    for _ in range(3):
        scan(d_in, d_off, None, [g], n, d_fl, d_ct)
    ctx.sync()
    ctx.reset_stats()
    ctx.set_profiling(True)
    for _ in range(10):
        scan(d_in, d_off, None, [g], n, d_fl, d_ct)
    ctx.sync()
    st = ctx.stats()
    ctx.set_profiling(False)
    t = st.scan_ms / st.scan_launches
    fl = np.zeros(n, np.uint8)
    ctx.d2h(fl, d_fl)
    print(f"{w}x{h} grid {g.grid_w}x{g.grid_h} ({g.grid_w * g.grid_h} cells) {layout}: {n} frames, {nrec} records, {t:.3f} ms/launch, "
          f"{(rb * nrec + 17 * n) / t / 1e6:.0f} GB/s, {nrec / t / 1e6:.1f} G rec/s, {n / t * 1e3:.0f} frames/s, active {int(fl.sum())}")
