#!/usr/bin/env python
"""Small driver for ncu captures of K-A: builds a device-resident stream and launches ka_scan_kernel a few times.
    python tools/ka_profile_run.py [records] [native|packed] [preset] [launches] [frames]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "motion-estimated-video-trimmer_b200"))
import motionscan as ms  # noqa: E402

records = float(sys.argv[1]) if len(sys.argv) > 1 else 4e8
layout = sys.argv[2] if len(sys.argv) > 2 else "native"
preset = int(sys.argv[3]) if len(sys.argv) > 3 else 4
launches = int(sys.argv[4]) if len(sys.argv) > 4 else 5
p = ms.shipped_env_params()
ctx = ms.Context(0, p, 1 << 16, 64 << 20)
spec = ms.synth_preset(preset, 5)
probe = 2048
d = ctx.dev_alloc(4 * probe)
ctx.synth_counts(spec, 0, probe, d)
ctx.sync()
pc = np.zeros(probe, np.uint32)
ctx.d2h(pc, d)
n = int(records / pc.mean())
if len(sys.argv) > 5 and int(sys.argv[5]) > 0:  # exact frame count (bench.py's stream: 100153 frames of preset 4, seed 5)
    n = int(sys.argv[5])
d_cnt = ctx.dev_alloc(4 * n)
d_off = ctx.dev_alloc(8 * (n + 1))
ctx.synth_counts(spec, 0, n, d_cnt)
ctx.offsets_from_counts(d_cnt, n, d_off)
ctx.sync()
off = np.zeros(n + 1, np.uint64)
ctx.d2h(off, d_off)
nrec = int(off[-1])
d_recs = ctx.dev_alloc(40 * nrec + 256)
d_fl = ctx.dev_alloc(n)
d_ct = ctx.dev_alloc(4 * n)
ctx.synth_fill(spec, 0, n, d_off, d_recs, 0)
ctx.sync()
g = ms.geometry_from_dims(p, spec.width, spec.height)
scan, d_in, rb = ctx.scan_device, d_recs, 40
if layout == "packed":
    d_r8 = ctx.dev_alloc(8 * nrec + 256)
    ctx.pack_records_device(d_recs, nrec, d_r8)
    ctx.sync()
    scan, d_in, rb = ctx.scan_device_packed, d_r8, 8
if layout == "elided":  # host-fed: per-frame submits in the static-elided form → K-A<mvz> launches on the slab streams
    n = min(n, 6000)
    off_h = off[: n + 1]
    recs_h = np.zeros(int(off_h[-1]), ms.MV_DTYPE)
    ctx.d2h(recs_h, d_recs)
    cnt_h = np.diff(off_h).astype(np.uint32)
    pts_h = np.arange(n) / 30.0
    ctx.set_staging_mode(ms.STAGING_ELIDE)
    ctx.set_profiling(True)
    for rep in range(launches):
        ctx.video_open(1, spec.width, spec.height)
        ctx.submit(1, pts_h, cnt_h, recs_h)
        fl, _ = ctx.collect(1)
        ctx.video_close(1)
    st = ctx.stats()
    ms_ = st.scan_ms / st.scan_launches
    print(f"elided preset {preset}: {int(off_h[-1])} records, {n} frames per pass, {st.scan_launches} K-A<mvz> launches, {ms_:.3f} ms/launch, "
          f"{st.elided_bytes / st.records_elided:.2f} B/record, {st.records_elided / (st.scan_ms * 1e-3) / 1e9:.1f} G rec/s in the kernel, active {int(fl.sum())}")
    sys.exit(0)
ctx.set_profiling(True)
for _ in range(launches):
    scan(d_in, d_off, None, [g], n, d_fl, d_ct)
ctx.sync()
st = ctx.stats()
ms_ = st.scan_ms / st.scan_launches
print(f"{layout} preset {preset}: {nrec} records, {n} frames, {ms_:.3f} ms/launch, {(rb * nrec + 17 * n) / ms_ / 1e6:.0f} GB/s, {nrec / ms_ / 1e6:.1f} G rec/s")
