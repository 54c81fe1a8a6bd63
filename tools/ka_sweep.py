#!/usr/bin/env python
"""K-A tuning sweep: times ka_scan_kernel (CUDA events inside the library) on a device-resident stream for
several (CTAs/SM, ring stages, consumer warps) plans chosen through MSCAN_KA_CTAS / MSCAN_KA_STAGES /
MSCAN_KA_WARPS (a plan that does not fit shared memory falls back to the automatic one). One process per
plan (the plan is read when the context scans). Usage: python tools/ka_sweep.py [records] [workload]"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

CHILD = r"""
import sys, os, json
sys.path[:0] = [os.path.join(ROOT, "motion-estimated-video-trimmer_b200")]
import numpy as np, motionscan as ms
records, preset, seed, fixed = float(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
packed = len(sys.argv) > 5 and sys.argv[5] == "packed"
p = ms.shipped_env_params()
ctx = ms.Context(0, p, 1 << 16, 64 << 20)
spec = ms.synth_preset(preset, seed)
if fixed:
    n = fixed
else:
    probe = 2048
    d = ctx.dev_alloc(4 * probe); ctx.synth_counts(spec, 0, probe, d); ctx.sync()
    pc = np.zeros(probe, np.uint32); ctx.d2h(pc, d); ctx.dev_free(d)
    n = int(records / pc.mean())
d_cnt = ctx.dev_alloc(4 * n); d_off = ctx.dev_alloc(8 * (n + 1))
ctx.synth_counts(spec, 0, n, d_cnt); ctx.offsets_from_counts(d_cnt, n, d_off); ctx.sync()
off = np.zeros(n + 1, np.uint64); ctx.d2h(off, d_off); nrec = int(off[-1])
d_recs = ctx.dev_alloc(40 * nrec + 256); d_fl = ctx.dev_alloc(n); d_ct = ctx.dev_alloc(4 * n)
ctx.synth_fill(spec, 0, n, d_off, d_recs, 0); ctx.sync()
g = ms.geometry_from_dims(p, spec.width, spec.height)
scan, d_in, rb = ctx.scan_device, d_recs, 40
if packed:
    d_r8 = ctx.dev_alloc(8 * nrec + 256); ctx.pack_records_device(d_recs, nrec, d_r8); ctx.sync()
    scan, d_in, rb = ctx.scan_device_packed, d_r8, 8
for _ in range(3): scan(d_in, d_off, None, [g], n, d_fl, d_ct)
ctx.sync(); ctx.reset_stats(); ctx.set_profiling(True)
for _ in range(10): scan(d_in, d_off, None, [g], n, d_fl, d_ct)
ctx.sync(); st = ctx.stats()
ms_ = st.scan_ms / st.scan_launches
fl = np.zeros(n, np.uint8); ctx.d2h(fl, d_fl)
print(json.dumps({"ms": ms_, "gbs": (rb * nrec + 17 * n) / ms_ / 1e6, "grecs": nrec / ms_ / 1e6, "records": nrec, "frames": n, "active": int(fl.sum())}))
"""

WORKLOADS = {"stream": (4, 5, 0), "dense4k": (2, 3, 1200), "cctv": (1, 2, 18000)}


def main():
    records = sys.argv[1] if len(sys.argv) > 1 else "4e8"
    wl = sys.argv[2] if len(sys.argv) > 2 else "stream"
    layout = sys.argv[3] if len(sys.argv) > 3 else "native"  # or "packed" (mscan_mv8 records)
    preset, seed, fixed = WORKLOADS[wl]
    plans = [(0, 0, 0, 0), (3, 2, 8, 1), (3, 3, 8, 1), (3, 4, 8, 1), (4, 2, 8, 1), (4, 3, 8, 1), (4, 4, 8, 1), (5, 2, 8, 1), (5, 3, 8, 1), (6, 2, 8, 1),
             (2, 4, 16, 1), (3, 2, 16, 1), (3, 4, 16, 1), (2, 8, 8, 1)] if layout == "packed" else [(0, 0, 0, 0), (1, 4, 16, 0), (1, 4, 16, 1), (1, 6, 16, 1), (1, 7, 16, 1), (1, 8, 16, 0), (1, 8, 16, 1), (1, 9, 16, 0),
             (1, 9, 16, 1), (1, 10, 16, 1), (2, 3, 8, 0), (2, 4, 8, 0), (2, 4, 8, 1), (2, 5, 8, 1), (3, 2, 8, 0), (3, 3, 8, 1)]
    for ctas, st, warps, c16 in plans:
        env = dict(os.environ)
        if ctas:
            env["MSCAN_KA_CTAS"], env["MSCAN_KA_STAGES"], env["MSCAN_KA_WARPS"] = str(ctas), str(st), str(warps)
            env["MSCAN_KA_CNT16"] = str(c16)
        r = subprocess.run([sys.executable, "-c", f"ROOT={str(ROOT)!r}\n" + CHILD, records, str(preset), str(seed), str(fixed), layout],
                           env=env, capture_output=True, text=True)
        if r.returncode != 0:
            print(f"ctas={ctas} stages={st} warps={warps} cnt16={c16}: FAILED {r.stderr[-300:]}")
            continue
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"{wl}/{layout} ctas={ctas or 'auto'} stages={st or 'auto'} warps={warps or 'auto'} cnt16={c16}: {d['ms']:.3f} ms  {d['gbs']:.0f} GB/s  "
              f"{d['grecs']:.1f} G rec/s  active={d['active']}", flush=True)


if __name__ == "__main__":
    main()
