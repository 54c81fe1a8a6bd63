#!/usr/bin/env python
"""Host microbenchmark (no GPU): rates of the two host passes of the host-fed path on cache-hot frames —
the projection AVMotionVector → mscan_mv8 (mscan_pack_records, csrc/host_project.cpp) and the decode stand-in's
record writer (mscan_feed_expand, csrc/feed_harness.cpp) — per thread count, scalar vs AVX-512 (re-executes itself
with MSCAN_NO_AVX512 / MSCAN_FEED_NO_AVX512 / MSCAN_PROJECT_STORES)."""
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "motion-estimated-video-trimmer_b200"))


def child():
    import numpy as np

    import motionscan as ms

    spec = ms.synth_preset(4, 5)
    cnt, off, recs, pts = ms.synth_host(spec, 1, 1)
    n = len(recs)
    L, F = ms.lib(), ms.feed_lib()
    out = []
    for T in (1, len(os.sched_getaffinity(0))):
        for what in ("pack", "expand", "expand+pack"):
            iters = 3000

            def work():
                r = recs.copy()
                r8 = np.zeros(n, ms.MV8_DTYPE)
                L.mscan_pack_records(r.ctypes.data, n, r8.ctypes.data)
                for _ in range(iters):
                    if what != "pack":
                        F.mscan_feed_expand(r8.ctypes.data, n, r.ctypes.data)
                    if what != "expand":
                        L.mscan_pack_records(r.ctypes.data, n, r8.ctypes.data)

            th = [threading.Thread(target=work) for _ in range(T)]
            t0 = time.perf_counter()
            for t in th:
                t.start()
            for t in th:
                t.join()
            dt = time.perf_counter() - t0
            out.append(f"{what} T={T}: {T * iters * n / dt / 1e9:.2f} G rec/s")
    print(f"avx512={'off' if os.environ.get('MSCAN_NO_AVX512') == '1' else 'on'} stores={os.environ.get('MSCAN_PROJECT_STORES', 'nt')}: " + " | ".join(out), flush=True)


if __name__ == "__main__":
    if os.environ.get("EXP_HOSTPROJ_CHILD"):
        child()
    else:
        for avx in ("0", "1"):
            for stores in ("nt", "plain"):
                env = dict(os.environ, EXP_HOSTPROJ_CHILD="1", MSCAN_NO_AVX512=avx, MSCAN_FEED_NO_AVX512=avx, MSCAN_PROJECT_STORES=stores)
                subprocess.run([sys.executable, __file__], env=env, check=True)
