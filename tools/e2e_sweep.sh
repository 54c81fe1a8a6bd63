#!/bin/bash
# Host-fed sweep (one GPU): staging slab size x projection store kind x producer threads x frames per submit.
# Each line: the producers / packed_pinned e2e values of one bench.py run (device-resident part cut short).
out=${1:-gpurun_out/r03_e2e_sweep.log}
: > "$out"
run() {
  local tag="$1"; shift
  local line
  line=$(env "$@" python bench.py --steps 3 --warmup 3 --records 1e8 --no-cpu --no-packed --no-spec-stream --e2e-steps 8 $EXTRA 2>/dev/null | tail -1)
  python - "$tag" "$line" >> "$out" <<'PY'
import json, sys
tag, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    m = d["e2e"]["modes"]
    s = " ".join(f"{k}={v['value']/1e9:.2f}" + (f"(h2d {v['h2d_gbs']:.1f} GB/s, standin {v.get('standin_share_of_thread_time', 0):.2f}, hot-only {v.get('hot_path_only_value', 0)/1e9:.1f})" if k == "producers" else "") for k, v in m.items())
    print(f"{tag}: {s} ok={d['e2e']['matches_device_resident']}")
except Exception as e:
    print(f"{tag}: FAILED {e} {line[:200]}")
PY
  tail -1 "$out"
}
for stores in nt plain; do
  for slab in 2 4 8 16 64; do
    EXTRA="--slab-mb $slab --e2e-modes producers,producers_elided,packed_pinned" run "stores=$stores slab=${slab}MiB" MSCAN_PROJECT_STORES=$stores
  done
done
for thr in 4 8 12; do
  EXTRA="--slab-mb 8 --e2e-modes producers --feed-threads $thr" run "stores=plain slab=8MiB threads=$thr" MSCAN_PROJECT_STORES=plain
  EXTRA="--slab-mb 64 --e2e-modes producers --feed-threads $thr" run "stores=nt slab=64MiB threads=$thr" MSCAN_PROJECT_STORES=nt
done
for fb in 4 16; do
  EXTRA="--slab-mb 8 --e2e-modes producers --feed-batch $fb" run "stores=plain slab=8MiB batch=$fb" MSCAN_PROJECT_STORES=plain
done
for win in 256 1024 4096; do
  EXTRA="--slab-mb 8 --e2e-modes producers" run "stores=plain slab=8MiB window=${win}KiB" MSCAN_PROJECT_STORES=plain MSCAN_COPY_WINDOW_KB=$win
done
