"""CPU, world_size 2 over gloo: the N>1 host logic — videos dealt to ranks with no data-path collective,
per-rank results gathered on rank 0 equal the single-process results, max-over-ranks timing and summed
work counters aggregate as bench.py reports them. (No GPU here, so each rank's scan uses the oracle —
the thing under test is the sharding/aggregation, not the kernel.)"""
import json
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent

WORKER = r"""
import json, os, sys
sys.path[:0] = [os.path.join(ROOT, "motion-estimated-video-trimmer_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import motionscan as ms
import oracle_lib as orc
from motionscan.dist import Dist, shard_videos, throughput

D = Dist("gloo")
N_VIDEOS, FRAMES = 7, 90
p = ms.shipped_env_params()
mine = shard_videos(N_VIDEOS, D.world, D.rank)
results, n_rec = {}, 0
for v in mine:
    spec = ms.synth_preset(0, 100 + v)
    cnt, off, recs, pts = ms.synth_host(spec, 0, FRAMES, n_threads=2)
    gw, gh, m = orc.geometry(spec.width, spec.height)
    flags, counts = orc.scan_frames(orc.make_cfg(p, gw, gh, m), recs, off)
    segs, res = orc.video_tail(pts, flags, FRAMES / spec.fps, p.max_gap_sec, p.padding_sec, p.min_savings_pct)
    results[v] = [int(res.decision), int(flags.sum()), [[float(a).hex(), float(b).hex()] for a, b in zip(segs["start"], segs["end"])]]
    n_rec += int(off[-1])
D.barrier()
my_ms = 10.0 + 5.0 * D.rank          # pretend timings: rank 1 is the slowest
total = D.allsum(n_rec)
slowest = D.allmax(my_ms)
gathered = D.gather_objects(results)
if D.rank == 0:
    merged = {}
    for part in gathered:
        for k, val in part.items():
            assert k not in merged        # shards are disjoint
            merged[k] = val
    print(json.dumps({"world": D.world, "videos": sorted(int(k) for k in merged), "results": {str(k): v for k, v in merged.items()},
                      "total_records": total, "slowest_ms": slowest, "value": throughput(total, 3, slowest)}))
D.close()
"""


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def run_world(world):
    port = free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", f"ROOT={str(ROOT)!r}\n" + WORKER], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    return json.loads(outs[0][0].strip().splitlines()[-1])


def test_two_ranks_equal_one_rank():
    one = run_world(1)
    two = run_world(2)
    assert two["world"] == 2 and one["world"] == 1
    assert one["videos"] == two["videos"] == list(range(7))        # every video scanned exactly once
    assert one["results"] == two["results"]                          # identical per-video decisions and segments
    assert one["total_records"] == two["total_records"]
    assert two["slowest_ms"] == 15.0 and one["slowest_ms"] == 10.0   # max over ranks, not the mean
    assert two["value"] == two["total_records"] * 3 / 15e-3
    assert any(v[0] == 1 for v in one["results"].values())           # the batch contains real cuts


def test_shard_videos_partition():
    from motionscan.dist import shard_videos

    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 64):
            parts = [shard_videos(n, world, r) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(x) for x in parts) - min(len(x) for x in parts) <= 1


def test_strong_scaling_shares_and_video_pieces():
    """bench.py --scaling strong: one stream cut into contiguous frame shares; the videos a share cuts
    through become pieces, and the pieces of all shares tile every video exactly once."""
    from motionscan.dist import strong_share, video_pieces

    for total in (1, 7, 61275, 100153):
        for world in (1, 2, 3, 4, 8):
            shares = [strong_share(total, world, r) for r in range(world)]
            assert shares[0][0] == 0 and sum(n for _, n in shares) == total
            assert all(shares[r][0] + shares[r][1] == shares[r + 1][0] for r in range(world - 1))
            assert max(n for _, n in shares) - min(n for _, n in shares) <= 1
    starts = list(range(0, 100153 + 18000, 18000))  # 10-minute videos
    covered = []
    for r in range(8):
        f0, n = strong_share(100153, 8, r)
        cuts = video_pieces(starts, f0, n)
        assert cuts[0] == 0 and cuts[-1] == n and cuts == sorted(set(cuts))
        covered += [(f0 + a, f0 + b) for a, b in zip(cuts[:-1], cuts[1:])]
    assert covered[0][0] == 0 and covered[-1][1] == 100153
    assert all(a[1] == b[0] for a, b in zip(covered[:-1], covered[1:]))          # the pieces tile the stream
    assert all(a // 18000 == (b - 1) // 18000 for a, b in covered)                # and none spans two videos
