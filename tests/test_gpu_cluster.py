"""GPU: K-A for block grids that do not fit one CTA's shared memory (8K, 16K, wide strips) — a thread-block cluster
keeps the grid in distributed shared memory by row bands (csrc/ka_scan_cluster.cu). Bit-exact against the oracle for
both record layouts, with clusters that straddle band boundaries, votes that cross CTAs, saturation, the 8-neighbour
extension, degenerate knobs, and mixed with small grids in the same launches. MSCAN_KA_NO_CLUSTER=1 (the older
global-counter path) must agree too."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import kats
import motionscan as ms
import oracle_lib as orc
from test_gpu_parity import np_count_adj
from test_oracle_kats import random_frame

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

SHAPES = {"8k": (7680, 4320), "16k": (15360, 8640), "strip": (30000, 1200), "tall": (1280, 20000)}


def cfg_for(p, w, h):
    gw, gh, m = orc.geometry(w, h, p.block_size, p.block_shift, p.vertical_mask)
    return orc.make_cfg(p, gw, gh, m), gw, gh, m


def scan(p, w, h, frames, mode):
    cnt = np.array([0 if f is None else len(f) for f in frames], dtype=np.uint32)
    recs = kats.cat(*[f for f in frames if f is not None and len(f)])
    with ms.Context(0, p) as ctx:
        ctx.set_staging_mode(ms.STAGING_NATIVE if mode == "native" else ms.STAGING_AUTO)
        ctx.video_open(1, w, h)
        ctx.submit(1, np.arange(len(frames)) / 30.0, cnt, recs if len(recs) else None)
        return ctx.collect(1)


@pytest.mark.parametrize("mode", ["native", "projected"])
@pytest.mark.parametrize("shape", sorted(SHAPES))
def test_random_frames_big_grids(shape, mode):
    w, h = SHAPES[shape]
    rng = np.random.default_rng(hash(shape) % 1000)
    p = kats.env_params(vectors_needed=int(rng.integers(1, 4)), clusters_needed=int(rng.integers(1, 4)))
    frames = []
    for i in range(14):
        if i % 6 == 3:
            frames.append(None)
        else:  # raster-ordered halves (votes stay in their band) and shuffled halves (votes cross CTAs)
            f = random_frame(rng, int(rng.integers(1, 40000)), w, h, int(rng.integers(2, 12)))
            if i % 2:
                f = f[np.lexsort((f["dst_x"], f["dst_y"]))]
            frames.append(f)
    cfg, gw, gh, m = cfg_for(p, w, h)
    flags, counts = scan(p, w, h, frames, mode)
    for i, f in enumerate(frames):
        assert counts[i] == orc.full_count(cfg, f), (shape, mode, i)
        assert flags[i] == orc.check_frame(cfg, f), (shape, mode, i)
    assert counts.max() > 0


@pytest.mark.parametrize("shape", ["8k", "16k"])
def test_clusters_across_band_boundaries(shape):
    """Vertical pairs whose two cells belong to different CTAs of the cluster (every band boundary for 2, 4, 8 and 16
    CTAs), horizontal pairs on the boundary rows, and a diagonal pair across a boundary (not a cluster)."""
    w, h = SHAPES[shape]
    p = kats.env_params()
    cfg, gw, gh, m = cfg_for(p, w, h)
    frames, want = [], []
    for C in (2, 4, 8, 16):
        rpr = -(-gh // C)
        for b in range(1, C):
            y = b * rpr  # first row of band b; y-1 is the last row of band b-1
            if not (m < y < gh - m):
                continue
            frames.append(kats.cat(kats.cell(50, y - 1), kats.cell(50, y)))          # vertical pair across the boundary
            frames.append(kats.cat(kats.cell(31, y - 1), kats.cell(32, y - 1), kats.cell(200, y), kats.cell(201, y)))
            frames.append(kats.cat(kats.cell(70, y - 1), kats.cell(71, y)))          # diagonal across the boundary
            frames.append(kats.cat(kats.cell(5, y - 2), kats.cell(5, y - 1), kats.cell(5, y), kats.cell(5, y + 1)))
    want = [orc.full_count(cfg, f) for f in frames]
    assert 2 in want and 4 in want and 0 in want
    for mode in ("native", "projected"):
        flags, counts = scan(p, w, h, frames, mode)
        assert list(counts) == want, mode
    # with the 8-neighbour extension the diagonal pairs count
    p8 = kats.env_params(adjacency=8)
    flags8, counts8 = scan(p8, w, h, frames, "native")
    for i, f in enumerate(frames):
        assert counts8[i] == np_count_adj(p8, gw, gh, m, f, True), i


def test_saturation_and_degenerate_knobs_8k():
    w, h = SHAPES["8k"]
    p = kats.env_params()
    cfg, gw, gh, m = cfg_for(p, w, h)
    yb = -(-gh // 2)  # band boundary of the 2-CTA cluster
    heavy = kats.cat(kats.cell(100, yb, 120000), kats.cell(101, yb, 3))        # 16-bit halves of one word
    pair = kats.cat(kats.cell(100, yb - 1, 70000), kats.cell(100, yb, 70000))  # heavy cells in two different CTAs
    frames = [heavy, pair, heavy[::-1].copy()]
    want = [orc.full_count(cfg, f) for f in frames]
    assert want == [0, 2, 0]
    for mode in ("native", "projected"):
        flags, counts = scan(p, w, h, frames, mode)
        assert list(counts) == want and list(flags) == [0, 1, 0]
    # VECTORS_NEEDED=0: every cell is active, masked rows count as neighbours
    p0 = kats.env_params(vectors_needed=0, clusters_needed=1)
    cfg0, *_ = cfg_for(p0, w, h)
    one = kats.cell(10, 100)
    flags, counts = scan(p0, w, h, [one, None], "native")
    assert counts[0] == orc.full_count(cfg0, one) == (gw - 2) * (gh - 2 * m) and flags[0] == 1 and counts[1] == 0


def test_mixed_with_small_grids_and_global_fallback_agree():
    """8K, 4K and 1080p videos interleaved in the same launches under the cluster plan; the same run with
    MSCAN_KA_NO_CLUSTER=1 (global-memory counters) gives the same log."""
    code = r"""
import sys, json
sys.path[:0] = [%r, %r]
import numpy as np, kats, motionscan as ms
from test_oracle_kats import random_frame
rng = np.random.default_rng(77)
p = kats.env_params(vectors_needed=2)
shapes = {1: (7680, 4320), 2: (1920, 1080), 3: (3840, 2160)}
frames = {v: [random_frame(rng, int(rng.integers(1, 30000)), w, h, 6) if i %% 4 else None for i in range(16)] for v, (w, h) in shapes.items()}
with ms.Context(0, p) as ctx:
    for v, (w, h) in shapes.items():
        ctx.video_open(v, w, h)
    for a in range(0, 16, 4):
        for v in shapes:
            fr = frames[v][a:a + 4]
            cnt = np.array([0 if f is None else len(f) for f in fr], np.uint32)
            ctx.submit(v, np.arange(a, a + 4) / 30.0, cnt, kats.cat(*[f for f in fr if f is not None]))
    print(json.dumps({str(v): [ctx.collect(v)[1].tolist()] for v in shapes}))
""" % (str(ROOT / "motion-estimated-video-trimmer_b200"), str(ROOT / "tests"))
    outs = []
    # cluster plan / global-counter plan / cluster plan whose launch is refused (falls back at launch time)
    for extra in ({"MSCAN_KA_NO_CLUSTER": "0"}, {"MSCAN_KA_NO_CLUSTER": "1"}, {"MSCAN_KA_FAIL_CLUSTER_LAUNCH": "1"}):
        env = dict(os.environ, **extra)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip().splitlines()[-1])
    assert outs[0] == outs[1] == outs[2]
    # and against the oracle (same seed, regenerated here)
    import json

    got = json.loads(outs[0])
    rng = np.random.default_rng(77)
    p = kats.env_params(vectors_needed=2)
    shapes = {1: (7680, 4320), 2: (1920, 1080), 3: (3840, 2160)}
    frames = {v: [random_frame(rng, int(rng.integers(1, 30000)), w, h, 6) if i % 4 else None for i in range(16)] for v, (w, h) in shapes.items()}
    for v, (w, h) in shapes.items():
        cfg, *_ = cfg_for(p, w, h)
        assert got[str(v)][0] == [orc.full_count(cfg, f) for f in frames[v]], v


@pytest.mark.parametrize("big,small", [("16k", (320, 240)), ("8k", (320, 64)), ("16k", (1920, 64)), ("strip", (640, 480))])
@pytest.mark.parametrize("mode", ["native", "projected"])
def test_tiny_grid_shares_a_context_with_a_cluster_sized_one(big, small, mode):
    """A video whose grid has no more rows than the cluster has CTAs (one row per band, or bands without rows) in the
    same context — and the same launches — as a grid that needs the cluster: every vote must reach the band that owns
    its row, and the frames after the small video's must not see leftovers (the cluster size is chosen per context
    from the largest geometry, so the small grid runs with rows-per-band == 1)."""
    bw, bh = SHAPES[big]
    sw, sh = small
    rng = np.random.default_rng(hash((big, small)) % 100000)
    p = kats.env_params(vectors_needed=2, clusters_needed=1, vertical_mask=0.26)  # margin >= 1 also for 4-row grids
    cfg_b, *_ = cfg_for(p, bw, bh)
    cfg_s, gws, ghs, ms_ = cfg_for(p, sw, sh)
    assert ms_ >= 1
    small_frames, big_frames = [], []
    for i in range(10):
        f = random_frame(rng, int(rng.integers(50, 3000)), sw, sh, int(rng.integers(1, 4)))
        # make sure live rows beyond row 0 get clusters: stack heavy neighbouring cells in every live row
        rows = range(ms_, ghs - ms_)
        f = kats.cat(f, *[kats.cell(1 + (i % max(gws - 3, 1)), y, 6) for y in rows], *[kats.cell(2 + (i % max(gws - 3, 1)), y, 6) for y in rows])
        small_frames.append(f)
        big_frames.append(random_frame(rng, int(rng.integers(1, 20000)), bw, bh, int(rng.integers(2, 8))))
    with ms.Context(0, p) as ctx:
        ctx.set_staging_mode(ms.STAGING_NATIVE if mode == "native" else ms.STAGING_AUTO)
        ctx.video_open(1, bw, bh)
        ctx.video_open(2, sw, sh)
        for i in range(10):  # interleaved: small frame, then a big one in the same segment
            ctx.submit(2, np.array([i / 30.0]), np.array([len(small_frames[i])], np.uint32), small_frames[i])
            ctx.submit(1, np.array([i / 30.0]), np.array([len(big_frames[i])], np.uint32), big_frames[i])
        fs, cs = ctx.collect(2)
        fb, cb = ctx.collect(1)
    want_s = [orc.full_count(cfg_s, f) for f in small_frames]
    want_b = [orc.full_count(cfg_b, f) for f in big_frames]
    assert max(want_s) > 0
    assert list(cs) == want_s, (big, small, mode)
    assert list(cb) == want_b, (big, small, mode)
    assert list(fs) == [orc.check_frame(cfg_s, f) for f in small_frames]


def test_remote_votes_saturate_without_carry_16k():
    """Many votes into ONE cell owned by another CTA of the cluster (red.shared::cluster with the carry guard), from a
    frame whose records are shuffled so that every CTA's slice votes remotely: the 16-bit half must not carry into its
    word-mate — the neighbouring cell stays inactive."""
    w, h = SHAPES["16k"]
    p = kats.env_params(vectors_needed=200, clusters_needed=1)
    cfg, gw, gh, m = cfg_for(p, w, h)
    gx, gy = 200, gh - m - 2                      # a cell of the last bands
    gx -= gx & 1                                  # even x: its word-mate is (gx+1, gy)
    heavy = kats.cell(gx, gy, 150000)
    mate_light = kats.cell(gx + 1, gy, 150)       # below VECTORS_NEEDED: must stay inactive whatever the left half does
    other = kats.cat(kats.cell(50, m + 1, 300), kats.cell(51, m + 1, 300))  # one real cluster pair elsewhere
    rng = np.random.default_rng(5)
    f = kats.cat(heavy, mate_light, other)
    f = f[rng.permutation(len(f))]
    want = orc.full_count(cfg, f)
    assert want == 2
    for mode in ("native", "projected"):
        flags, counts = scan(p, w, h, [f, kats.cell(gx + 1, gy, 150), f], mode)
        assert list(counts) == [want, 0, want], mode
