"""GPU: the CUDA path, called through the C ABI, against the CPU oracle — bit-exact flags, full
cluster counts, segment doubles and decisions (integer/byte work: no tolerance anywhere)."""
import ctypes as C

import numpy as np
import pytest

import kats
import motionscan as ms
import oracle_lib as orc

pytestmark = pytest.mark.gpu


def cfg_for(p, w, h):
    gw, gh, m = orc.geometry(w, h, p.block_size, p.block_shift, p.vertical_mask)
    return orc.make_cfg(p, gw, gh, m)


def seg_bytes(a):
    return np.ascontiguousarray(a).tobytes()


def oracle_tail(p, pts, flags, duration):
    return orc.video_tail(pts, flags, duration, p.max_gap_sec, p.padding_sec, p.min_savings_pct)


def assert_result_equal(res, ores):
    assert res.decision == ores.decision
    assert res.n_motion_frames == ores.n_motion_frames
    assert res.n_segments == ores.n_segments
    for f in ("out_dur", "time_removed", "saved_pct"):
        a, b = np.float64(getattr(res, f)), np.float64(getattr(ores, f))
        assert a.tobytes() == b.tobytes(), f


def run_frames(ctx, vid, w, h, frames, pts=None):
    """frames: list of record arrays (None = no side data). Returns (flags, counts)."""
    cnt = np.array([0 if f is None else len(f) for f in frames], dtype=np.uint32)
    recs = kats.cat(*[f for f in frames if f is not None and len(f)])
    if pts is None:
        pts = np.arange(len(frames), dtype=np.float64) / 30.0
    ctx.video_open(vid, w, h)
    ctx.submit(vid, pts, cnt, recs if len(recs) else None)
    return ctx.collect(vid)


# ------------------------------------------------------------------------------------ K-A --------
def test_frame_kats_gpu():
    K = kats.frame_kats()
    # group by params so one context handles many KAT frames in one launch
    groups = {}
    for name, (p, recs, flag, count) in K.items():
        key = (p.mv_threshold_sq, p.vectors_needed, p.clusters_needed)
        groups.setdefault(key, (p, []))[1].append((name, recs, flag, count))
    for p, items in groups.values():
        with ms.Context(0, p) as ctx:
            flags, counts = run_frames(ctx, 1, kats.W, kats.H, [r for _, r, _, _ in items])
            for i, (name, _, flag, count) in enumerate(items):
                assert flags[i] == flag, name
                assert counts[i] == count, name


@pytest.mark.parametrize("seed", range(8))
def test_random_frames_vs_oracle(seed):
    from test_oracle_kats import random_frame

    rng = np.random.default_rng(1000 + seed)
    w, h = [(1920, 1080), (3840, 2160), (1280, 720), (352, 288)][seed % 4]
    p = kats.env_params() if seed % 2 else kats.code_defaults()
    p.vectors_needed = int(rng.integers(0, 6))
    p.clusters_needed = int(rng.integers(0, 5))
    p.mv_threshold_sq = float(rng.choice([0.0, 1.0, 4.0, 4.5, 16.0, 25.0]))
    frames = []
    for i in range(40):
        n = int(rng.integers(0, 6000)) if i % 7 else 0  # ragged, with empty frames in between
        frames.append(random_frame(rng, n, w, h, int(rng.integers(1, 6))) if n else None)
    cfg = cfg_for(p, w, h)
    with ms.Context(0, p) as ctx:
        flags, counts = run_frames(ctx, 5, w, h, frames)
    for i, f in enumerate(frames):
        want = orc.full_count(cfg, f)
        assert counts[i] == want, i
        assert flags[i] == orc.check_frame(cfg, f), i


@pytest.mark.parametrize("config,n_frames", [(0, 1800), (3, 600), (2, 24)])
def test_synthetic_clip_vs_oracle(config, n_frames):
    """BASELINE configs as synthetic MV streams: 60 s 1080p30 (config 0, the reference's CPU case),
    a batch clip and the 4K dense field; flags/counts/segments/decision bit-exact."""
    spec = ms.synth_preset(config, 1 + config)
    cnt, off, recs, pts = ms.synth_host(spec, 0, n_frames)
    p = kats.env_params()
    cfg = cfg_for(p, spec.width, spec.height)
    oflags, ocounts = orc.scan_frames(cfg, recs, off, threads=8)
    duration = n_frames / spec.fps
    with ms.Context(0, p) as ctx:
        ctx.video_open(9, spec.width, spec.height)
        ctx.submit(9, pts, cnt, recs)
        flags, counts = ctx.collect(9)
        assert np.array_equal(counts, ocounts)
        assert np.array_equal(flags, oflags)
        segs, res = ctx.motion_segments(9, duration)
        jsegs, jres = ctx.segments(9, duration)
    osegs, ores = oracle_tail(p, pts, oflags, duration)
    assert_result_equal(res, ores)
    assert seg_bytes(segs) == seg_bytes(osegs)
    if ores.decision == ms.CUT:
        assert seg_bytes(jsegs) == seg_bytes(osegs)
    elif ores.decision == ms.FULL_COPY:
        assert [(s["start"], s["end"]) for s in jsegs] == [(0.0, duration)]
    else:
        assert len(jsegs) == 0
    if config == 0:  # forced static spans [300,900) and [1200,1500) ⇒ >= 2 segments, a CUT
        assert oflags.sum() > 0 and not oflags[300:900].any() and not oflags[1200:1500].any()
        assert ores.n_segments >= 2 and ores.decision == ms.CUT


def test_submit_chunked_pinned_and_pageable_agree():
    """Frames fed in many small submits (pageable) and in one pinned submit give identical logs; tiny
    slabs force frames to straddle several K-A launches."""
    spec = ms.synth_preset(3, 77)
    n = 300
    cnt, off, recs, pts = ms.synth_host(spec, 0, n)
    p = kats.env_params()
    with ms.Context(0, p, 0, 8 << 20) as ctx:  # 8 MiB slabs ⇒ ~18 frames per launch
        ctx.video_open(1, spec.width, spec.height)
        for a in range(0, n, 37):
            b = min(n, a + 37)
            ctx.submit(1, pts[a:b], cnt[a:b], recs[int(off[a]) : int(off[b])])
        f1, c1 = ctx.collect(1)
        # pinned, one call
        hp = ctx.pinned_array(len(recs), ms.MV_DTYPE)
        hp[:] = recs
        ctx.video_open(2, spec.width, spec.height)
        ctx.submit(2, pts, cnt, hp)
        f2, c2 = ctx.collect(2)
        st = ctx.stats()
        ctx.host_free(hp.ctypes.data)
    cfg = cfg_for(p, spec.width, spec.height)
    of, oc = orc.scan_frames(cfg, recs, off, threads=8)
    assert np.array_equal(c1, oc) and np.array_equal(f1, of)
    assert np.array_equal(c2, oc) and np.array_equal(f2, of)
    assert st.scan_launches > 10 and st.frames_scanned == 2 * n


def test_mixed_resolutions_interleaved():
    """Frames of a 1080p, a 4K and a 720p video interleaved in the same launches (per-frame geometry)."""
    p = kats.env_params()
    specs = {1: ms.synth_preset(0, 5), 2: ms.synth_preset(2, 6), 3: ms.synth_preset(0, 7)}
    specs[3].width, specs[3].height = 1280, 720
    data = {v: ms.synth_host(s, 0, 45) for v, s in specs.items()}
    with ms.Context(0, p) as ctx:
        for v, s in specs.items():
            ctx.video_open(v, s.width, s.height)
        for a in range(0, 45, 5):
            for v in (1, 2, 3):
                cnt, off, recs, pts = data[v]
                ctx.submit(v, pts[a : a + 5], cnt[a : a + 5], recs[int(off[a]) : int(off[a + 5])])
        out = {v: ctx.collect(v) for v in specs}
        segs, soff, res = ctx.segments_batch([1, 2, 3], [1.5, 1.5, 1.5])
    for i, (v, s) in enumerate(specs.items()):
        cnt, off, recs, pts = data[v]
        cfg = cfg_for(p, s.width, s.height)
        of, oc = orc.scan_frames(cfg, recs, off, threads=4)
        assert np.array_equal(out[v][1], oc), v
        assert np.array_equal(out[v][0], of), v
        osegs, ores = oracle_tail(p, pts, of, 1.5)
        assert res["decision"][i] == ores.decision
        assert np.float64(res["saved_pct"][i]).tobytes() == np.float64(ores.saved_pct).tobytes()


# ------------------------------------------------------------------------------------ K-C --------
def feed_flags(ctx, vid, pts, flags):
    """Drive K-C with chosen flags: frames whose records form one horizontal cluster pair, or nothing."""
    active = kats.cat(kats.cell(10, 10), kats.cell(11, 10))
    frames = [active if f else None for f in flags]
    got, _ = run_frames(ctx, vid, kats.W, kats.H, frames, np.asarray(pts, dtype=np.float64))
    assert np.array_equal(got, np.asarray(flags, dtype=np.uint8))


@pytest.mark.parametrize("name", sorted(kats.segment_kats()))
def test_segment_kats_gpu(name):
    ts, duration, segs, out_dur, removed, decision = kats.segment_kats()[name]
    p = kats.env_params()
    with ms.Context(0, p) as ctx:
        if ts:
            feed_flags(ctx, 3, ts, [1] * len(ts))
        else:
            feed_flags(ctx, 3, [0.0, 1.0], [0, 0])
            ts = [0.0, 1.0]
            flags0 = np.zeros(2, np.uint8)
        got, res = ctx.motion_segments(3, duration)
        job, _ = ctx.segments(3, duration)
    flags = np.ones(len(ts), np.uint8) if decision != ms.NO_MOTION else flags0
    osegs, ores = oracle_tail(p, np.array(ts), flags, duration)
    assert res.decision == decision
    assert_result_equal(res, ores)
    assert seg_bytes(got) == seg_bytes(osegs)
    assert len(got) == len(segs)
    if decision == ms.FULL_COPY:
        assert [(s["start"], s["end"]) for s in job] == [(0.0, duration)]
    if decision == ms.NO_MOTION:
        assert len(job) == 0 and len(got) == 0


@pytest.mark.parametrize("seed", range(6))
def test_random_tails_vs_oracle(seed):
    rng = np.random.default_rng(500 + seed)
    duration = float(rng.choice([60.0, 600.0, 37.25]))
    fps = float(rng.choice([30.0, 25.0, 29.97]))
    n = int(rng.integers(1, 5000))
    frames = np.sort(rng.choice(int(duration * fps), size=min(n, int(duration * fps)), replace=False))
    pts = frames / fps
    if seed % 2 == 0:  # unsorted with duplicates: chunk workers finish in any order (pipeline.cpp:302-304)
        pts = np.concatenate([pts, pts[: len(pts) // 3]])
        rng.shuffle(pts)
    flags = (rng.random(len(pts)) < 0.6).astype(np.uint8)
    p = kats.env_params(max_gap_sec=float(rng.choice([5.0, 1.0, 0.0])), padding_sec=float(rng.choice([0.5, 2.0, 0.0])))
    with ms.Context(0, p) as ctx:
        feed_flags(ctx, 4, pts, flags)
        got, res = ctx.motion_segments(4, duration)
    osegs, ores = oracle_tail(p, pts, flags, duration)
    assert_result_equal(res, ores)
    assert seg_bytes(got) == seg_bytes(osegs)


# --------------------------------------------------------------------------- device-resident -----
def test_device_generator_matches_host_and_scan_device():
    """The on-device generator writes the same bytes as the host one; K-A / K-C on caller-owned device
    memory (the decode-free stream path of configs[4]) agree with the oracle."""
    spec = ms.synth_preset(4, 5)
    spec.frames_per_video = 120
    n = 360
    cnt, off, recs, pts = ms.synth_host(spec, 1000, n)
    p = kats.env_params()
    with ms.Context(0, p) as ctx:
        d_cnt = ctx.dev_alloc(4 * n)
        d_off = ctx.dev_alloc(8 * (n + 1))
        ctx.synth_counts(spec, 1000, n, d_cnt)
        ctx.offsets_from_counts(d_cnt, n, d_off)
        g_cnt = np.zeros(n, np.uint32)
        g_off = np.zeros(n + 1, np.uint64)
        ctx.sync()
        ctx.d2h(g_cnt, d_cnt)
        ctx.d2h(g_off, d_off)
        assert np.array_equal(g_cnt, cnt) and np.array_equal(g_off, off)
        d_recs = ctx.dev_alloc(40 * int(off[-1]) + 16)
        d_pts = ctx.dev_alloc(8 * n)
        ctx.synth_fill(spec, 1000, n, d_off, d_recs, d_pts)
        ctx.sync()
        g_recs = np.zeros(int(off[-1]), ms.MV_DTYPE)
        g_pts = np.zeros(n, np.float64)
        ctx.d2h(g_recs, d_recs)
        ctx.d2h(g_pts, d_pts)
        assert g_recs.tobytes() == recs.tobytes()
        assert g_pts.tobytes() == pts.tobytes()
        d_flags = ctx.dev_alloc(n)
        d_counts = ctx.dev_alloc(4 * n)
        geom = ms.geometry_from_dims(p, spec.width, spec.height)
        ctx.scan_device(d_recs, d_off, None, [geom], n, d_flags, d_counts)
        voff = np.array([0, 120, 240, 360], np.uint64)
        durs = np.array([4.0, 4.0, 4.0])
        d_segs = ctx.dev_alloc(16 * n)
        d_res = ctx.dev_alloc(40 * 3)
        ctx.segments_device(voff, durs, d_pts, d_flags, d_segs, d_res)
        ctx.sync()
        flags = np.zeros(n, np.uint8)
        counts = np.zeros(n, np.uint32)
        segs = np.zeros(n, ms.SEG_DTYPE)
        res = np.zeros(3, ms.RESULT_DTYPE)
        ctx.d2h(flags, d_flags)
        ctx.d2h(counts, d_counts)
        ctx.d2h(segs, d_segs)
        ctx.d2h(res, d_res)
        for d in (d_cnt, d_off, d_recs, d_pts, d_flags, d_counts, d_segs, d_res):
            ctx.dev_free(d)
    cfg = cfg_for(p, spec.width, spec.height)
    of, oc = orc.scan_frames(cfg, recs, off, threads=8)
    assert np.array_equal(counts, oc) and np.array_equal(flags, of)
    for v in range(3):
        a, b = int(voff[v]), int(voff[v + 1])
        osegs, ores = oracle_tail(p, pts[a:b], of[a:b], 4.0)
        assert res["decision"][v] == ores.decision
        assert res["n_segments"][v] == ores.n_segments
        assert np.float64(res["out_dur"][v]).tobytes() == np.float64(ores.out_dur).tobytes()
        assert seg_bytes(segs[a : a + ores.n_segments]) == seg_bytes(osegs)


def test_errors_are_loud():
    p = kats.env_params()
    with ms.Context(0, p) as ctx:
        with pytest.raises(ms.MscanError) as e:
            ctx.submit(99, np.zeros(1), np.zeros(1, np.uint32), None)  # video not open
        assert e.value.code == ms.ERR_INVALID
        ctx.video_open(1, 1920, 1080)
        with pytest.raises(ms.MscanError):
            ctx.video_open(1, 1920, 1080)  # already open
        g = ms.Geometry(32767, 32767, 10, 0)  # 10^9 cells: beyond even the global-memory counter scratch
        with pytest.raises(ms.MscanError) as e:
            ctx.video_open_geometry(2, g)
        assert e.value.code == ms.ERR_UNSUPPORTED


def test_large_grid_uses_global_counters():
    """8K video (480x270 = 129 600 cells) does not fit shared memory: the counters move to a global
    scratch, results stay bit-exact; a 1080p video interleaved in the same launches is unaffected."""
    from test_oracle_kats import random_frame

    rng = np.random.default_rng(88)
    p = kats.env_params(vectors_needed=3)
    shapes = {1: (7680, 4320), 2: (1920, 1080)}
    frames = {v: [random_frame(rng, int(rng.integers(1, 20000)), w, h, int(rng.integers(1, 8))) if i % 5 else None
                  for i in range(30)] for v, (w, h) in shapes.items()}
    with ms.Context(0, p) as ctx:
        for v, (w, h) in shapes.items():
            ctx.video_open(v, w, h)
        for a in range(0, 30, 6):
            for v in shapes:
                fr = frames[v][a : a + 6]
                cnt = np.array([0 if f is None else len(f) for f in fr], np.uint32)
                ctx.submit(v, np.arange(a, a + 6) / 30.0, cnt, kats.cat(*[f for f in fr if f is not None]))
        out = {v: ctx.collect(v) for v in shapes}
    for v, (w, h) in shapes.items():
        cfg = cfg_for(p, w, h)
        for i, f in enumerate(frames[v]):
            assert out[v][1][i] == orc.full_count(cfg, f), (v, i)
            assert out[v][0][i] == orc.check_frame(cfg, f), (v, i)
    assert out[1][0].any()


def test_4k_u16_counters_saturation_guard():
    """4K grids use 16-bit shared-memory counters (two per word). 120 000 votes into one cell must not
    carry into its word-mate: the neighbour cell (same 32-bit word) stays below VECTORS_NEEDED."""
    p = kats.env_params()
    heavy = kats.cat(kats.cell(100, 50, 120000), kats.cell(101, 50, 3))   # cells 100/101 of a row share a word
    pair = kats.cat(kats.cell(100, 50, 70000), kats.cell(101, 50, 70000))
    lone = kats.cat(kats.cell(101, 60, 90000))                           # odd cell alone: even mate must stay 0
    cfg = cfg_for(p, 3840, 2160)
    with ms.Context(0, p) as ctx:
        flags, counts = run_frames(ctx, 1, 3840, 2160, [heavy, pair, lone, heavy])
    want = [orc.full_count(cfg, f) for f in (heavy, pair, lone, heavy)]
    assert want == [0, 2, 0, 0]
    assert list(counts) == want and list(flags) == [0, 1, 0, 0]


def np_count_adj(p, gw, gh, margin, recs, adj8):
    """numpy restatement with selectable connectivity (8 = extension, not in the reference)."""
    tx, ty = recs["dst_x"].astype(np.int64), recs["dst_y"].astype(np.int64)
    mag = (tx - recs["src_x"].astype(np.int64)) ** 2 + (ty - recs["src_y"].astype(np.int64)) ** 2
    gx, gy = tx >> p.block_shift, ty >> p.block_shift
    keep = ~(mag.astype(np.float64) < p.mv_threshold_sq) & (gx >= 0) & (gx < gw) & (gy >= margin) & (gy < gh - margin)
    grid = np.zeros((gh, gw), np.int64)
    np.add.at(grid, (gy[keep], gx[keep]), 1)
    act = grid >= (p.vectors_needed & 0xFF)
    pad = np.zeros((gh + 2, gw + 2), bool)
    pad[1:-1, 1:-1] = act
    nb = pad[1:-1, :-2] | pad[1:-1, 2:] | pad[:-2, 1:-1] | pad[2:, 1:-1]
    if adj8:
        nb |= pad[:-2, :-2] | pad[:-2, 2:] | pad[2:, :-2] | pad[2:, 2:]
    cl = act & nb
    cl[:, 0] = cl[:, gw - 1] = False
    cl[:margin, :] = False
    cl[gh - margin :, :] = False
    return int(cl.sum())


def test_adjacency8_extension():
    """CLUSTER_ADJACENCY=8 (extension behind a knob, default 4 = reference): K5's diagonal pair becomes a
    cluster; random frames agree with a numpy restatement; adjacency 4 is unchanged."""
    from test_oracle_kats import random_frame

    p8 = kats.env_params(adjacency=8)
    diag = kats.cat(kats.cell(10, 10), kats.cell(11, 11))
    # word-boundary diagonals: cells (31,20)/(32,21) and (64,30)/(63,31) straddle 32-bit bit-row words
    edge = kats.cat(kats.cell(31, 20), kats.cell(32, 21), kats.cell(64, 30), kats.cell(63, 31))
    rng = np.random.default_rng(5)
    rnd = [random_frame(rng, 4000, kats.W, kats.H, 6) for _ in range(6)]
    with ms.Context(0, p8) as ctx:
        f8, c8 = run_frames(ctx, 1, kats.W, kats.H, [diag, edge] + rnd)
    with ms.Context(0, kats.env_params()) as ctx:
        f4, c4 = run_frames(ctx, 1, kats.W, kats.H, [diag, edge] + rnd)
    assert (f4[0], c4[0]) == (0, 0) and (f8[0], c8[0]) == (1, 2)
    assert c4[1] == 0 and c8[1] == 4
    for i, r in enumerate(rnd):
        assert c8[2 + i] == np_count_adj(p8, 120, 68, 3, r, True)
        assert c4[2 + i] == np_count_adj(p8, 120, 68, 3, r, False)
