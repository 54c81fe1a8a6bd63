"""CPU: the oracle against the known-answer vectors of SURVEY.md §4 and against an independent numpy
restatement. The reference has no tests of its own; see also test_ref_golden.py (reference-built
fixtures)."""
import numpy as np
import pytest

import kats
import motionscan as ms
import oracle_lib as orc

GW, GH, MARGIN = 120, 68, 3


def cfg_for(p, w=kats.W, h=kats.H):
    gw, gh, m = orc.geometry(w, h, p.block_size, p.block_shift, p.vertical_mask)
    return orc.make_cfg(p, gw, gh, m)


def test_geometry_kats():
    # 1080p 120x68 margin 3; 4K 240x135 margin 6; 720p 80x45 margin 2
    assert orc.geometry(1920, 1080) == (120, 68, 3)
    assert orc.geometry(3840, 2160) == (240, 135, 6)
    assert orc.geometry(1280, 720) == (80, 45, 2)
    # product-side host derivation must agree bit for bit (no GPU needed)
    p = ms.default_params()
    for w, h in [(1920, 1080), (3840, 2160), (1280, 720), (640, 360), (7680, 4320), (17, 33), (16, 16)]:
        g = ms.geometry_from_dims(p, w, h)
        assert (g.grid_w, g.grid_h, g.vertical_margin) == orc.geometry(w, h)
    for mask in [0.0, 0.05, 0.1, 0.123, 0.25, 0.49]:
        p.vertical_mask = mask
        for h in range(16, 2200, 97):
            g = ms.geometry_from_dims(p, 1920, h)
            assert (g.grid_w, g.grid_h, g.vertical_margin) == orc.geometry(1920, h, 16, 4, mask)


@pytest.mark.parametrize("name", sorted(kats.frame_kats()))
def test_frame_kat(name):
    p, recs, flag, count = kats.frame_kats()[name]
    cfg = cfg_for(p)
    assert (cfg.grid_w, cfg.grid_h, cfg.vertical_margin) == (GW, GH, MARGIN)
    assert orc.check_frame(cfg, recs) == flag, name
    assert orc.full_count(cfg, recs) == count, name
    # closed form: flag == full_count >= max(1, CLUSTERS_NEEDED)
    assert flag == int(count >= max(1, p.clusters_needed))


@pytest.mark.parametrize("name", sorted(kats.segment_kats()))
def test_segment_kat(name):
    ts, duration, segs, out_dur, removed, decision = kats.segment_kats()[name]
    p = kats.env_params()
    pts = np.array(ts, dtype=np.float64)
    got, res = orc.video_tail(pts, np.ones(len(ts), np.uint8), duration, p.max_gap_sec, p.padding_sec, p.min_savings_pct)
    assert res.decision == decision
    assert len(got) == len(segs)
    for g, (a, b) in zip(got, segs):
        assert g["start"] == pytest.approx(a, abs=1e-12) and g["end"] == pytest.approx(b, abs=1e-12)
    if out_dur is not None:
        assert res.out_dur == pytest.approx(out_dur, abs=1e-12)
        assert res.time_removed == pytest.approx(removed, abs=1e-12)
    if name == "S3":
        assert res.out_dur == 60.0 - (59.8 - 0.5)  # 0.70000000000000284 exactly
    if name == "S4":  # a gap of exactly 5.0 does not split, 5.000000001 does
        assert got["end"][0] == 11.5 and got["start"][1] == 16.000000001 - 0.5
    if name == "S6":
        assert 4.2 < res.saved_pct < 4.25
    if name == "S8":
        assert res.n_motion_frames == 2  # duplicate removed by unique


def np_full_count(p, gw, gh, margin, recs):
    """Independent numpy restatement of Appendix A (no shared code with the C oracle)."""
    if recs is None:
        return 0
    sx, sy = recs["src_x"].astype(np.int64), recs["src_y"].astype(np.int64)
    tx, ty = recs["dst_x"].astype(np.int64), recs["dst_y"].astype(np.int64)
    mag = (tx - sx) ** 2 + (ty - sy) ** 2
    keep = ~(mag.astype(np.float64) < p.mv_threshold_sq)
    gx, gy = tx >> p.block_shift, ty >> p.block_shift
    keep &= (gx >= 0) & (gx < gw) & (gy >= margin) & (gy < gh - margin)
    grid = np.zeros((gh, gw), dtype=np.int64)
    np.add.at(grid, (gy[keep], gx[keep]), 1)
    act = np.minimum(grid, 255) >= (p.vectors_needed & 0xFF)
    pad = np.zeros((gh + 2, gw + 2), dtype=bool)
    pad[1:-1, 1:-1] = act
    nb = pad[1:-1, :-2] | pad[1:-1, 2:] | pad[:-2, 1:-1] | pad[2:, 1:-1]
    cl = act & nb
    cl[:, 0] = False
    cl[:, gw - 1] = False
    cl[:margin, :] = False
    cl[gh - margin :, :] = False
    return int(cl.sum())


def random_frame(rng, n, w, h, hot):
    r = np.zeros(n, dtype=ms.MV_DTYPE)
    # cluster dst around a few hot spots so cells reach VECTORS_NEEDED, plus uniform + out-of-frame
    centres = rng.integers(0, [w, h], size=(hot, 2))
    pick = rng.integers(0, hot, size=n)
    jitter = rng.integers(-24, 25, size=(n, 2))
    dst = centres[pick] + jitter
    uni = rng.random(n) < 0.2
    dst[uni] = rng.integers([-40, -40], [w + 40, h + 40], size=(int(uni.sum()), 2))
    d = rng.integers(-4, 5, size=(n, 2))
    r["dst_x"], r["dst_y"] = dst[:, 0], dst[:, 1]
    r["src_x"], r["src_y"] = dst[:, 0] - d[:, 0], dst[:, 1] - d[:, 1]
    return r


@pytest.mark.parametrize("seed", range(12))
def test_oracle_vs_numpy_random(seed):
    rng = np.random.default_rng(seed)
    w, h = [(1920, 1080), (3840, 2160), (1280, 720), (352, 288)][seed % 4]
    p = kats.env_params() if seed % 2 else kats.code_defaults()
    p.vectors_needed = int(rng.integers(0, 6))
    p.clusters_needed = int(rng.integers(0, 5))
    p.mv_threshold_sq = float(rng.choice([0.0, 1.0, 4.0, 4.5, 16.0, 25.0]))
    cfg = cfg_for(p, w, h)
    for _ in range(6):
        recs = random_frame(rng, int(rng.integers(1, 4000)), w, h, int(rng.integers(1, 6)))
        want = np_full_count(p, cfg.grid_w, cfg.grid_h, cfg.vertical_margin, recs)
        assert orc.full_count(cfg, recs) == want
        assert orc.check_frame(cfg, recs) == int(want >= max(1, p.clusters_needed))


def test_scan_frames_mt_matches_single():
    spec = ms.synth_preset(3, 7)
    cnt, off, recs, pts = ms.synth_host(spec, 0, 240)
    p = kats.env_params()
    cfg = cfg_for(p)
    f1, c1 = orc.scan_frames(cfg, recs, off)
    f4, c4 = orc.scan_frames(cfg, recs, off, threads=4)
    fe, _ = orc.scan_frames(cfg, recs, off, early_exit=True)
    assert np.array_equal(f1, f4) and np.array_equal(c1, c4)
    assert np.array_equal(f1, fe)  # early-exit flag == closed form on the full count
    assert f1.sum() > 0 and f1.sum() < len(f1)  # the preset produces both active and static frames


def py_tail(ts, duration, gap, pad, min_pct):
    """Independent pure-Python restatement of Appendix B."""
    ts = sorted(set(ts))
    if not ts:
        return [], 0.0, 0.0, 0.0, ms.NO_MOTION
    segs, cs, last = [], ts[0], ts[0]
    for t in ts[1:]:
        if t - last > gap:
            segs.append([max(0.0, cs - pad), last + pad])
            cs = t
        last = t
    segs.append([max(0.0, cs - pad), last + pad])
    out = 0.0
    for s in segs:
        s[1] = min(s[1], duration)
        s[0] = min(s[0], s[1])
        out += s[1] - s[0]
    removed = duration - out
    pct = removed / duration * 100.0 if duration > 0 else 0.0
    return segs, out, removed, pct, (ms.CUT if pct > min_pct else ms.FULL_COPY)


@pytest.mark.parametrize("seed", range(10))
def test_tail_vs_python_random(seed):
    rng = np.random.default_rng(100 + seed)
    duration = float(rng.choice([60.0, 600.0, 37.25]))
    n = int(rng.integers(0, 3000))
    fps = float(rng.choice([30.0, 25.0, 29.97]))
    frames = np.sort(rng.choice(int(duration * fps), size=min(n, int(duration * fps)), replace=False))
    pts = frames / fps
    if seed % 3 == 0:  # unsorted with duplicates, as chunk workers may deliver them
        pts = np.concatenate([pts, pts[: len(pts) // 3]])
        rng.shuffle(pts)
    flags = (rng.random(len(pts)) < 0.7).astype(np.uint8)
    gap, pad, min_pct = float(rng.choice([5.0, 1.0, 0.0])), float(rng.choice([0.5, 2.0, 0.0])), 5.0
    segs, res = orc.video_tail(pts, flags, duration, gap, pad, min_pct)
    wsegs, wout, wrem, wpct, wdec = py_tail([float(t) for t, f in zip(pts, flags) if f], duration, gap, pad, min_pct)
    assert res.decision == wdec
    assert [(s["start"], s["end"]) for s in segs] == [tuple(s) for s in wsegs]
    assert (res.out_dur, res.time_removed, res.saved_pct) == (wout, wrem, wpct)
