"""Property tests (hypothesis): arbitrary geometries, knobs and record streams — extreme int16 coordinates,
out-of-frame and negative destinations, fractional / negative / huge thresholds, VECTORS_NEEDED beyond uint8,
unsorted and duplicated timestamps.
  CPU: the C oracle against an independent numpy / pure-Python model (no shared code).
  GPU: the CUDA path (both record layouts) against the oracle, bit for bit."""
import math

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import kats
import motionscan as ms
import oracle_lib as orc
from test_oracle_kats import np_full_count

COMMON = dict(deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)


@st.composite
def scan_case(draw, max_dim=2200, max_recs=3000):
    w = draw(st.integers(16, max_dim))
    h = draw(st.integers(16, max_dim))
    p = ms.default_params()
    p.mv_threshold_sq = draw(st.sampled_from([0.0, 0.5, 1.0, 3.999, 4.0, 4.0001, 16.0, 100.0, -1.0, 1e12, 2147483647.0, 2147483648.0]))
    p.vectors_needed = draw(st.sampled_from([0, 1, 2, 3, 4, 8, 255, 256, 257, 300]))  # wraps to uint8 (config.hpp:75)
    p.clusters_needed = draw(st.integers(-1, 5))
    p.vertical_mask = draw(st.sampled_from([0.0, 0.02, 0.05, 0.1, 0.25, 0.49]))
    seed = draw(st.integers(0, 2**32 - 1))
    n = draw(st.integers(0, max_recs))
    style = draw(st.sampled_from(["hot", "uniform", "extreme", "one-cell"]))
    rng = np.random.default_rng(seed)
    r = np.zeros(n, dtype=ms.MV_DTYPE)
    if style == "hot":
        c = rng.integers(0, [w, h], size=(max(1, n // 200 + 1), 2))
        dst = c[rng.integers(0, len(c), n)] + rng.integers(-20, 21, size=(n, 2))
    elif style == "uniform":
        dst = rng.integers([-64, -64], [w + 64, h + 64], size=(n, 2))
    elif style == "extreme":
        dst = rng.choice(np.array([-32768, -1, 0, 1, 15, 16, 17, w - 1, w, h - 1, h, 32767]), size=(n, 2))
    else:
        dst = np.tile(rng.integers(0, [w, h], size=(1, 2)), (n, 1))
    dst = np.clip(dst, -32768, 32767)
    d = rng.integers(-9, 10, size=(n, 2)) if style != "extreme" else rng.choice(np.array([-200, -2, -1, 0, 1, 2, 200]), size=(n, 2))
    src = np.clip(dst - d, -32768, 32767)
    r["dst_x"], r["dst_y"], r["src_x"], r["src_y"] = dst[:, 0], dst[:, 1], src[:, 0], src[:, 1]
    return w, h, p, r


def oracle_cfg(p, w, h):
    gw, gh, m = orc.geometry(w, h, p.block_size, p.block_shift, p.vertical_mask)
    return orc.make_cfg(p, gw, gh, m), gw, gh, m


@settings(max_examples=150, **COMMON)
@given(scan_case())
def test_oracle_matches_numpy_model(case):
    w, h, p, recs = case
    cfg, gw, gh, m = oracle_cfg(p, w, h)
    # rec_count == 0 stands for "no MV side data" ⇒ false (motion_scanner.cpp:219-221), whatever the knobs
    want = np_full_count(p, gw, gh, m, recs) if (len(recs) and gw > 2 and gh > 2 * m) else 0
    got = orc.full_count(cfg, recs if len(recs) else None)
    assert got == want
    assert orc.check_frame(cfg, recs if len(recs) else None) == int(want >= max(1, p.clusters_needed))


def py_tail(ts, duration, max_gap, pad, min_pct):
    """Pure-Python restatement of pipeline.cpp:302-404 (sort, unique, gap merge, clamp, savings, decision)."""
    ts = sorted(set(ts))
    if not ts:
        return [], 0, 0.0
    segs, start, last = [], ts[0], ts[0]
    for t in ts[1:]:
        if t - last > max_gap:
            segs.append((max(0.0, start - pad), last + pad))
            start = t
        last = t
    segs.append((max(0.0, start - pad), last + pad))
    out, clamped = 0.0, []
    for a, b in segs:
        b = min(b, duration)
        a = min(a, b)
        out += b - a
        clamped.append((a, b))
    pct = (duration - out) / duration * 100.0 if duration > 0 else 0.0
    return clamped, (1 if pct > min_pct else 2), pct


tail_case = st.tuples(
    st.lists(st.one_of(st.floats(0, 500, allow_nan=False), st.integers(0, 15000).map(lambda k: k / 30.0)), max_size=400),
    st.sampled_from([0.0, 0.5, 5.0, 5.000000001, 33.3]), st.sampled_from([0.0, 0.5, 2.0]), st.sampled_from([0.0, 5.0, 50.0, 99.9]),
    st.sampled_from([1.0, 60.0, 500.0, 600.0]))


@settings(max_examples=200, **COMMON)
@given(tail_case)
def test_oracle_tail_matches_python_model(case):
    ts, gap, pad, min_pct, duration = case
    pts = np.array(ts, dtype=np.float64)
    segs, res = orc.video_tail(pts, np.ones(len(ts), np.uint8), duration, gap, pad, min_pct)
    want, decision, pct = py_tail(ts, duration, gap, pad, min_pct)
    assert [(s["start"], s["end"]) for s in segs] == want
    assert res.decision == (decision if ts else 0)
    if ts:
        assert res.saved_pct == pct or (math.isnan(res.saved_pct) and math.isnan(pct))


# ------------------------------------------------------------------------------------------ GPU ----
@pytest.mark.gpu
@settings(max_examples=60, **COMMON)
@given(scan_case(max_dim=4400, max_recs=6000), st.sampled_from(["native", "projected"]))
def test_gpu_matches_oracle_on_arbitrary_frames(case, mode):
    w, h, p, recs = case
    cfg, gw, gh, m = oracle_cfg(p, w, h)
    frames = [recs, None, recs[: len(recs) // 2], recs[::-1].copy()]
    cnt = np.array([0 if f is None else len(f) for f in frames], np.uint32)
    allr = kats.cat(*[f for f in frames if f is not None and len(f)])
    with ms.Context(0, p, 1 << 12, 8 << 20) as ctx:
        ctx.set_staging_mode(ms.STAGING_NATIVE if mode == "native" else ms.STAGING_AUTO)
        ctx.video_open(1, w, h)
        ctx.submit(1, np.arange(4) / 30.0, cnt, allr if len(allr) else None)
        flags, counts = ctx.collect(1)
    for i, f in enumerate(frames):
        f = f if (f is not None and len(f)) else None
        assert counts[i] == orc.full_count(cfg, f), (i, mode)
        assert flags[i] == orc.check_frame(cfg, f), (i, mode)


@pytest.mark.gpu
@settings(max_examples=60, **COMMON)
@given(tail_case, st.integers(0, 2**32 - 1))
def test_gpu_tail_matches_oracle(case, seed):
    ts, gap, pad, min_pct, duration = case
    if not ts:
        ts = [1.0]
    rng = np.random.default_rng(seed)
    pts = np.array(ts, dtype=np.float64)
    flags = (rng.random(len(ts)) < 0.7).astype(np.uint8)
    p = kats.env_params(max_gap_sec=gap, padding_sec=pad, min_savings_pct=min_pct)
    osegs, ores = orc.video_tail(pts, flags, duration, gap, pad, min_pct)
    n = len(ts)
    with ms.Context(0, p, 1 << 12, 8 << 20) as ctx:
        d_pts, d_flags, d_segs, d_res = ctx.dev_alloc(8 * n), ctx.dev_alloc(n), ctx.dev_alloc(16 * n), ctx.dev_alloc(40)
        ctx.h2d(d_pts, pts)
        ctx.h2d(d_flags, flags)
        ctx.segments_device([0, n], [duration], d_pts, d_flags, d_segs, d_res)
        ctx.sync()
        res = np.zeros(1, ms.RESULT_DTYPE)
        ctx.d2h(res, d_res)
        segs = np.zeros(int(res["n_segments"][0]), ms.SEG_DTYPE)
        if len(segs):
            ctx.d2h(segs, d_segs)
        for d in (d_pts, d_flags, d_segs, d_res):
            ctx.dev_free(d)
    assert int(res["decision"][0]) == ores.decision and int(res["n_motion_frames"][0]) == ores.n_motion_frames
    assert segs.tobytes() == osegs.tobytes()
    assert np.float64(res["saved_pct"][0]).tobytes() == np.float64(ores.saved_pct).tobytes()
    assert np.float64(res["out_dur"][0]).tobytes() == np.float64(ores.out_dur).tobytes()
