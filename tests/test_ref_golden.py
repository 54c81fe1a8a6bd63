"""CPU: pins the oracle against outputs of the REFERENCE's own code (tests/golden/ref_golden.json, made
by tests/golden/make_golden.py from oracle/_ref/ref_scan). When the reference binary is present (build
container) a few cases are also re-run live so the fixture cannot silently rot."""
import json
import tempfile
from functools import lru_cache
from pathlib import Path

import numpy as np
import pytest

import golden_cases as gc
import kats
from motionscan import mvs_io
import oracle_lib as orc
import ref_runner

GOLDEN = json.loads((Path(__file__).parent / "golden" / "ref_golden.json").read_text())["cases"]


@lru_cache(maxsize=None)
def cases():
    return {c.name: c for c in gc.all_cases()}


def fx(x):
    return float.fromhex(x)


def expected(name):
    g = GOLDEN[name]
    return dict(ts=np.array([fx(t) for t in g["ts"]]), segs=np.array([[fx(a), fx(b)] for a, b in g["segments"]]).reshape(-1, 2),
                decision=g["decision"], duration=fx(g["duration"]), time_removed=fx(g["time_removed"]), saved_pct=fx(g["saved_pct"]))


def job_segments(segs, res, duration):
    """What the FFmpegJob carries (pipeline.cpp:360-369 / :384-394)."""
    if res.decision == 1:
        return np.stack([segs["start"], segs["end"]], 1).reshape(-1, 2)
    if res.decision == 2:
        return np.array([[0.0, duration]])
    return np.zeros((0, 2))


def test_fixture_covers_every_case():
    assert sorted(GOLDEN) == sorted(cases())
    assert len(GOLDEN) >= 25


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_oracle_matches_reference_outputs(name):
    c, g, e = cases()[name], GOLDEN[name], expected(name)
    assert gc.digest(c.cnt, c.recs, c.ticks) == g["digest"], "regenerated input differs from the one the reference saw"
    assert gc.params_dict(c.params) == g["params"]
    assert c.duration == e["duration"]
    p = c.params
    gw, gh, m = orc.geometry(c.width, c.height, p.block_size, p.block_shift, p.vertical_mask)
    cfg = orc.make_cfg(p, gw, gh, m)
    # leg 1: one scan_range(0, duration): frame selection (:303-371) + check_frame (:217-295)
    sel1 = c.selected(chunked=False)
    cnt1, off1, recs1, pts1 = c.subset(sel1)
    flags_ee, _ = orc.scan_frames(cfg, recs1, off1, early_exit=True)
    flags_fc, counts = orc.scan_frames(cfg, recs1, off1, early_exit=False)
    assert np.array_equal(flags_ee, flags_fc)
    assert pts1[flags_ee.astype(bool)].tobytes() == e["ts"].tobytes()  # exactly what scan_range returned (:382-383)
    # leg 2: the chunked pipeline → FFmpegJob
    sel2 = c.selected(chunked=True)
    cnt2, off2, recs2, pts2 = c.subset(sel2)
    flags2, _ = orc.scan_frames(cfg, recs2, off2, early_exit=True)
    if not c.target_fps:
        assert np.array_equal(np.sort(sel1), np.sort(sel2))  # without skipping, chunking selects the same frames
    segs, res = orc.video_tail(pts2, flags2, e["duration"], p.max_gap_sec, p.padding_sec, p.min_savings_pct)
    assert res.decision == e["decision"]
    assert job_segments(segs, res, e["duration"]).tobytes() == e["segs"].tobytes()
    if res.decision != 0:  # no-motion returns before the savings are computed (pipeline.cpp:308-319)
        assert np.float64(res.time_removed).tobytes() == np.float64(e["time_removed"]).tobytes()
        assert np.float64(res.saved_pct).tobytes() == np.float64(e["saved_pct"]).tobytes()


def test_kat_table_agrees_with_reference_flags():
    """SURVEY §4's K-table flags are what the real reference returns."""
    table = kats.frame_kats()
    for c in gc.kat_frame_cases():
        e = expected(c.name)
        got = set(np.round(e["ts"] * 30).astype(int))
        for i, kname in enumerate(c.kat_names):
            assert (i in got) == bool(table[kname][2]), kname


@pytest.mark.skipif(not ref_runner.available(), reason="oracle/_ref/ref_scan not built (needs /root/reference)")
@pytest.mark.parametrize("name", ["kat_frames_0", "kat_seg_S4", "batchclip_seed100", "rand_params_5", "skip_tfps7_chunk2p5"])
def test_live_reference_reproduces_fixture(name):
    c, e = cases()[name], expected(name)
    with tempfile.TemporaryDirectory() as d:
        path = Path(d) / "x.mvs"
        mvs_io.write_mvs(path, c.width, c.height, c.fps[0], c.fps[1], c.ticks, c.cnt, c.recs, has_mvs=c.has_mvs,
                         tb_num=c.tb[0], tb_den=c.tb[1], duration_us=c.duration_us)
        r = ref_runner.run(path, c.params, threads=c.threads, chunk_sec=c.chunk_sec, target_fps=c.target_fps)
    assert r["ts"].tobytes() == e["ts"].tobytes()
    assert r["segs"].tobytes() == e["segs"].tobytes()
    assert r["decision"] == e["decision"] and r["saved_pct"] == e["saved_pct"]


@pytest.mark.skipif(not ref_runner.available(), reason="oracle/_ref/ref_scan not built (needs /root/reference)")
@pytest.mark.parametrize("seed", range(4))
def test_live_reference_vs_oracle_random_streams(seed):
    """Fresh random streams (not in the fixture): real reference vs oracle, live."""
    from test_oracle_kats import random_frame

    rng = np.random.default_rng(7000 + seed)
    w, h = [(1920, 1080), (1280, 720), (352, 288), (3840, 2160)][seed]
    p = kats.env_params(vectors_needed=int(rng.integers(1, 5)), clusters_needed=int(rng.integers(1, 4)),
                        mv_threshold_sq=float(rng.choice([1.0, 4.0, 9.5])), max_gap_sec=1.0)
    n = 240
    frames = [random_frame(rng, int(rng.integers(1, 3000)), w, h, int(rng.integers(1, 5))) if i % 30 else None for i in range(n)]
    cnt = np.array([0 if f is None else len(f) for f in frames], np.uint32)
    recs = kats.cat(*[f for f in frames if f is not None])
    c = gc.Case("live", w, h, (30, 1), (1, 30), np.arange(n), cnt, recs, p)
    with tempfile.TemporaryDirectory() as d:
        path = Path(d) / "x.mvs"
        mvs_io.write_mvs(path, w, h, 30, 1, c.ticks, cnt, recs)
        r = ref_runner.run(path, p, threads=3, chunk_sec=2.0)
    gw, gh, m = orc.geometry(w, h, p.block_size, p.block_shift, p.vertical_mask)
    flags, _ = orc.scan_frames(orc.make_cfg(p, gw, gh, m), recs, c.off)
    assert c.pts[flags.astype(bool)].tobytes() == r["ts"].tobytes()
    segs, res = orc.video_tail(c.pts, flags, r["duration"], p.max_gap_sec, p.padding_sec, p.min_savings_pct)
    assert res.decision == r["decision"]
    assert job_segments(segs, res, r["duration"]).tobytes() == r["segs"].tobytes()
