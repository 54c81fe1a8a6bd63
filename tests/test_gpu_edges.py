"""GPU: edge cases and size-independent properties — degenerate knobs, empty inputs, concurrent submits,
idempotence / batching invariance at a large device-resident size with oracle spot checks."""
import os
import threading
import zlib

import numpy as np
import pytest

import kats
import motionscan as ms
import oracle_lib as orc
from test_gpu_parity import cfg_for, oracle_tail, run_frames
from test_oracle_kats import random_frame

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("knobs", [
    dict(mv_threshold_sq=float("nan")),          # every `<` against NaN is false ⇒ all vectors kept
    dict(mv_threshold_sq=float("inf")),          # nothing is ever kept
    dict(mv_threshold_sq=-5.0),                  # negative threshold keeps zero-length vectors too
    dict(mv_threshold_sq=3e9),                   # above INT32_MAX
    dict(mv_threshold_sq=0.5),                   # fractional: ceil() on the device
    dict(vectors_needed=0),                      # every cell active (also masked rows, as neighbours)
    dict(vectors_needed=256),                    # wraps to 0 (config.hpp:75)
    dict(vectors_needed=260),                    # wraps to 4
    dict(vectors_needed=255),
    dict(clusters_needed=0),
    dict(clusters_needed=-3),
    dict(clusters_needed=100000),
    dict(block_size=8, block_shift=3),           # 240x135 grid on 1080p
    dict(block_size=32, block_shift=5),
    dict(vertical_mask=0.2),
    dict(vertical_mask=0.49),                    # 2 live rows on 1080p (margin 33 of 68)
])
def test_degenerate_knobs(knobs):
    p = kats.env_params(**knobs)
    rng = np.random.default_rng(zlib.crc32(repr(sorted(knobs.items())).encode()))
    frames = [random_frame(rng, int(rng.integers(1, 5000)), kats.W, kats.H, 4) for _ in range(10)] + [None, kats.cell(10, 10, 300)]
    cfg = cfg_for(p, kats.W, kats.H)
    assert cfg.vertical_margin >= 1
    with ms.Context(0, p) as ctx:
        flags, counts = run_frames(ctx, 1, kats.W, kats.H, frames)
    for i, f in enumerate(frames):
        assert counts[i] == orc.full_count(cfg, f), (knobs, i)
        assert flags[i] == orc.check_frame(cfg, f), (knobs, i)


def test_empty_inputs():
    p = kats.env_params()
    with ms.Context(0, p) as ctx:
        ctx.video_open(1, 1920, 1080)
        flags, counts = ctx.collect(1)                       # a video nobody submitted to
        assert len(flags) == 0 and len(counts) == 0
        segs, res = ctx.segments(1, 60.0)
        assert res.decision == ms.NO_MOTION and len(segs) == 0
        assert ctx.submit(1, np.zeros(0), np.zeros(0, np.uint32), None) == 0   # zero frames is a no-op
        ctx.submit(1, np.arange(5) / 30.0, np.zeros(5, np.uint32), None)       # frames without side data
        flags, counts = ctx.collect(1)
        assert not flags.any() and not counts.any() and len(flags) == 5
        segs, res = ctx.segments(1, 60.0)
        assert res.decision == ms.NO_MOTION and res.n_motion_frames == 0
        ctx.video_close(1)
        segs, off, res = ctx.segments_batch([], [])
        assert len(segs) == 0 and len(res) == 0


def test_concurrent_submits_from_many_threads():
    """Decode workers of several videos call mscan_submit concurrently (motion_scanner.hpp:8-13: one scanner
    per thread); chunks of one video may arrive in any order — K-C sorts them like pipeline.cpp:302-304."""
    p = kats.env_params()
    specs = {v: ms.synth_preset(0, 200 + v) for v in range(4)}
    data = {v: ms.synth_host(s, 0, 240) for v, s in specs.items()}
    errors = []
    with ms.Context(0, p, 0, 16 << 20) as ctx:
        for v, s in specs.items():
            ctx.video_open(v, s.width, s.height)

        def worker(v, chunks):
            try:
                cnt, off, recs, pts = data[v]
                for a in chunks:
                    b = a + 30
                    ctx.submit(v, pts[a:b], cnt[a:b], recs[int(off[a]) : int(off[b])])
            except Exception as e:  # pragma: no cover
                errors.append(e)

        threads = []
        for v in specs:  # two workers per video, interleaved chunk order
            threads.append(threading.Thread(target=worker, args=(v, [0, 60, 120, 180])))
            threads.append(threading.Thread(target=worker, args=(v, [210, 150, 90, 30])))
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errors
        segs, soff, res = ctx.segments_batch(list(specs), [8.0] * 4)
        motion = {v: ctx.motion_segments(v, 8.0) for v in specs}
        collected = {v: ctx.collect(v) for v in specs}
    for i, v in enumerate(specs):
        cnt, off, recs, pts = data[v]
        of, oc = orc.scan_frames(cfg_for(p, 1920, 1080), recs, off, threads=4)
        assert collected[v][0].sum() == of.sum() and collected[v][1].sum() == oc.sum()  # same frames, arrival order differs
        osegs, ores = oracle_tail(p, pts, of, 8.0)
        assert res["decision"][i] == ores.decision and res["n_motion_frames"][i] == ores.n_motion_frames
        assert motion[v][0].tobytes() == osegs.tobytes()
        assert np.float64(res["saved_pct"][i]).tobytes() == np.float64(ores.saved_pct).tobytes()


@pytest.mark.parametrize("n_threads,batch", [(16, 1), (24, 3), (5, 1)])
def test_decode_worker_standins_submit_per_frame(n_threads, batch):
    """16+ decode-worker stand-ins (csrc/feed_harness.cpp) each write their frames' native records into a private,
    pageable side-data buffer and call mscan_submit PER FRAME, concurrently, on shared videos: the reserve / fill /
    commit submit must give every frame exactly the oracle's result, whatever the interleaving. Small slabs and copy
    windows so that slabs flip, windows close and videos wrap many times."""
    p = kats.env_params()
    specs = [ms.synth_preset(0, 300 + v) for v in range(3)]
    data = [ms.synth_host(s, 0, 360) for s in specs]
    cnt = np.concatenate([d[0] for d in data])
    pts = np.concatenate([d[3] for d in data])
    recs = kats.cat(*[d[2] for d in data])
    off = np.zeros(len(cnt) + 1, np.uint64)
    np.cumsum(cnt, out=off[1:])
    voff = np.array([0, 360, 720, 1080], np.uint64)
    r8 = ms.pack_records(recs)
    cfg = cfg_for(p, 1920, 1080)
    of, oc = orc.scan_frames(cfg, recs, off, threads=4)
    os.environ["MSCAN_COPY_WINDOW_KB"] = "256"
    try:
        with ms.Context(0, p, 0, 4 << 20) as ctx:
            for rep in range(3):
                for v in range(3):
                    ctx.video_open(v, 1920, 1080)
                res, index = ms.feed_run(ctx, [0, 1, 2], voff, pts, cnt, off, r8, n_threads=n_threads, frames_per_submit=batch)
                assert res.records == int(off[-1]) and res.frames == 1080 and res.submits >= 1080 // batch
                segs, soff, vres = ctx.segments_batch([0, 1, 2], [12.0] * 3)
                for v in range(3):
                    a, b = int(voff[v]), int(voff[v + 1])
                    fl, cn = ctx.collect(v)
                    idx = index[a:b].astype(np.int64)
                    assert sorted(idx.tolist()) == list(range(b - a))        # every frame got its own slot
                    assert np.array_equal(fl[idx], of[a:b]) and np.array_equal(cn[idx], oc[a:b]), (rep, v)
                    osegs, ores = oracle_tail(p, pts[a:b], of[a:b], 12.0)
                    assert vres["decision"][v] == ores.decision and vres["n_motion_frames"][v] == ores.n_motion_frames
                    job = segs[int(soff[v]) : int(soff[v + 1])]
                    if ores.decision == ms.CUT:
                        assert job.tobytes() == osegs.tobytes()
                    ctx.video_close(v)
            st = ctx.stats()
            assert st.records_projected == 3 * int(off[-1])  # pageable native records: projected by the submitting threads
    finally:
        del os.environ["MSCAN_COPY_WINDOW_KB"]


def test_video_tail_does_not_drain_other_videos():
    """collect / segments / close of one video wait only for the slabs holding ITS frames: a second video with a large
    pinned submit in flight is not drained by the first video's tail (its results are still correct afterwards)."""
    p = kats.env_params()
    sa, sb = ms.synth_preset(0, 410), ms.synth_preset(0, 411)
    a = ms.synth_host(sa, 0, 60)
    b = ms.synth_host(sb, 0, 900)
    cfg = cfg_for(p, 1920, 1080)
    with ms.Context(0, p, 0, 8 << 20) as ctx:
        ctx.video_open(1, 1920, 1080)
        ctx.video_open(2, 1920, 1080)
        for rep in range(4):
            ctx.submit(1, a[3], a[0], a[2])
            ctx.submit(2, b[3] + 30.0 * rep, b[0], b[2])
            fl, cn = ctx.collect_range(1, 60 * rep, 60)
            of, oc = orc.scan_frames(cfg, a[2], a[1])
            assert np.array_equal(fl, of) and np.array_equal(cn, oc)
        fl2, cn2 = ctx.collect(2)
        of2, oc2 = orc.scan_frames(cfg, b[2], b[1], threads=4)
        assert np.array_equal(fl2, np.tile(of2, 4)) and np.array_equal(cn2, np.tile(oc2, 4))
        ctx.video_close(1)
        ctx.video_close(2)


def test_fullsize_properties_device_resident():
    """~2x10^8 records on the device (8 GB): scanning is idempotent, independent of how the frames are split
    into launches, and agrees with the oracle on randomly chosen frames regenerated on the host."""
    p = kats.env_params()
    spec = ms.synth_preset(4, 5)
    n = 20000
    with ms.Context(0, p) as ctx:
        d_cnt, d_off = ctx.dev_alloc(4 * n), ctx.dev_alloc(8 * (n + 1))
        ctx.synth_counts(spec, 0, n, d_cnt)
        ctx.offsets_from_counts(d_cnt, n, d_off)
        ctx.sync()
        off = np.zeros(n + 1, np.uint64)
        ctx.d2h(off, d_off)
        n_rec = int(off[-1])
        assert n_rec > 1.9e8
        d_recs, d_pts = ctx.dev_alloc(40 * n_rec + 256), ctx.dev_alloc(8 * n)
        ctx.synth_fill(spec, 0, n, d_off, d_recs, d_pts)
        geom = ms.geometry_from_dims(p, spec.width, spec.height)
        outs = []
        for rep in range(2):  # idempotence
            d_f, d_c = ctx.dev_alloc(n), ctx.dev_alloc(4 * n)
            ctx.scan_device(d_recs, d_off, None, [geom], n, d_f, d_c)
            ctx.sync()
            f, c = np.zeros(n, np.uint8), np.zeros(n, np.uint32)
            ctx.d2h(f, d_f)
            ctx.d2h(c, d_c)
            outs.append((f, c))
            ctx.dev_free(d_f)
            ctx.dev_free(d_c)
        assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
        # the same frames in three unequal launches (offsets shifted: each launch sees a sub-range of rec_off)
        d_f, d_c = ctx.dev_alloc(n), ctx.dev_alloc(4 * n)
        for a, b in [(0, 7), (7, 12345), (12345, n)]:
            ctx.scan_device(d_recs, d_off + 8 * a, None, [geom], b - a, d_f + a, d_c + 4 * a)
        ctx.sync()
        f3, c3 = np.zeros(n, np.uint8), np.zeros(n, np.uint32)
        ctx.d2h(f3, d_f)
        ctx.d2h(c3, d_c)
        assert np.array_equal(f3, outs[0][0]) and np.array_equal(c3, outs[0][1])
        # layout invariance: the 8-byte projection of the stream (device-side) scans to the same log
        d_r8 = ctx.dev_alloc(8 * n_rec + 256)
        ctx.pack_records_device(d_recs, n_rec, d_r8)
        ctx.scan_device_packed(d_r8, d_off, None, [geom], n, d_f, d_c)
        ctx.sync()
        f8, c8 = np.zeros(n, np.uint8), np.zeros(n, np.uint32)
        ctx.d2h(f8, d_f)
        ctx.d2h(c8, d_c)
        assert np.array_equal(f8, outs[0][0]) and np.array_equal(c8, outs[0][1])
        for d in (d_cnt, d_off, d_recs, d_pts, d_f, d_c, d_r8):
            ctx.dev_free(d)
    flags, counts = outs[0]
    assert 0 < flags.sum() < n and counts.max() > 10
    assert not flags[::30].any()  # I-frames (no records) are never active
    # oracle spot checks: 200 random frames + the busiest ones, regenerated on the host
    rng = np.random.default_rng(1)
    pick = np.unique(np.concatenate([rng.integers(0, n, 200), np.argsort(counts)[-20:]]))
    cfg = cfg_for(p, spec.width, spec.height)
    for i in pick:
        cnt1, off1, recs1, _ = ms.synth_host(spec, int(i), 1, n_threads=1)
        assert int(cnt1[0]) == int(off[i + 1] - off[i])
        assert counts[i] == orc.full_count(cfg, recs1 if len(recs1) else None), i
        assert flags[i] == orc.check_frame(cfg, recs1 if len(recs1) else None), i


def test_host_register_and_fence_allow_buffer_reuse():
    """A caller-owned buffer pinned with mscan_host_register is DMA'd in place; after mscan_host_fence it may
    be overwritten and submitted again (the double-buffering contract of the staging path)."""
    p = kats.env_params()
    a = ms.synth_host(ms.synth_preset(0, 31), 0, 90)
    b = ms.synth_host(ms.synth_preset(0, 32), 0, 90)
    n = max(len(a[2]), len(b[2]))
    buf = np.zeros(n, ms.MV_DTYPE)
    with ms.Context(0, p, 0, 8 << 20) as ctx:
        ctx.host_register(buf)
        ctx.video_open(1, 1920, 1080)
        ctx.video_open(2, 1920, 1080)
        buf[: len(a[2])] = a[2]
        ctx.submit(1, a[3], a[0], buf)
        ctx.host_fence()                 # every byte of video 1 has left `buf`
        buf[: len(b[2])] = b[2]          # overwrite while video 1's kernels may still be running
        ctx.submit(2, b[3], b[0], buf)
        f1, c1 = ctx.collect(1)
        f2, c2 = ctx.collect(2)
        st = ctx.stats()
        ctx.host_unregister(buf)
    cfg = cfg_for(p, 1920, 1080)
    for (cnt, off, recs, pts), f, c in ((a, f1, c1), (b, f2, c2)):
        of, oc = orc.scan_frames(cfg, recs, off, threads=4)
        assert np.array_equal(c, oc) and np.array_equal(f, of)
    assert f1.any() and f2.any() and not np.array_equal(c1, c2)
    assert st.h2d_bytes >= 40 * (len(a[2]) + len(b[2]))


@pytest.mark.parametrize("scrambled", [False, True])
def test_kc_ten_hour_video(scrambled):
    """K-C at its largest realistic size: a 10-hour 30 fps video (1 080 000 frames), in order and with its
    900-frame chunks delivered in random order plus duplicated chunks (sort + unique over 2^21 slots);
    segments, savings and decision bit-exact with the oracle's restatement of pipeline.cpp:302-404."""
    n, fps = 1_080_000, 30.0
    rng = np.random.default_rng(11)
    pts = np.arange(n, dtype=np.float64) / fps
    flags = np.zeros(n, np.uint8)
    for a in rng.integers(0, n - 3000, 400):  # 400 motion events of 1..100 s
        flags[a : a + int(rng.integers(30, 3000))] = 1
    flags &= (rng.random(n) < 0.9).astype(np.uint8)  # flicker inside events
    if scrambled:
        order = rng.permutation(n // 900)
        order = np.concatenate([order, order[:25]])  # some chunks arrive twice
        idx = (order[:, None] * 900 + np.arange(900)[None, :]).reshape(-1)
        pts, flags = pts[idx], flags[idx]
    p = kats.env_params()
    duration = n / fps
    osegs, ores = orc.video_tail(pts, flags, duration, p.max_gap_sec, p.padding_sec, p.min_savings_pct)
    m = len(pts)
    with ms.Context(0, p) as ctx:
        d_pts, d_flags = ctx.dev_alloc(8 * m), ctx.dev_alloc(m)
        d_segs, d_res = ctx.dev_alloc(16 * m), ctx.dev_alloc(40)
        ctx.h2d(d_pts, pts)
        ctx.h2d(d_flags, flags)
        ctx.set_profiling(True)
        for _ in range(2):
            ctx.segments_device([0, m], [duration], d_pts, d_flags, d_segs, d_res)
        ctx.sync()
        st = ctx.stats()
        res = np.zeros(1, ms.RESULT_DTYPE)
        ctx.d2h(res, d_res)
        segs = np.zeros(int(res["n_segments"][0]), ms.SEG_DTYPE)
        ctx.d2h(segs, d_segs)
        for d in (d_pts, d_flags, d_segs, d_res):
            ctx.dev_free(d)
    assert int(res["decision"][0]) == ores.decision and int(res["n_motion_frames"][0]) == ores.n_motion_frames
    assert segs.tobytes() == osegs.tobytes() and len(segs) > 100
    assert np.float64(res["saved_pct"][0]).tobytes() == np.float64(ores.saved_pct).tobytes()
    print(f"K-C {m} frames scrambled={scrambled}: {st.segment_ms / st.segment_launches:.3f} ms per launch")
    assert st.segment_ms / st.segment_launches < 2000.0


def test_frame_log_is_reused_after_videos_close():
    """A long-running context (watch mode, many streams) never sees a moment without open videos, so the frame log
    cannot wait for one to rewind: frames of closed videos are reused first-fit, a video's frames may wrap around the
    log in several extents, and only a log whose every frame belongs to an open video reports MSCAN_ERR_CAPACITY."""
    p = kats.env_params()
    spec = ms.synth_preset(3, 9)
    cnt, off, recs, pts = ms.synth_host(spec, 0, 200)
    cfg = cfg_for(p, spec.width, spec.height)
    of, oc = orc.scan_frames(cfg, recs, off, threads=4)

    def feed(ctx, vid, a, b):
        ctx.submit(vid, pts[a:b], cnt[a:b], recs[int(off[a]) : int(off[b])])

    with ms.Context(0, p, max_log_frames=64, slab_bytes=8 << 20) as ctx:
        for v in (1, 2, 3):
            ctx.video_open(v, spec.width, spec.height)
        feed(ctx, 1, 0, 40)
        feed(ctx, 2, 40, 60)          # log: [1: 0..40) [2: 40..60) free 60..64
        ctx.video_close(1)            # frees 0..40 while video 2 stays open
        feed(ctx, 3, 100, 130)        # 4 frames at 60..64, then wraps into 0..26
        f3, c3 = ctx.collect(3)
        assert np.array_equal(c3, oc[100:130]) and np.array_equal(f3, of[100:130])
        f2, c2 = ctx.collect(2)
        assert np.array_equal(c2, oc[40:60]) and np.array_equal(f2, of[40:60])  # untouched by the reuse
        segs3, res3 = ctx.motion_segments(3, 30 / spec.fps)
        osegs, ores = orc.video_tail(pts[100:130], of[100:130], 30 / spec.fps, p.max_gap_sec, p.padding_sec, p.min_savings_pct)
        assert segs3.tobytes() == osegs.tobytes() and res3.n_motion_frames == ores.n_motion_frames
        feed(ctx, 2, 60, 74)          # fills 26..40: the log is now completely owned by videos 2 and 3
        with pytest.raises(ms.MscanError) as e:
            feed(ctx, 3, 130, 131)
        assert e.value.code == ms.ERR_CAPACITY
        ctx.video_close(3)            # 30 frames come back
        ctx.video_open(4, spec.width, spec.height)
        feed(ctx, 4, 150, 180)
        f4, c4 = ctx.collect(4)
        assert np.array_equal(c4, oc[150:180]) and np.array_equal(f4, of[150:180])
        f2, c2 = ctx.collect(2)
        assert np.array_equal(c2, np.concatenate([oc[40:60], oc[60:74]]))
        # many generations: the context never has zero open videos, the log (64 frames) is recycled ~30 times
        ctx.video_close(4)
        ctx.video_close(2)
        ctx.video_open(100, spec.width, spec.height)
        for g in range(60):
            ctx.video_open(101 + g, spec.width, spec.height)
            a = (g * 7) % 150
            feed(ctx, 101 + g, a, a + 29)  # two generations are alive at a time: 58 of the 64 log frames
            fg, cg = ctx.collect(101 + g)
            assert np.array_equal(cg, oc[a : a + 29]) and np.array_equal(fg, of[a : a + 29]), g
            ctx.video_close(100 + g)  # the previous generation goes, this one stays open
