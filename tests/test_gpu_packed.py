"""GPU: every way records can reach K-A gives the oracle's answer bit for bit.

Feed modes (include/motionscan.h, mscan_set_staging_mode / mscan_submit_packed):
  native-inplace   native 40-byte records in pinned memory, DMA'd in place        → K-A<native>
  native-staged    native records in pageable memory, memcpy'd to staging (mode NATIVE) → K-A<native>
  projected        native records in pageable memory, staging pass projects to 8 B (mode AUTO, default) → K-A<packed>
  projected-pinned native records in pinned memory, projected anyway (mode PACK)   → K-A<packed>
  packed-pinned    caller-projected mscan_mv8 in pinned memory, DMA'd in place     → K-A<packed>
  packed-pageable  caller-projected mscan_mv8 in pageable memory                   → K-A<packed>
  elided           native records in pageable memory, sent in the static-elided form (mode ELIDE) → K-A<mvz>
  elided-pinned    native records in pinned memory, same                           → K-A<mvz>
  compact          native records in pageable memory, only the moving ones sent (mode COMPACT) → K-A<packed>
  compact-pinned   native records in pinned memory, same                           → K-A<packed>
"""
import numpy as np
import pytest

import kats
import motionscan as ms
import oracle_lib as orc

pytestmark = pytest.mark.gpu

MODES = ["native-inplace", "native-staged", "projected", "projected-pinned", "packed-pinned", "packed-pageable", "elided", "elided-pinned", "compact",
         "compact-pinned"]


def cfg_for(p, w, h):
    gw, gh, m = orc.geometry(w, h, p.block_size, p.block_shift, p.vertical_mask)
    return orc.make_cfg(p, gw, gh, m)


class Feeder:
    """Submits (pts, cnt, recs) to a video in one of MODES; owns the pinned copies until close()."""

    def __init__(self, ctx, mode):
        self.ctx, self.mode, self.pinned = ctx, mode, []
        ctx.set_staging_mode({"native-staged": ms.STAGING_NATIVE, "projected-pinned": ms.STAGING_PACK, "elided": ms.STAGING_ELIDE,
                              "elided-pinned": ms.STAGING_ELIDE, "compact": ms.STAGING_COMPACT,
                              "compact-pinned": ms.STAGING_COMPACT}.get(mode, ms.STAGING_AUTO))

    def _pin(self, a):
        h = self.ctx.pinned_array(max(len(a), 1), a.dtype)[: len(a)]
        h[:] = a
        self.pinned.append(h)
        return h

    def submit(self, vid, pts, cnt, recs):
        if recs is None:
            recs = np.zeros(0, ms.MV_DTYPE)
        m = self.mode
        if m in ("native-inplace", "projected-pinned", "elided-pinned", "compact-pinned"):
            return self.ctx.submit(vid, pts, cnt, self._pin(recs))
        if m in ("native-staged", "projected", "elided", "compact"):
            return self.ctx.submit(vid, pts, cnt, recs)
        r8 = ms.pack_records(recs)
        return self.ctx.submit_packed(vid, pts, cnt, self._pin(r8) if m == "packed-pinned" else r8)

    def close(self):
        self.ctx.host_fence()
        for h in self.pinned:  # views starting at the allocation's first byte
            self.ctx.host_free(h.ctypes.data)
        self.pinned = []


@pytest.mark.parametrize("mode", MODES)
def test_frame_kats_every_feed_mode(mode):
    K = kats.frame_kats()
    groups = {}
    for name, (p, recs, flag, count) in K.items():
        key = (p.mv_threshold_sq, p.vectors_needed, p.clusters_needed)
        groups.setdefault(key, (p, []))[1].append((name, recs, flag, count))
    for p, items in groups.values():
        with ms.Context(0, p) as ctx:
            fd = Feeder(ctx, mode)
            frames = [r for _, r, _, _ in items]
            cnt = np.array([0 if f is None else len(f) for f in frames], dtype=np.uint32)
            recs = kats.cat(*[f for f in frames if f is not None and len(f)])
            ctx.video_open(1, kats.W, kats.H)
            fd.submit(1, np.arange(len(frames)) / 30.0, cnt, recs)
            flags, counts = ctx.collect(1)
            fd.close()
        for i, (name, _, flag, count) in enumerate(items):
            assert flags[i] == flag, (mode, name)
            assert counts[i] == count, (mode, name)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("seed", range(3))
def test_random_ragged_frames_every_feed_mode(mode, seed):
    """Ragged frames whose sizes sit on and around the ring-stage boundaries of both layouts (512 native /
    2560 packed records per stage), starting at odd record indices (the 8-byte phase inside a 16-byte copy)."""
    from test_oracle_kats import random_frame

    rng = np.random.default_rng(4200 + seed)
    w, h = [(1920, 1080), (3840, 2160), (640, 360)][seed]
    p = kats.env_params() if seed != 1 else kats.code_defaults()
    p.vectors_needed = int(rng.integers(1, 5))
    p.clusters_needed = int(rng.integers(1, 4))
    sizes = [1, 0, 511, 512, 513, 3, 2559, 2560, 2561, 0, 5119, 5120, 5121, 7, 7679, 7680, 7681, 1023, 20479, 2, 10241]
    sizes += [int(x) for x in rng.integers(0, 9000, 20)]
    frames = [random_frame(rng, n, w, h, int(rng.integers(1, 6))) if n else None for n in sizes]
    cnt = np.array(sizes, dtype=np.uint32)
    recs = kats.cat(*[f for f in frames if f is not None])
    pts = np.arange(len(sizes)) / 25.0
    cfg = cfg_for(p, w, h)
    with ms.Context(0, p) as ctx:
        fd = Feeder(ctx, mode)
        ctx.video_open(5, w, h)
        fd.submit(5, pts, cnt, recs)
        flags, counts = ctx.collect(5)
        fd.close()
    for i, f in enumerate(frames):
        assert counts[i] == orc.full_count(cfg, f), (mode, i, sizes[i])
        assert flags[i] == orc.check_frame(cfg, f), (mode, i, sizes[i])


@pytest.mark.parametrize("mode", MODES)
def test_synthetic_clip_every_feed_mode(mode):
    """configs[0]-style 1080p clip through each feed mode, tiny slabs so frames span many launches;
    flags, full counts, segments and decision bit-exact."""
    spec = ms.synth_preset(0, 21)
    n = 420
    cnt, off, recs, pts = ms.synth_host(spec, 0, n)
    p = kats.env_params()
    of, oc = orc.scan_frames(cfg_for(p, spec.width, spec.height), recs, off, threads=8)
    osegs, ores = orc.video_tail(pts, of, n / spec.fps, p.max_gap_sec, p.padding_sec, p.min_savings_pct)
    with ms.Context(0, p, 0, 4 << 20) as ctx:
        fd = Feeder(ctx, mode)
        ctx.video_open(2, spec.width, spec.height)
        for a in range(0, n, 53):
            b = min(n, a + 53)
            fd.submit(2, pts[a:b], cnt[a:b], recs[int(off[a]) : int(off[b])])
        flags, counts = ctx.collect(2)
        segs, res = ctx.motion_segments(2, n / spec.fps)
        st = ctx.stats()
        fd.close()
    assert np.array_equal(counts, oc) and np.array_equal(flags, of)
    assert res.decision == ores.decision and segs.tobytes() == osegs.tobytes()
    assert np.float64(res.saved_pct).tobytes() == np.float64(ores.saved_pct).tobytes()
    n_rec = int(off[-1])
    if mode.startswith("native"):
        assert st.h2d_bytes >= 40 * n_rec and st.records_projected == 0
    elif mode.startswith("elided"):
        assert 4 * n_rec <= st.h2d_bytes < 8 * n_rec  # static records travel as 4 bytes + a mask bit
        assert st.records_projected == n_rec == st.records_elided
    elif mode.startswith("compact"):
        r8 = ms.pack_records(recs)
        n_moving = int(((r8["src_x"] != r8["dst_x"]) | (r8["src_y"] != r8["dst_y"])).sum())
        assert st.records_projected == n_rec == st.records_elided and st.elided_bytes == 8 * n_moving
        assert 8 * n_moving <= st.h2d_bytes < 8 * n_moving + 64 * n  # static records do not travel at all
        assert st.records_scanned == n_moving
    else:
        assert 8 * n_rec <= st.h2d_bytes < 9 * n_rec  # 8 B/record + per-frame metadata cross PCIe
        assert st.records_projected == (n_rec if mode.startswith("projected") else 0)


def test_formats_mixed_inside_one_video():
    """Chunk workers of one video may feed different formats; a slab never mixes them."""
    spec = ms.synth_preset(3, 9)
    n = 240
    cnt, off, recs, pts = ms.synth_host(spec, 0, n)
    p = kats.env_params()
    of, oc = orc.scan_frames(cfg_for(p, spec.width, spec.height), recs, off, threads=8)
    with ms.Context(0, p, 0, 16 << 20) as ctx:
        hp = ctx.pinned_array(len(recs), ms.MV_DTYPE)
        hp[:] = recs
        r8 = ms.pack_records(recs)
        ctx.video_open(1, spec.width, spec.height)
        for k, a in enumerate(range(0, n, 20)):
            b = a + 20
            r0, r1 = int(off[a]), int(off[b])
            if k % 3 == 0:
                ctx.submit(1, pts[a:b], cnt[a:b], hp[r0:r1])       # native, in place
            elif k % 3 == 1:
                ctx.submit(1, pts[a:b], cnt[a:b], recs[r0:r1])     # projected by the staging pass
            else:
                ctx.submit_packed(1, pts[a:b], cnt[a:b], r8[r0:r1])  # caller-projected
        flags, counts = ctx.collect(1)
        st = ctx.stats()
        ctx.host_free(hp.ctypes.data)
    assert np.array_equal(counts, oc) and np.array_equal(flags, of)
    assert st.scan_launches >= 8  # every format switch closes a segment (one K-A launch each)


@pytest.mark.parametrize("threads", [1, 3, 0])
def test_large_submit_uses_the_projection_pool(threads):
    """One submit of > 256 Ki records is projected by the pool (any size of it) with identical results."""
    spec = ms.synth_preset(2, 31)  # 4K dense: 129 600 records per P-frame
    n = 14
    cnt, off, recs, pts = ms.synth_host(spec, 0, n)
    assert int(off[-1]) > 1 << 20
    p = kats.env_params()
    of, oc = orc.scan_frames(cfg_for(p, spec.width, spec.height), recs, off, threads=8)
    with ms.Context(0, p) as ctx:
        ctx.set_pack_threads(threads)
        ctx.video_open(1, spec.width, spec.height)
        ctx.submit(1, pts, cnt, recs)
        ctx.submit(1, pts + 1.0, cnt, recs)  # pool reused for a second job
        flags, counts = ctx.collect(1)
        st = ctx.stats()
    assert np.array_equal(counts, np.concatenate([oc, oc])) and np.array_equal(flags, np.concatenate([of, of]))
    assert st.records_projected == 2 * int(off[-1]) and st.project_ms > 0


def test_scan_device_packed_matches_native():
    """Device-resident: K-A<packed> on the projection of a stream == K-A<native> on the stream == oracle."""
    spec = ms.synth_preset(1, 3)
    n = 900
    cnt, off, recs, pts = ms.synth_host(spec, 0, n)
    r8 = ms.pack_records(recs)
    p = kats.env_params()
    of, oc = orc.scan_frames(cfg_for(p, spec.width, spec.height), recs, off, threads=8)
    with ms.Context(0, p) as ctx:
        d_recs = ctx.dev_alloc(recs.nbytes + 256)
        d_r8 = ctx.dev_alloc(r8.nbytes + 256)
        d_off = ctx.dev_alloc(off.nbytes)
        d_flags, d_counts = ctx.dev_alloc(n), ctx.dev_alloc(4 * n)
        ctx.h2d(d_recs, recs)
        ctx.pack_records_device(d_recs, len(recs), d_r8)  # device projection == host projection
        ctx.sync()
        back = np.zeros(len(recs), ms.MV8_DTYPE)
        ctx.d2h(back, d_r8)
        assert back.tobytes() == r8.tobytes()
        ctx.h2d(d_off, off)
        geom = ms.geometry_from_dims(p, spec.width, spec.height)
        out = []
        for fn, d in ((ctx.scan_device, d_recs), (ctx.scan_device_packed, d_r8)):
            fn(d, d_off, None, [geom], n, d_flags, d_counts)
            ctx.sync()
            f, c = np.zeros(n, np.uint8), np.zeros(n, np.uint32)
            ctx.d2h(f, d_flags)
            ctx.d2h(c, d_counts)
            out.append((f, c))
        for d in (d_recs, d_r8, d_off, d_flags, d_counts):
            ctx.dev_free(d)
    for f, c in out:
        assert np.array_equal(c, oc) and np.array_equal(f, of)



def test_many_tiny_segments_across_the_slab_ring():
    """300 single-frame submits that alternate record formats: every submit closes a segment (one K-A launch each),
    slabs rotate many times, launches pile up on three streams — the log must still be the oracle's, in order."""
    spec = ms.synth_preset(3, 55)
    n = 300
    cnt, off, recs, pts = ms.synth_host(spec, 0, n)
    p = kats.env_params()
    of, oc = orc.scan_frames(cfg_for(p, spec.width, spec.height), recs, off, threads=8)
    with ms.Context(0, p, 0, 2 << 20) as ctx:  # 2 MiB slabs: a handful of frames each
        hp = ctx.pinned_array(len(recs), ms.MV_DTYPE)
        hp[:] = recs
        r8 = ms.pack_records(recs)
        ctx.video_open(1, spec.width, spec.height)
        for i in range(n):
            r0, r1 = int(off[i]), int(off[i + 1])
            if i % 2:
                ctx.submit(1, pts[i : i + 1], cnt[i : i + 1], hp[r0:r1])
            else:
                ctx.submit_packed(1, pts[i : i + 1], cnt[i : i + 1], r8[r0:r1])
        flags, counts = ctx.collect(1)
        segs, res = ctx.motion_segments(1, n / spec.fps)
        st = ctx.stats()
        ctx.host_free(hp.ctypes.data)
    osegs, ores = orc.video_tail(pts, of, n / spec.fps, p.max_gap_sec, p.padding_sec, p.min_savings_pct)
    assert np.array_equal(counts, oc) and np.array_equal(flags, of)
    assert segs.tobytes() == osegs.tobytes() and res.decision == ores.decision
    assert st.scan_launches >= 250  # I-frames carry no records and need no launch of their own


def test_elided_transport_sends_fewer_bytes_and_counts_them():
    """MSCAN_STAGING_ELIDE on a CCTV-style clip: identical results, ~4.3 B/record on the wire instead of 8, and the
    encoder's output decodes (numpy reference decoder) to exactly the projected records."""
    p = kats.env_params()
    spec = ms.synth_preset(1, 2)
    n = 400
    cnt, off, recs, pts = ms.synth_host(spec, 0, n)
    of, oc = orc.scan_frames(cfg_for(p, spec.width, spec.height), recs, off, threads=4)
    with ms.Context(0, p, 1 << 16, 8 << 20) as ctx:
        ctx.set_staging_mode(ms.STAGING_ELIDE)
        ctx.video_open(1, spec.width, spec.height)
        # per-frame submits (a decode thread's pattern), then one multi-frame submit that spans several pieces and slabs
        for f in range(100):
            ctx.submit(1, pts[f : f + 1], cnt[f : f + 1], recs[int(off[f]) : int(off[f + 1])])
        ctx.submit(1, pts[100:], cnt[100:], recs[int(off[100]) :])
        flags, counts = ctx.collect(1)
        st = ctx.stats()
    assert np.array_equal(flags, of) and np.array_equal(counts, oc) and of.any()
    assert st.records_elided == int(off[-1]) == st.records_projected
    assert 4.0 < st.elided_bytes / st.records_elided < 5.0
    fr = np.ascontiguousarray(recs[int(off[1]) : int(off[2])])
    enc, te = ms.elide_records(fr)
    assert ms.unelide_records(enc, te, len(fr)).tobytes() == ms.pack_records(fr).tobytes()


def test_elided_mode_falls_back_to_mv8_for_cluster_sized_grids():
    """The cluster kernel (8K and larger grids) reads native and mv8 records only: ELIDE then projects to mv8."""
    from test_oracle_kats import random_frame

    p = kats.env_params(vectors_needed=2)
    w, h = 7680, 4320
    rng = np.random.default_rng(12)
    frames = [random_frame(rng, 20000, w, h, 6) for _ in range(4)]
    cnt = np.array([len(f) for f in frames], np.uint32)
    cfg = cfg_for(p, w, h)
    with ms.Context(0, p) as ctx:
        ctx.set_staging_mode(ms.STAGING_ELIDE)
        ctx.video_open(1, w, h)
        ctx.submit(1, np.arange(4) / 30.0, cnt, kats.cat(*frames))
        flags, counts = ctx.collect(1)
        st = ctx.stats()
    assert list(counts) == [orc.full_count(cfg, f) for f in frames]
    assert st.records_elided == 0 and st.records_projected == int(cnt.sum())


@pytest.mark.parametrize("pinned", [True, False])
def test_submit_elided_takes_the_callers_own_encoding(pinned):
    """mscan_submit_elided: frames the caller encoded itself (mscan_elide_records per frame) — DMA'd in place from pinned
    memory or copied from pageable memory — give the oracle's results; mixed with other layouts in one video."""
    p = kats.env_params()
    spec = ms.synth_preset(1, 6)
    n = 300
    cnt, off, recs, pts = ms.synth_host(spec, 0, n)
    of, oc = orc.scan_frames(cfg_for(p, spec.width, spec.height), recs, off, threads=4)
    with ms.Context(0, p, 1 << 16, 4 << 20) as ctx:
        enc, enc_off, te = ms.elide_frames(recs, off)
        enc_pageable = np.array(enc)  # (the pinned copy below is freed before the argument checks)
        buf = None
        if pinned:
            buf = ctx.pinned_array(len(enc) + 16, np.uint8)
            buf[: len(enc)] = enc
            enc = buf[: len(enc)]
        ctx.video_open(1, spec.width, spec.height)
        # frames [0,100) natively, [100,250) pre-encoded in two calls, the rest projected by the library
        ctx.set_staging_mode(ms.STAGING_NATIVE)
        ctx.submit(1, pts[:100], cnt[:100], recs[: int(off[100])])
        tiles = np.concatenate([[0], np.cumsum((cnt.astype(np.int64) + 1023) // 1024)])
        for a, b in ((100, 180), (180, 250)):
            first = ctx.submit_elided(1, pts[a:b], cnt[a:b], enc, enc_off[a : b + 1], te[int(tiles[a]) : int(tiles[b])])
            assert first == a
        ctx.set_staging_mode(ms.STAGING_PACK)
        ctx.submit(1, pts[250:], cnt[250:], recs[int(off[250]) :])
        flags, counts = ctx.collect(1)
        st = ctx.stats()
        if pinned:
            ctx.host_fence()
            ctx.host_free(buf.ctypes.data)
    assert np.array_equal(flags, of) and np.array_equal(counts, oc)
    assert st.records_elided == int(off[250] - off[100])
    # argument checks
    with ms.Context(0, p, 1 << 12, 1 << 20) as ctx:
        ctx.video_open(1, spec.width, spec.height)
        bad_off = enc_off[:3].copy()
        bad_off[1] += 8
        with pytest.raises(ms.MscanError) as e:
            ctx.submit_elided(1, pts[:2], cnt[:2], enc_pageable, bad_off, te)
        assert e.value.code == ms.ERR_INVALID
        with pytest.raises(ms.MscanError) as e:
            ctx.submit_elided(9, pts[:2], cnt[:2], enc_pageable, enc_off[:3], te)
        assert e.value.code == ms.ERR_INVALID


def test_compact_transport_sends_only_moving_records():
    """MSCAN_STAGING_COMPACT on a CCTV-style clip: identical flags and counts with well under 1 B/record on the wire;
    per-frame submits, then a multi-frame submit that spans several pieces and slabs, keep submission order."""
    p = kats.env_params()
    spec = ms.synth_preset(1, 2)
    n = 400
    cnt, off, recs, pts = ms.synth_host(spec, 0, n)
    of, oc = orc.scan_frames(cfg_for(p, spec.width, spec.height), recs, off, threads=4)
    with ms.Context(0, p, 1 << 16, 1 << 20) as ctx:
        ctx.set_staging_mode(ms.STAGING_COMPACT)
        ctx.video_open(1, spec.width, spec.height)
        for f in range(100):
            assert ctx.submit(1, pts[f : f + 1], cnt[f : f + 1], recs[int(off[f]) : int(off[f + 1])]) == f
        assert ctx.submit(1, pts[100:], cnt[100:], recs[int(off[100]) :]) == 100
        flags, counts = ctx.collect(1)
        st = ctx.stats()
    assert np.array_equal(flags, of) and np.array_equal(counts, oc) and of.any()
    assert st.records_elided == int(off[-1]) == st.records_projected
    assert st.elided_bytes / st.records_elided < 1.5


@pytest.mark.parametrize("thr", [0.0, -1.0, float("nan"), 0.25, 1e300])
def test_compact_mode_only_drops_what_cannot_vote(thr):
    """Static records vote when MV_THRESHOLD_SQ <= 0 (or NaN): COMPACT and AUTO must then send them (static-elided
    form); with any positive threshold — also one no int32 magnitude reaches — the compaction is exact."""
    p = kats.env_params(vectors_needed=2)
    p.mv_threshold_sq = thr
    spec = ms.synth_preset(1, 8)
    n = 60
    cnt, off, recs, pts = ms.synth_host(spec, 0, n)
    cfg = cfg_for(p, spec.width, spec.height)
    of, oc = orc.scan_frames(cfg, recs, off, threads=4)
    for mode in (ms.STAGING_COMPACT, ms.STAGING_AUTO):
        with ms.Context(0, p, 1 << 12, 8 << 20) as ctx:
            ctx.set_staging_mode(mode)
            ctx.video_open(1, spec.width, spec.height)
            for f in range(n):
                ctx.submit(1, pts[f : f + 1], cnt[f : f + 1], recs[int(off[f]) : int(off[f + 1])])
            flags, counts = ctx.collect(1)
            st = ctx.stats()
        assert np.array_equal(flags, of) and np.array_equal(counts, oc), (thr, mode)
        per_rec = st.elided_bytes / st.records_elided
        assert (per_rec < 1.5) if thr > 0 else (per_rec >= 4.0), (thr, mode, per_rec)
    if not thr > 0:
        assert of.all() or oc.max() > 0  # static macroblocks really did vote here


def test_compact_mode_feeds_the_cluster_kernel_too():
    """Compacted records are plain mscan_mv8: grids beyond one CTA's shared memory (8K) take them like any packed submit."""
    from test_oracle_kats import random_frame

    p = kats.env_params(vectors_needed=2)
    w, h = 7680, 4320
    rng = np.random.default_rng(13)
    frames = [random_frame(rng, 20000, w, h, 6) for _ in range(4)]
    for f in frames:  # make two thirds of the records static
        keep = np.arange(len(f)) % 3 != 0
        f["src_x"][keep], f["src_y"][keep] = f["dst_x"][keep], f["dst_y"][keep]
    cnt = np.array([len(f) for f in frames], np.uint32)
    cfg = cfg_for(p, w, h)
    with ms.Context(0, p) as ctx:
        ctx.set_staging_mode(ms.STAGING_COMPACT)
        ctx.video_open(1, w, h)
        ctx.submit(1, np.arange(4) / 30.0, cnt, kats.cat(*frames))
        flags, counts = ctx.collect(1)
        st = ctx.stats()
    assert list(counts) == [orc.full_count(cfg, f) for f in frames]
    assert st.records_elided == int(cnt.sum()) and 0 < st.records_scanned < int(cnt.sum()) // 2


def test_concurrent_multi_piece_compact_submits_keep_each_calls_frames_together():
    """Two threads submit large COMPACT calls (several pieces each) to ONE video at the same time: every call's frames
    occupy the contiguous index range that starts at its first_frame_out, in the call's own order."""
    import threading

    p = kats.env_params()
    spec = ms.synth_preset(1, 17)
    n = 240  # ~2.4 M records per call: about ten pieces of 256 Ki records
    cnt, off, recs, pts = ms.synth_host(spec, 0, n)
    of, oc = orc.scan_frames(cfg_for(p, spec.width, spec.height), recs, off, threads=8)
    assert int(off[-1]) > 4 * (1 << 18)
    with ms.Context(0, p, 1 << 14, 2 << 20) as ctx:
        ctx.set_staging_mode(ms.STAGING_COMPACT)
        ctx.video_open(1, spec.width, spec.height)
        firsts = {}

        def work(k):
            for rep in range(2):
                firsts[(k, rep)] = ctx.submit(1, pts + 100.0 * (2 * k + rep), cnt, recs)

        th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        flags, counts = ctx.collect(1)
    assert sorted(firsts.values()) == [0, n, 2 * n, 3 * n]
    for first in firsts.values():
        assert np.array_equal(flags[first : first + n], of) and np.array_equal(counts[first : first + n], oc)
