"""GPU: mscan_submit_device — frames whose records already lie in device memory join a video's frame log and are
scanned in place (no host staging). Results must equal the host-fed ones and the oracle, for both record layouts, with
a producer stream, mixed with host submits in one video, and through collect / segments / append_from."""
import ctypes as C

import numpy as np
import pytest

import kats
import motionscan as ms
import oracle_lib as orc
from test_gpu_parity import cfg_for, oracle_tail

pytestmark = pytest.mark.gpu


def upload(ctx, arr):
    d = ctx.dev_alloc(arr.nbytes + 256)
    ctx.h2d(d, arr)
    return d


@pytest.mark.parametrize("packed", [False, True])
def test_device_submit_equals_host_submit_and_oracle(packed):
    p = kats.env_params()
    spec = ms.synth_preset(0, 77)
    n = 300
    cnt, off, recs, pts = ms.synth_host(spec, 0, n)
    cfg = cfg_for(p, spec.width, spec.height)
    of, oc = orc.scan_frames(cfg, recs, off, threads=4)
    src = ms.pack_records(recs) if packed else recs
    with ms.Context(0, p, 1 << 16, 8 << 20) as ctx:
        d = upload(ctx, src)
        ctx.video_open(1, spec.width, spec.height)
        ctx.video_open(2, spec.width, spec.height)
        # video 1: three device submits (the second and third start in the middle of the buffer: record offsets)
        cuts = [0, 100, 101, n]
        for a, b in zip(cuts[:-1], cuts[1:]):
            stride = 8 if packed else 40
            # sub-ranges are addressed by passing the whole buffer's base and skipping frames with a 0-frame prefix is
            # not possible through the ABI: a piece must start at a 16-byte aligned address
            base = d + int(off[a]) * stride
            if base % 16:
                # unaligned piece start: hand the frames over from a copy that starts aligned
                piece = np.ascontiguousarray(src[int(off[a]) : int(off[b])])
                base = upload(ctx, piece)
            first = ctx.submit_device(1, pts[a:b], cnt[a:b], base, packed=packed)
            assert first == a
        ctx.submit(2, pts, cnt, recs)  # video 2: the same frames host-fed
        f1, c1 = ctx.collect(1)
        f2, c2 = ctx.collect(2)
        assert np.array_equal(f1, of) and np.array_equal(c1, oc)
        assert np.array_equal(f2, of) and np.array_equal(c2, oc)
        s1, r1 = ctx.motion_segments(1, n / spec.fps)
        s2, r2 = ctx.motion_segments(2, n / spec.fps)
        osegs, ores = oracle_tail(p, pts, of, n / spec.fps)
        assert s1.tobytes() == s2.tobytes() == osegs.tobytes()
        assert r1.decision == r2.decision == ores.decision
        st = ctx.stats()
        assert st.records_projected == int(off[-1])  # only the host-fed video went through the staging pass


def test_device_submit_after_producer_stream_and_mixed_with_host_frames():
    """The records are produced on the caller's stream (here: the library's own device-side projection kernel) and
    handed over with that stream as `ready_stream`; host-fed frames of the same video come before and after."""
    import torch

    p = kats.env_params()
    spec = ms.synth_preset(3, 5)
    n = 240
    cnt, off, recs, pts = ms.synth_host(spec, 0, n)
    cfg = cfg_for(p, spec.width, spec.height)
    of, oc = orc.scan_frames(cfg, recs, off, threads=4)
    a, b = 80, 160
    while int(off[a]) % 2:  # an even record index keeps the 8-byte-record piece 16-byte aligned
        a += 1
    stream = torch.cuda.Stream()
    with ms.Context(0, p, 1 << 16, 4 << 20) as ctx:
        d_native = upload(ctx, recs)
        d_packed = ctx.dev_alloc(8 * len(recs) + 256)
        ctx.video_open(7, spec.width, spec.height)
        ctx.submit(7, pts[:a], cnt[:a], recs[: int(off[a])])
        ctx.pack_records_device(d_native, len(recs), d_packed, stream.cuda_stream)  # producer work on the caller's stream
        first = ctx.submit_device(7, pts[a:b], cnt[a:b], d_packed + 8 * int(off[a]), packed=True, ready_stream=stream.cuda_stream)
        assert first == a
        ctx.submit(7, pts[b:], cnt[b:], recs[int(off[b]) :])
        fl, cn = ctx.collect(7)
        assert np.array_equal(fl, of) and np.array_equal(cn, oc)
        segs, res = ctx.segments(7, n / spec.fps)
        osegs, ores = oracle_tail(p, pts, of, n / spec.fps)
        assert res.decision == ores.decision and res.n_motion_frames == ores.n_motion_frames
        # stitched into another video of the context: the adopted frames keep their results
        ctx.video_open(8, spec.width, spec.height)
        ctx.video_append_from(8, ctx, 7)
        f8, c8 = ctx.collect(8)
        assert np.array_equal(f8, of) and np.array_equal(c8, oc)
        ctx.video_close(7)
        ctx.video_close(8)


def test_device_submit_argument_checks():
    p = kats.env_params()
    with ms.Context(0, p, 1 << 12, 1 << 20) as ctx:
        d = ctx.dev_alloc(4096)
        one = np.array([1], np.uint32)
        t = np.array([0.0])
        with pytest.raises(ms.MscanError) as e:
            ctx.submit_device(5, t, one, d)  # video not open
        assert e.value.code == ms.ERR_INVALID
        ctx.video_open(5, 1920, 1080)
        with pytest.raises(ms.MscanError) as e:
            ctx.submit_device(5, t, one, d + 8)  # not 16-byte aligned
        assert e.value.code == ms.ERR_INVALID
        assert ctx.submit_device(5, t[:0], one[:0], d) == 0  # empty call: index of the next frame
        ctx.submit_device(5, t, np.array([0], np.uint32), 0)  # a frame without side data needs no records
        fl, cn = ctx.collect(5)
        assert list(fl) == [0] and list(cn) == [0]
