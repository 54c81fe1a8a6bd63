"""GPU: one video split over several contexts/GPUs and stitched with mscan_video_append_from,
and the host mirror's MOTION_TRIM_SPLIT_GPUS / WATCH_MODE on top of it. The stitched result must equal the
oracle's for the whole video, whatever the order in which chunks were scanned or appended."""
import ctypes as C
import os
import subprocess
import tempfile
import threading
import time
from pathlib import Path

import numpy as np
import pytest

import kats
import motionscan as ms
import oracle_lib as orc
from test_host_cli import BIN, parse, run_cli, write_case
from test_ref_golden import cases, expected

pytestmark = pytest.mark.gpu


def n_devices():
    n = C.c_int()
    ms.lib().mscan_device_count(C.byref(n))
    return n.value


def cfg_for(p, w, h):
    gw, gh, m = orc.geometry(w, h, p.block_size, p.block_shift, p.vertical_mask)
    return orc.make_cfg(p, gw, gh, m)


@pytest.mark.parametrize("n_ctx", [2, 3])
def test_split_video_stitched_equals_oracle(n_ctx):
    spec = ms.synth_preset(0, 17)
    n = 900
    cnt, off, recs, pts = ms.synth_host(spec, 0, n)
    p = kats.env_params()
    of, oc = orc.scan_frames(cfg_for(p, spec.width, spec.height), recs, off, threads=8)
    osegs, ores = orc.video_tail(pts, of, n / spec.fps, p.max_gap_sec, p.padding_sec, p.min_savings_pct)
    nd = n_devices()
    ctxs = [ms.Context(k % nd, p, 1 << 16, 32 << 20) for k in range(n_ctx)]  # real peers when the box has them
    try:
        for c in ctxs:
            c.video_open(7, spec.width, spec.height)
        chunks = [(a, min(n, a + 75)) for a in range(0, n, 75)]
        order = list(range(len(chunks)))
        np.random.default_rng(3).shuffle(order)  # chunk workers finish in any order
        for j in order:
            a, b = chunks[j]
            ctxs[j % n_ctx].submit(7, pts[a:b], cnt[a:b], recs[int(off[a]) : int(off[b])])
        for c in ctxs[:0:-1]:  # append in reverse, too
            ctxs[0].video_append_from(7, c, 7)
        segs, res = ctxs[0].motion_segments(7, n / spec.fps)
        flags, counts = ctxs[0].collect(7)
        st = ctxs[0].stats()
    finally:
        for c in ctxs:
            c.close()
    assert res.decision == ores.decision and res.n_motion_frames == ores.n_motion_frames
    assert segs.tobytes() == osegs.tobytes()
    assert np.float64(res.saved_pct).tobytes() == np.float64(ores.saved_pct).tobytes()
    assert len(flags) == n and int(flags.sum()) == int(of.sum()) and int(counts.sum()) == int(oc.sum())
    assert st.peer_bytes == 13 * sum(b - a for j, (a, b) in enumerate(chunks) if j % n_ctx)


def test_append_within_one_context_and_errors():
    p = kats.env_params()
    active = kats.cat(kats.cell(10, 10), kats.cell(11, 10))
    with ms.Context(0, p) as ctx:
        for vid, ts in ((1, [1.0, 2.0]), (2, [10.0])):
            ctx.video_open(vid, kats.W, kats.H)
            ctx.submit(vid, np.array(ts), np.full(len(ts), len(active), np.uint32), kats.cat(*[active] * len(ts)))
        ctx.video_append_from(1, ctx, 2)
        segs, res = ctx.motion_segments(1, 60.0)
        assert [(s["start"], s["end"]) for s in segs] == [(0.5, 2.5), (9.5, 10.5)]  # KAT S1
        f2, _ = ctx.collect(2)
        assert len(f2) == 1  # the source video is untouched
        for bad in ((1, 1), (1, 99), (99, 2)):
            with pytest.raises(ms.MscanError) as e:
                ctx.video_append_from(bad[0], ctx, bad[1])
            assert e.value.code == ms.ERR_INVALID


def test_cli_split_one_video_over_gpus():
    """MOTION_TRIM_SPLIT_GPUS: chunk workers bound to different GPUs, stitched on GPU 0 — same job as the reference's."""
    if n_devices() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    for name in ("clip60s_1080p_config0", "skip_tfps7_chunk2p5", "rand_params_5"):
        c, e = cases()[name], expected(name)
        with tempfile.TemporaryDirectory() as d:
            path, out = Path(d) / "in.mvs", Path(d) / "out.mp4"
            write_case(c, path)
            r = run_cli(["--print-segments", str(path), str(out)], c.params, chunk_sec=c.chunk_sec, threads=max(c.threads or 0, 4),
                        target_fps=c.target_fps, extra_env={"MOTION_TRIM_SPLIT_GPUS": "2"})
            assert r.returncode == 0, r.stdout + r.stderr
            assert "Splitting the video over 2 GPUs" in r.stdout
            res, segs = parse(r.stdout)
            assert int(res["decision"]) == e["decision"]
            assert np.array(segs).reshape(-1, 2).tobytes() == e["segs"].tobytes()


def test_cli_batch_identical_on_one_and_two_gpus():
    """a 16-clip batch sharded by video over 2 GPUs (PARALLEL_STREAMS stream threads per GPU,
    one context per GPU, no exchange) gives, per video, exactly what one GPU gives — and what the oracle gives."""
    if n_devices() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import motionscan as ms
    import oracle_lib as orc
    from motionscan import mvs_io

    p = cases()["batchclip_seed100"].params
    n_clips, n_frames = 16, 450
    want = {}
    with tempfile.TemporaryDirectory() as d:
        ind = Path(d) / "in"
        ind.mkdir()
        for k in range(n_clips):
            spec = ms.synth_preset(3, 500 + k)
            cnt, off, recs, pts = ms.synth_host(spec, 0, n_frames)
            mvs_io.write_mvs(ind / f"clip{k:02d}.mvs", spec.width, spec.height, 30, 1, np.arange(n_frames), cnt, recs)
            gw, gh, m = orc.geometry(spec.width, spec.height, p.block_size, p.block_shift, p.vertical_mask)
            of, _ = orc.scan_frames(orc.make_cfg(p, gw, gh, m), recs, off, threads=4)
            pts = np.arange(n_frames, dtype=np.float64) * (1.0 / 30.0)  # ticks * time_base, as the host computes it (:361)
            _, ores = orc.video_tail(pts, of, n_frames / 30.0, p.max_gap_sec, p.padding_sec, p.min_savings_pct)
            want[f"clip{k:02d}"] = (ores.decision, ores.saved_pct)
        runs = {}
        for gpus in (1, 2):
            outd = Path(d) / f"out{gpus}"
            r = run_cli(["--print-segments", str(ind), str(outd)], p, chunk_sec=5.0,
                        extra_env={"PARALLEL_STREAMS": "2", "MOTION_TRIM_GPUS": str(gpus)})
            assert r.returncode == 0, r.stdout + r.stderr
            assert f"{gpus} GPU(s)" in r.stdout
            got = {}
            for line in r.stdout.splitlines():
                if line.startswith("RESULT "):
                    parts = line.split()
                    kv = dict(x.split("=") for x in parts[2:])
                    got[parts[1][:-4]] = (kv["rc"], kv["decision"], kv["segments"], kv["saved_pct"], kv["gpu"])
            runs[gpus] = got
            lists = {f.name: f.read_text().replace(str(ind.resolve()), "IN") for f in sorted(outd.glob("*.concat.txt"))}
            runs[(gpus, "lists")] = lists
        assert sorted(runs[1]) == sorted(runs[2]) == sorted(want)
        assert {v[4] for v in runs[2].values()} == {"0", "1"}, "both GPUs must have scanned videos"
        for name in want:
            assert runs[1][name][:4] == runs[2][name][:4], name                      # identical per-video results
            assert int(runs[1][name][1]) == want[name][0]                            # and the oracle's decision
            if want[name][0]:
                assert float.fromhex(runs[1][name][3]) == want[name][1], name
        assert runs[(1, "lists")] == runs[(2, "lists")]                              # identical cut lists (segments)


def test_cli_watch_mode_picks_up_new_files():
    """WATCH_MODE=1 (batch_processor.cpp:237-305): files that appear later are processed; outputs that exist are
    skipped; MOTION_TRIM_WATCH_IDLE_EXIT_SEC ends the loop (the reference's never ends)."""
    names = ["kat_seg_S1", "batchclip_seed100", "kat_seg_S4"]
    p = cases()[names[0]].params
    with tempfile.TemporaryDirectory() as d:
        ind, outd = Path(d) / "in", Path(d) / "out"
        ind.mkdir()
        outd.mkdir()
        write_case(cases()[names[0]], ind / f"{names[0]}.mvs")
        write_case(cases()[names[2]], ind / f"{names[2]}.mvs")
        (outd / f"{names[2]}.mvs").write_text("already there")  # resume semantics: skipped
        env = dict(os.environ)
        import ref_runner

        env.update(ref_runner.env_for(p, 10.0, None))
        env.update({"WATCH_MODE": "1", "MOTION_TRIM_WATCH_IDLE_EXIT_SEC": "6", "PARALLEL_STREAMS": "1"})
        proc = subprocess.Popen([str(BIN), "--print-segments", str(ind), str(outd)], env=env, stdout=subprocess.PIPE,
                                stderr=subprocess.STDOUT, text=True)

        def late():
            time.sleep(2.5)
            tmp = ind / "late.partial"
            write_case(cases()[names[1]], tmp)
            tmp.rename(ind / f"{names[1]}.mvs")

        t = threading.Thread(target=late)
        t.start()
        out, _ = proc.communicate(timeout=120)
        t.join()
    assert proc.returncode == 0, out
    assert "Starting Watch Mode" in out and f"New file detected: {names[1]}.mvs" in out
    got = {}
    for line in out.splitlines():
        if line.startswith("RESULT "):
            parts = line.split()
            got[parts[1][:-4]] = dict(kv.split("=") for kv in parts[2:])
    assert sorted(got) == sorted(names[:2]), out  # the third had an output already
    for n in names[:2]:
        assert int(got[n]["decision"]) == expected(n)["decision"]
