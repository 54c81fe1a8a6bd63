"""The FFmpeg decode front-end of the host mirror (host/src/ffmpeg_frontend.cpp).

The image has no FFmpeg, so the front-end — written against the real libavformat/libavcodec API — is
compiled here against the fake libav of oracle/ffshim (test infrastructure: it demuxes MVS1 files and
"decodes" by attaching the stored AVMotionVector records as export_mvs side data). That is the same shim
the reference's own sources were compiled against to produce the golden fixtures, so on one input file
   reference:  FFmpeg API → MotionScanner::scan_range → check_frame → pipeline.cpp segment builder   (golden)
   here:       FFmpeg API → FFmpegFrontEnd::scan → mscan_pack_records → mscan_submit_packed → K-A → K-C
must agree bit for bit: frame selection (seek, TARGET_FPS skip, chunk edges), decisions, segments.
CPU: the build and the no-GPU refusal. GPU: parity."""
import subprocess
import tempfile
from pathlib import Path

import numpy as np
import pytest

from test_host_cli import HOST, parse, run_cli, write_case
from test_ref_golden import cases, expected

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "tests" / "_build" / "motion_trim_b200_ffshim"
FF_ENV = {"MOTION_TRIM_FRONTEND": "ffmpeg"}


@pytest.fixture(scope="module", autouse=True)
def _built():
    OUT.parent.mkdir(exist_ok=True)
    srcs = list((HOST / "src").glob("*.cpp")) + list((HOST / "include" / "motion_trim").glob("*.hpp")) + [
        ROOT / "oracle" / "ffshim" / "fake_libav.cpp", HOST.parent / "libmotionscan.so"]
    if OUT.exists() and all(s.stat().st_mtime <= OUT.stat().st_mtime for s in srcs):
        return
    subprocess.run(
        ["make", "-C", str(HOST), "FFMPEG=1", f"FF_CFLAGS=-I {ROOT / 'oracle' / 'ffshim'}",
         f"FF_LIBS={ROOT / 'oracle' / 'ffshim' / 'fake_libav.cpp'}", f"OUT={OUT}", f"RPATH={HOST.parent}"],
        check=True, capture_output=True)


def test_frontend_builds_and_refuses_without_gpu(have_gpu):
    assert OUT.exists()
    if have_gpu:
        pytest.skip("GPU present")
    c = cases()["kat_seg_S1"]
    with tempfile.TemporaryDirectory() as d:
        path = Path(d) / "x.mp4"  # the front-end probes content, not the name
        write_case(c, path)
        r = run_cli([str(path), str(Path(d) / "out.mp4")], c.params, binary=OUT, extra_env=FF_ENV)
    assert r.returncode == 1 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["kat_seg_S1", "kat_seg_S4", "kat_seg_S7", "clip60s_1080p_config0", "batchclip_seed100",
                                  "dense_4k_24f", "rand_params_0", "rand_params_3", "rand_params_5", "rand_params_7",
                                  "skip_tfps10_chunk10", "skip_tfps7_chunk2p5", "skip_tfps4_chunk7", "skip_tfps12p5_720p"])
def test_decode_fed_pipeline_matches_reference(name):
    c, e = cases()[name], expected(name)
    with tempfile.TemporaryDirectory() as d:
        path, out = Path(d) / "in.mp4", Path(d) / "out.mp4"
        write_case(c, path)
        r = run_cli(["--print-segments", str(path), str(out)], c.params, chunk_sec=c.chunk_sec, threads=c.threads,
                    target_fps=c.target_fps, binary=OUT, extra_env=FF_ENV)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "FFmpeg export_mvs front-end" in r.stdout
        res, segs = parse(r.stdout)
        assert int(res["decision"]) == e["decision"]
        assert np.array(segs).reshape(-1, 2).tobytes() == e["segs"].tobytes()
        if e["decision"]:
            assert float.fromhex(res["saved_pct"]) == e["saved_pct"]
            assert float.fromhex(res["time_removed"]) == e["time_removed"]
        else:
            assert "No motion found." in r.stdout and not Path(str(out) + ".concat.txt").exists()


@pytest.mark.gpu
def test_decode_fed_batch_directory():
    """Directory mode picks up the reference's media extensions (src/main.cpp:68-69) in FFmpeg builds."""
    names = ["batchclip_seed100", "batchclip_seed101", "kat_seg_S7", "kat_seg_S5"]
    p = cases()[names[0]].params
    with tempfile.TemporaryDirectory() as d:
        ind, outd = Path(d) / "in", Path(d) / "out"
        ind.mkdir()
        for n, ext in zip(names, (".mp4", ".mkv", ".ts", ".mov")):
            write_case(cases()[n], ind / f"{n}{ext}")
        (ind / "notes.txt").write_text("not media")
        r = run_cli(["--print-segments", str(ind), str(outd)], p, chunk_sec=10.0, binary=OUT,
                    extra_env={**FF_ENV, "PARALLEL_STREAMS": "2"})
        assert r.returncode == 0, r.stdout + r.stderr
        got = {}
        for line in r.stdout.splitlines():
            if line.startswith("RESULT "):
                parts = line.split()
                got[Path(parts[1]).stem] = dict(kv.split("=") for kv in parts[2:])
        assert sorted(got) == sorted(names)
        for n in names:
            e = expected(n)
            assert int(got[n]["decision"]) == e["decision"], n
            if e["decision"]:
                assert float.fromhex(got[n]["saved_pct"]) == e["saved_pct"], n
