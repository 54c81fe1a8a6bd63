"""The C++ host mirror (motion-estimated-video-trimmer_b200/host → motion_trim_b200): the reference's
CLI contract for this path on top of the C ABI. CPU: it builds and refuses to run without a GPU
(no fallback). GPU: its decisions/segments equal the reference's own pipeline outputs (golden fixture),
single-file and batch mode."""
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np
import pytest

from motionscan import mvs_io
import ref_runner
from test_ref_golden import GOLDEN, cases, expected

ROOT = Path(__file__).resolve().parent.parent
HOST = ROOT / "motion-estimated-video-trimmer_b200" / "host"
BIN = HOST / "motion_trim_b200"


@pytest.fixture(scope="module", autouse=True)
def _host_built():
    if not BIN.exists():
        subprocess.run(["make", "-C", str(HOST)], check=True)


def write_case(c, path):
    mvs_io.write_mvs(path, c.width, c.height, c.fps[0], c.fps[1], c.ticks, c.cnt, c.recs, has_mvs=c.has_mvs,
                     tb_num=c.tb[0], tb_den=c.tb[1], duration_us=c.duration_us)


def run_cli(args, params, chunk_sec=None, threads=None, extra_env=None, target_fps=None, binary=None):
    env = dict(os.environ)
    env.update(ref_runner.env_for(params, chunk_sec, target_fps))
    if threads:
        env["THREADS_PER_STREAM"] = str(threads)
    env.update(extra_env or {})
    return subprocess.run([str(binary or BIN), *args], env=env, capture_output=True, text=True, timeout=300)


def parse(stdout):
    res, segs = None, []
    for line in stdout.splitlines():
        if line.startswith("RESULT "):
            res = dict(kv.split("=") for kv in line.split()[1:] if "=" in kv)
        elif line.startswith("SEGMENT "):
            _, a, b = line.split()
            segs.append((float.fromhex(a), float.fromhex(b)))
    return res, segs


def test_cli_refuses_without_gpu(have_gpu):
    if have_gpu:
        pytest.skip("GPU present")
    c = cases()["kat_seg_S1"]
    with tempfile.TemporaryDirectory() as d:
        path = Path(d) / "x.mvs"
        write_case(c, path)
        r = run_cli([str(path), str(Path(d) / "out.mp4")], c.params)
    assert r.returncode == 1
    assert "no CUDA device" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["kat_seg_S1", "kat_seg_S4", "kat_seg_S5", "kat_seg_S7", "clip60s_1080p_config0",
                                  "batchclip_seed100", "batchclip_seed101", "dense_4k_24f", "rand_params_0", "rand_params_5",
                                  "rand_params_7", "skip_tfps10_chunk10", "skip_tfps7_chunk2p5", "skip_tfps4_chunk7",
                                  "skip_tfps12p5_720p"])
def test_cli_single_file_matches_reference_pipeline(name):
    c, e = cases()[name], expected(name)
    with tempfile.TemporaryDirectory() as d:
        path, out = Path(d) / "in.mvs", Path(d) / "out.mp4"
        write_case(c, path)
        r = run_cli(["--print-segments", str(path), str(out)], c.params, chunk_sec=c.chunk_sec, threads=c.threads,
                    target_fps=c.target_fps)
        assert r.returncode == 0, r.stdout + r.stderr
        res, segs = parse(r.stdout)
        concat = Path(str(out) + ".concat.txt")
        assert int(res["decision"]) == e["decision"]
        assert np.array(segs).reshape(-1, 2).tobytes() == e["segs"].tobytes()
        if e["decision"] == 0:
            assert not concat.exists() and "No motion found." in r.stdout  # pipeline.cpp:308-319: no job, no file
        else:
            assert float.fromhex(res["saved_pct"]) == e["saved_pct"]
            assert float.fromhex(res["time_removed"]) == e["time_removed"]
            # concat list in the reference's format (ffmpeg_executor.cpp:44-50)
            want = "".join(f"file '{path.resolve()}'\ninpoint {a:.2f}\noutpoint {b:.2f}\n" for a, b in e["segs"] if b > a)
            assert concat.read_text() == want


@pytest.mark.gpu
def test_cli_batch_directory_matches_reference():
    names = ["batchclip_seed100", "batchclip_seed101", "cctv_1080p_600f", "stream_cfg4_720f", "kat_seg_S7", "kat_seg_S5"]
    p = cases()[names[0]].params  # all of these use the shipped-env parameter set
    with tempfile.TemporaryDirectory() as d:
        ind, outd = Path(d) / "in", Path(d) / "out"
        ind.mkdir()
        for n in names:
            write_case(cases()[n], ind / f"{n}.mvs")
        r = run_cli(["--print-segments", str(ind), str(outd)], p, chunk_sec=10.0, extra_env={"PARALLEL_STREAMS": "2"})
        assert r.returncode == 0, r.stdout + r.stderr
        got = {}
        for line in r.stdout.splitlines():
            if line.startswith("RESULT "):
                parts = line.split()
                got[parts[1][:-4]] = dict(kv.split("=") for kv in parts[2:])
        assert sorted(got) == sorted(names)
        for n in names:
            e = expected(n)
            assert int(got[n]["decision"]) == e["decision"], n
            assert int(got[n]["rc"]) == 0
            if e["decision"]:
                assert float.fromhex(got[n]["saved_pct"]) == e["saved_pct"], n
                assert (outd / f"{n}.mvs.concat.txt").exists()
            else:
                assert not (outd / f"{n}.mvs.concat.txt").exists()
        assert "BATCH SUMMARY" in r.stdout


def test_unparsable_knob_is_fatal_before_anything_runs(tmp_path):
    """One parser for the scan knobs (the library's): garbage in the environment ends the program with a message, like
    the reference's std::stod/stoi (include/motion_trim/config.hpp:28-53) — it does not fall back to a default."""
    c = cases()["kat_seg_S1"]
    path = tmp_path / "x.mvs"
    write_case(c, path)
    for knob in ("MV_THRESHOLD_SQ", "VERTICAL_MASK", "CHUNK_DURATION_SEC", "PARALLEL_STREAMS"):
        r = run_cli([str(path), str(tmp_path / "out.mp4")], c.params, extra_env={knob: "abc"})
        assert r.returncode == 2, (knob, r.stdout, r.stderr)
        assert "[ERROR]" in r.stderr


def test_concat_list_escapes_single_quotes(tmp_path):
    """File names are data: a single quote in the input path is written in the concat demuxer's own quoting."""
    src = tmp_path / "t.cpp"
    src.write_text(r"""
#include <cstdio>
#include "motion_trim/ffmpeg_queue.hpp"
int main() {
  std::vector<motion_trim::TimeSegment> s{{1.0, 2.5}, {3.0, 3.0}};
  std::fputs(motion_trim::build_concat_list("/data/it's here/a.mp4", s).c_str(), stdout);
}
""")
    exe = tmp_path / "t"
    subprocess.run(["g++", "-std=c++17", "-I", str(HOST / "include"), "-I", str(ROOT / "include"), "-o", str(exe), str(src),
                    str(HOST / "src" / "ffmpeg_queue.cpp"), "-pthread"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    assert out == "file '/data/it'\\''s here/a.mp4'\ninpoint 1.00\noutpoint 2.50\n"


@pytest.mark.gpu
def test_crafted_stream_headers_are_rejected(tmp_path):
    """An MVS1 header whose counts would wrap the bounds arithmetic must be refused (watch mode opens whatever appears
    in the directory), not read or DMA'd out of bounds."""
    import struct

    c = cases()["kat_seg_S1"]
    good = tmp_path / "good.mvs"
    write_case(c, good)
    raw = bytearray(good.read_bytes())
    hdr = list(mvs_io.HDR.unpack_from(raw, 0))
    names = ["magic", "width", "height", "tb_num", "tb_den", "fps_num", "fps_den", "duration_us", "n_frames", "flags", "records_offset", "n_records"]
    idx = {n: i for i, n in enumerate(names)}
    attacks = {
        "n_records_wraps": {"n_records": (1 << 64) // 40 + 3},
        "records_offset_huge": {"records_offset": (1 << 64) - 40},
        "n_frames_huge": {"n_frames": 0xFFFFFFFF},
    }
    for name, patch in attacks.items():
        h = list(hdr)
        for k, v in patch.items():
            h[idx[k]] = v
        bad = tmp_path / f"{name}.mvs"
        bad.write_bytes(mvs_io.HDR.pack(*h) + bytes(raw[mvs_io.HDR.size:]))
        r = run_cli([str(bad), str(tmp_path / "o.mp4")], c.params)
        assert r.returncode == 1 and "Failed to initialize probe" in r.stdout, (name, r.stdout, r.stderr)
    # a frame entry pointing past the records
    fr = np.frombuffer(bytes(raw[mvs_io.HDR.size : mvs_io.HDR.size + 24 * hdr[idx["n_frames"]]]), dtype=mvs_io.FRAME_DTYPE).copy()
    fr["first_record"][-1] = (1 << 64) - 2
    bad = tmp_path / "frame_wraps.mvs"
    bad.write_bytes(bytes(raw[: mvs_io.HDR.size]) + fr.tobytes() + bytes(raw[mvs_io.HDR.size + fr.nbytes :]))
    r = run_cli([str(bad), str(tmp_path / "o.mp4")], c.params)
    assert r.returncode == 1 and "Failed to initialize probe" in r.stdout
