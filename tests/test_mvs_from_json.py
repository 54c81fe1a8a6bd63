"""CPU: tools/mvs_from_json.py turns the reference's tools/extract_mvs JSON dump format (extract_mvs.cpp:96-176)
into an MVS1 stream whose records reproduce the int16 fields check_frame reads."""
import importlib.util
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import kats
import motionscan as ms
from motionscan import mvs_io
import oracle_lib as orc

ROOT = Path(__file__).resolve().parent.parent


def dump_like_extract_mvs(frames, tb=(1, 15360)):
    """What extract_mvs would print for these frames (quarter-pel motion, source=-1)."""
    out = {"input": "synthetic.mp4", "time_base": f"{tb[0]}/{tb[1]}", "frames": []}
    for i, (pts_ticks, ftype, recs) in enumerate(frames):
        mvs = []
        for r in recs:
            mx, my = int(r["motion_x"]), int(r["motion_y"])
            mvs.append({"dst_x": int(r["dst_x"]), "dst_y": int(r["dst_y"]), "src_x": round(int(r["dst_x"]) + mx / 4, 3),
                        "src_y": round(int(r["dst_y"]) + my / 4, 3), "w": int(r["w"]), "h": int(r["h"]), "motion_x": mx,
                        "motion_y": my, "motion_scale": 4, "source": -1})
        out["frames"].append({"frame_index": i + 1, "pts_seconds": round(pts_ticks * tb[0] / tb[1], 6), "frame_type": ftype,
                              "num_mvs": len(mvs), "motion_vectors": mvs})
    return out


def test_json_round_trip(tmp_path):
    spec = ms.synth_preset(0, 9)
    spec.width, spec.height = 352, 288
    n = 40
    cnt, off, recs, _ = ms.synth_host(spec, 0, n)
    # give the records sub-pel motion whose C-truncated quotient is the generator's integer displacement
    recs = recs.copy()
    recs["motion_x"] = (recs["src_x"].astype(np.int32) - recs["dst_x"]) * 4 + np.sign(recs["src_x"].astype(np.int32) - recs["dst_x"]) * 3
    recs["motion_y"] = (recs["src_y"].astype(np.int32) - recs["dst_y"]) * 4
    frames = [(512 * i, "I" if cnt[i] == 0 else "P", recs[int(off[i]) : int(off[i + 1])]) for i in range(n)]
    jpath, mpath = tmp_path / "dump.json", tmp_path / "out.mvs"
    jpath.write_text(json.dumps(dump_like_extract_mvs(frames)))
    subprocess.run([sys.executable, str(ROOT / "tools" / "mvs_from_json.py"), str(jpath), str(mpath), "--width", "352", "--height",
                    "288", "--fps", "30"], check=True, capture_output=True)
    m = mvs_io.read_mvs(mpath, ms.MV_DTYPE)
    assert (m["width"], m["height"], m["tb"], m["fps"]) == (352, 288, (1, 15360), (30, 1))
    assert np.array_equal(m["frames"]["n_records"], cnt)
    assert np.array_equal(m["frames"]["pts"], 512 * np.arange(n))
    for f in ("src_x", "src_y", "dst_x", "dst_y", "w", "h", "motion_x", "motion_y"):
        assert np.array_equal(m["recs"][f], recs[f]), f
    assert bool(m["frames"]["flags"][0] & mvs_io.KEY) and bool(m["frames"]["flags"][30] & mvs_io.KEY)
    # and the scan of the converted stream equals the scan of the original records
    p = kats.env_params()
    cfg = orc.make_cfg(p, *orc.geometry(352, 288))
    assert np.array_equal(orc.scan_frames(cfg, np.ascontiguousarray(m["recs"]), off)[0], orc.scan_frames(cfg, recs, off)[0])


def test_c_division_truncates_toward_zero():
    spec = importlib.util.spec_from_file_location("mvs_from_json", ROOT / "tools" / "mvs_from_json.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert [mod.c_div(a, 4) for a in (-9, -8, -7, -1, 0, 1, 7, 8, 9)] == [-2, -2, -1, 0, 0, 0, 1, 2, 2]


@pytest.mark.gpu
def test_converted_dump_through_the_cuda_path(tmp_path):
    """SURVEY §8(f) N3 on the GPU: an extract_mvs-style JSON dump (reference tools/extract_mvs.cpp:146-165) →
    tools/mvs_from_json.py → `motion_trim_b200 --print-segments` gives the oracle's decision and segments, and the
    library's per-frame flags for the converted records equal the oracle's."""
    import os

    from test_gpu_parity import cfg_for, oracle_tail
    from test_host_cli import BIN, parse

    spec = ms.synth_preset(0, 21)
    spec.width, spec.height = 704, 576
    n = 180
    cnt, off, recs, _ = ms.synth_host(spec, 0, n)
    recs = recs.copy()
    recs["motion_x"] = (recs["src_x"].astype(np.int32) - recs["dst_x"]) * 4 + np.sign(recs["src_x"].astype(np.int32) - recs["dst_x"]) * 2
    recs["motion_y"] = (recs["src_y"].astype(np.int32) - recs["dst_y"]) * 4
    tb = (1, 15360)
    frames = [(512 * i, "I" if cnt[i] == 0 else "P", recs[int(off[i]) : int(off[i + 1])]) for i in range(n)]
    jpath, mpath = tmp_path / "dump.json", tmp_path / "clip.mvs"
    jpath.write_text(json.dumps(dump_like_extract_mvs(frames, tb)))
    subprocess.run([sys.executable, str(ROOT / "tools" / "mvs_from_json.py"), str(jpath), str(mpath), "--width", "704", "--height",
                    "576", "--fps", "30"], check=True, capture_output=True)
    p = kats.env_params()
    m = mvs_io.read_mvs(mpath, ms.MV_DTYPE)
    conv = np.ascontiguousarray(m["recs"])
    pts = m["frames"]["pts"].astype(np.float64) * (tb[0] / tb[1])
    cfg = cfg_for(p, 704, 576)
    of, oc = orc.scan_frames(cfg, conv, off)
    duration = m["duration_us"] / 1e6
    osegs, ores = oracle_tail(p, pts, of, duration)
    # 1. the library on the converted records
    with ms.Context(0, p, 1 << 12, 4 << 20) as ctx:
        ctx.video_open(1, 704, 576)
        ctx.submit(1, pts, cnt, conv)
        fl, cn = ctx.collect(1)
    assert np.array_equal(fl, of) and np.array_equal(cn, oc) and of.any()
    # 2. the product CLI on the converted file
    import ref_runner

    env = dict(os.environ)
    env.update(ref_runner.env_for(p, 2.0, None))
    r = subprocess.run([str(BIN), "--print-segments", str(mpath), str(tmp_path / "out.mp4")], env=env, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    res, segs = parse(r.stdout)
    assert int(res["decision"]) == ores.decision
    want = osegs if ores.decision == ms.CUT else (np.array([(0.0, duration)], dtype=osegs.dtype) if ores.decision == ms.FULL_COPY else osegs[:0])
    assert np.array(segs).reshape(-1, 2).tobytes() == np.ascontiguousarray(want).tobytes()
