"""CPU: the C-ABI library loads, exports every symbol include/motionscan.h declares, its pure-host
entry points work without a GPU, and the compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import motionscan as ms

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "motionscan.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mscan_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 35
    L = ms.lib()
    for n in names:
        assert hasattr(L, n), f"{n} declared in motionscan.h but not exported"
        assert n in ms.SYMBOLS, f"{n} has no ctypes prototype"
    assert L.mscan_abi_version() == 3


def test_struct_layouts():
    assert ms.MV_DTYPE.itemsize == 40
    assert C.sizeof(ms.Params) == 56
    assert C.sizeof(ms.Geometry) == 16
    assert C.sizeof(ms.VideoResult) == 40 and ms.RESULT_DTYPE.itemsize == 40
    assert C.sizeof(ms.Stats) == 112
    assert ms.MV8_DTYPE.itemsize == 8
    assert C.sizeof(ms.MvgenSpec) == 96  # + scatter, p_rec_move (SURVEY §8(d) config-5 shape)


def test_defaults_and_env(monkeypatch):
    p = ms.default_params()  # config.hpp:57-123
    assert (p.mv_threshold_sq, p.block_size, p.block_shift, p.vectors_needed, p.clusters_needed) == (16.0, 16, 4, 2, 2)
    assert (p.max_gap_sec, p.padding_sec, p.min_savings_pct) == (5.0, 0.5, 5.0)
    assert p.adjacency == 4  # the reference's 4-connectivity
    assert p.vertical_mask == np.float32(0.05)
    monkeypatch.setenv("MV_THRESHOLD_SQ", "4.0")
    monkeypatch.setenv("VECTORS_NEEDED", "4")
    monkeypatch.setenv("VERTICAL_MASK", "0.1")
    monkeypatch.setenv("MAX_GAP_SEC", "7.5")
    q = ms.Params()
    assert ms.lib().mscan_params_from_env(C.byref(q)) == ms.OK
    assert (q.mv_threshold_sq, q.vectors_needed, q.max_gap_sec) == (4.0, 4, 7.5)
    assert q.vertical_mask == np.float32(0.1)
    monkeypatch.setenv("CLUSTERS_NEEDED", "banana")  # reference: std::stoi throws → terminate
    assert ms.lib().mscan_params_from_env(C.byref(q)) == ms.ERR_INVALID


def test_no_gpu_means_error_not_fallback(have_gpu):
    if have_gpu:
        pytest.skip("GPU present; covered by the gpu tests")
    n = C.c_int(-1)
    assert ms.lib().mscan_device_count(C.byref(n)) == ms.ERR_CUDA
    with pytest.raises(ms.MscanError) as e:
        ms.Context(0, ms.default_params())
    assert e.value.code == ms.ERR_CUDA


def test_host_generator_deterministic_and_structured():
    spec = ms.synth_preset(0, 1)
    cnt, off, recs, pts = ms.synth_host(spec, 0, 95, n_threads=3)
    cnt2, off2, recs2, pts2 = ms.synth_host(spec, 0, 95, n_threads=1)
    assert np.array_equal(cnt, cnt2) and recs.tobytes() == recs2.tobytes() and np.array_equal(pts, pts2)
    # any sub-range regenerates identically (counter-based)
    c3, o3, r3, p3 = ms.synth_host(spec, 40, 20)
    assert r3.tobytes() == recs[int(off[40]) : int(off[60])].tobytes() and np.array_equal(p3, pts[40:60])
    assert cnt[0] == 0 and cnt[30] == 0 and cnt[60] == 0 and cnt[90] == 0  # I-frames carry no records
    assert cnt[1] >= 8160  # at least one record per macroblock on P-frames
    assert pts[31] == 31 / 30.0
    assert set(np.unique(recs["w"])) <= {8, 16} and (recs["motion_scale"] == 4).all()
    # padding bytes are zero so host and device generators can be compared byte-wise
    raw = recs.view(np.uint8).reshape(-1, 40)
    assert not raw[:, 14:16].any() and not raw[:, 34:40].any()


def test_pack_records_is_the_byte_range_6_to_14():
    """mscan_pack_records (no GPU needed): out[i] == bytes [6,14) of recs[i] == (src_x, src_y, dst_x, dst_y)."""
    rng = np.random.default_rng(7)
    for n in (0, 1, 7, 4096, 100_003):
        raw = rng.integers(0, 256, size=(n, 40), dtype=np.uint8)
        recs = raw.reshape(-1).view(ms.MV_DTYPE)
        out = ms.pack_records(recs)
        assert out.dtype == ms.MV8_DTYPE and len(out) == n
        assert out.view(np.uint8).reshape(n, 8).tobytes() == raw[:, 6:14].tobytes()
        for f in ("src_x", "src_y", "dst_x", "dst_y"):
            assert np.array_equal(out[f], recs[f])
    # misaligned output is refused, not silently mis-stored
    buf = np.zeros(8 * 4 + 4, np.uint8)
    assert ms.lib().mscan_pack_records(recs.ctypes.data, 4, buf.ctypes.data + 4) == ms.ERR_INVALID


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/motionscan.h compiles as C99 (-pedantic) and a C program links against libmotionscan.so: the ABI
    really is plain C (no C++ types, no torch types). The program only calls GPU-free entry points."""
    import subprocess

    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "motionscan.h"
int main(void) {
  mscan_params p;
  mscan_geometry g;
  mscan_mv recs[3];
  mscan_mv8 out[3];
  int i;
  if (mscan_abi_version() != MSCAN_ABI_VERSION) return 1;
  if (mscan_params_default(&p) != MSCAN_OK) return 2;
  if (mscan_geometry_from_dims(&p, 1920, 1080, &g) != MSCAN_OK || g.grid_w != 120 || g.grid_h != 68 || g.vertical_margin != 3) return 3;
  memset(recs, 0, sizeof recs);
  for (i = 0; i < 3; ++i) { recs[i].src_x = (int16_t)(10 + i); recs[i].src_y = -7; recs[i].dst_x = (int16_t)(100 * i); recs[i].dst_y = 32767; }
  if (mscan_pack_records(recs, 3, out) != MSCAN_OK) return 4;
  for (i = 0; i < 3; ++i) if (out[i].src_x != 10 + i || out[i].src_y != -7 || out[i].dst_x != 100 * i || out[i].dst_y != 32767) return 5;
  printf("%d %d %s\n", (int)sizeof(mscan_mv), (int)sizeof(mscan_mv8), mscan_status_string(MSCAN_ERR_CUDA));
  return 0;
}
''')
    exe = tmp_path / "abi"
    pkg = ROOT / "motion-estimated-video-trimmer_b200"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", str(ROOT / "include"), str(src), "-o", str(exe),
                    "-L", str(pkg), "-lmotionscan", f"-Wl,-rpath,{pkg}"], check=True, capture_output=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.split()[:2] == ["40", "8"]


def test_feed_harness_standin_writes_full_native_records():
    """The decode stand-in of the measurement harness (csrc/feed_harness.cpp) expands 8-byte coordinates into complete
    40-byte AVMotionVector records — byte-identical to the generator's own records — with and without AVX-512."""
    spec = ms.synth_preset(4, 5)
    cnt, off, recs, pts = ms.synth_host(spec, 1, 3)
    r8 = ms.pack_records(recs)
    ex = ms.feed_expand(r8)
    for f in ("source", "src_x", "src_y", "dst_x", "dst_y", "flags", "motion_x", "motion_y", "motion_scale"):
        assert np.array_equal(ex[f], recs[f]), f
    assert ms.pack_records(ex).tobytes() == r8.tobytes()
    assert hasattr(ms.feed_lib(), "mscan_feed_run")


def test_generator_library_alone_matches_and_spec_stream_shape():
    """libmvgen.so (what bench.py --impl reference loads instead of the product library) generates the same bytes;
    preset 5 is SURVEY §8(d) config 5 as specified: 16 320 records per 1080p frame, ~10 % moving, ~0.1 % out of frame."""
    from motionscan import mvgen

    for preset in (0, 4, 5):
        a = ms.synth_host(ms.synth_preset(preset, 9), 3, 6)
        b = mvgen.synth_host(mvgen.synth_preset(preset, 9), 3, 6)
        assert a[2].tobytes() == b[2].tobytes() and np.array_equal(a[0], b[0]) and np.array_equal(a[3], b[3])
    cnt, off, recs, pts = mvgen.synth_host(mvgen.synth_preset(5, 5), 0, 40)
    assert (cnt == 16320).all()
    moving = (recs["src_x"] != recs["dst_x"]) | (recs["src_y"] != recs["dst_y"])
    assert 0.09 < moving.mean() < 0.11
    d = np.stack([recs["dst_x"].astype(int) - recs["src_x"], recs["dst_y"].astype(int) - recs["src_y"]])
    assert d.min() == -8 and d.max() == 8
    oob = (recs["dst_x"] < 0) | (recs["dst_x"] >= 1920) | (recs["dst_y"] < 0) | (recs["dst_y"] >= 1088)
    assert 0.0003 < oob.mean() < 0.003
    # uniform dst: consecutive records are NOT raster-ordered
    assert (np.diff(recs["dst_y"][:1000].astype(int)) < 0).mean() > 0.3


def test_elided_form_round_trips_on_cpu():
    """mscan_elide_records (the wire form of MSCAN_STAGING_ELIDE) against the numpy reference decoder, at tile and
    block boundaries, for static, mixed and all-moving frames; it never takes more than 8 B/record + per-tile overhead."""
    from test_oracle_kats import random_frame

    rng = np.random.default_rng(8)
    cnt, off, recs, pts = ms.synth_host(ms.synth_preset(4, 5), 1, 2)
    cctv = np.ascontiguousarray(recs[: int(off[1])])
    moving = random_frame(rng, 3000, 1920, 1080, 5)
    static = cctv.copy()
    static["src_x"], static["src_y"] = static["dst_x"], static["dst_y"]
    for frame in (cctv, moving, static):
        for n in (0, 1, 31, 32, 33, 1023, 1024, 1025, 2049, len(frame)):
            if n > len(frame):
                continue
            sub = np.ascontiguousarray(frame[:n])
            enc, te = ms.elide_records(sub)
            assert len(te) == (n + 1023) // 1024 and len(enc) % 16 == 0
            assert ms.unelide_records(enc, te, n).tobytes() == ms.pack_records(sub).tobytes()
            assert len(enc) <= 8 * n + 352 * max(len(te), 1)
    enc, _ = ms.elide_records(static)
    assert len(enc) / len(static) < 4.4  # 4 B dst + 8 B header per 32 records


def _moving_mv8(recs):
    r8 = ms.pack_records(recs) if len(recs) else np.zeros(0, ms.MV8_DTYPE)
    keep = (r8["src_x"] != r8["dst_x"]) | (r8["src_y"] != r8["dst_y"])
    return r8[keep]


def test_compaction_keeps_exactly_the_moving_records_in_order():
    """mscan_compact_records (the wire form of MSCAN_STAGING_COMPACT): the projections of the records with src != dst,
    in record order — at every length around the 8-record vector step, for static, mixed and all-moving input."""
    from test_oracle_kats import random_frame

    rng = np.random.default_rng(11)
    cnt, off, recs, pts = ms.synth_host(ms.synth_preset(4, 5), 1, 2)
    cctv = np.ascontiguousarray(recs[: int(off[1])])
    moving = random_frame(rng, 3000, 1920, 1080, 5)
    static = cctv.copy()
    static["src_x"], static["src_y"] = static["dst_x"], static["dst_y"]
    half = cctv.copy()  # records that differ in ONE coordinate only, either one
    half["src_x"] = half["dst_x"]
    half["src_y"] = half["dst_y"] + (np.arange(len(half)) % 3 == 0)
    assert 0 < len(_moving_mv8(cctv)) < len(cctv) and len(_moving_mv8(static)) == 0
    for frame in (cctv, moving, static, half):
        for n in (0, 1, 7, 8, 9, 15, 16, 17, 63, 64, 65, 1000, len(frame)):
            if n > len(frame):
                continue
            sub = np.ascontiguousarray(frame[:n])
            got = ms.compact_records(sub)
            assert got.tobytes() == _moving_mv8(sub).tobytes(), n
    # every alignment of the first record against a 64-byte line (the vector loop starts at the first record on a line),
    # and a source that is not even 8-byte aligned
    for shift in range(9):
        sub = cctv[shift : shift + 700]
        assert sub.flags["C_CONTIGUOUS"] and sub.ctypes.data == cctv.ctypes.data + 40 * shift
        assert ms.compact_records(sub).tobytes() == _moving_mv8(np.ascontiguousarray(sub)).tobytes(), shift
    raw = np.zeros(40 * 300 + 64, np.uint8)
    start = (-raw.ctypes.data) % 8 + 4
    raw[start : start + 40 * 300] = cctv[:300].view(np.uint8)
    odd = np.empty(300, ms.MV8_DTYPE)
    n_odd = C.c_uint64(0)
    assert ms.lib().mscan_compact_records(raw.ctypes.data + start, 300, odd.ctypes.data, C.byref(n_odd)) == ms.OK
    assert odd[: n_odd.value].tobytes() == _moving_mv8(np.ascontiguousarray(cctv[:300])).tobytes()
    # scattered moving records (the SURVEY §8(d) stream: 10 % at random) and frames that alternate between quiet and busy
    # stretches: the vector loop switches between its two step forms every 64 steps
    _, soff, srecs, _ = ms.synth_host(ms.synth_preset(5, 5), 0, 3)
    assert ms.compact_records(srecs).tobytes() == _moving_mv8(srecs).tobytes()
    mix = np.ascontiguousarray(srecs[:16000]).copy()
    for a in range(0, 16000, 3000):  # quiet stretches of 1500 records
        mix["src_x"][a : a + 1500], mix["src_y"][a : a + 1500] = mix["dst_x"][a : a + 1500], mix["dst_y"][a : a + 1500]
    busy = np.arange(16000) % 3000 >= 2200  # and stretches where everything moves
    mix["src_x"][busy] = mix["dst_x"][busy] + 1
    for shift in (0, 3):
        assert ms.compact_records(mix[shift:]).tobytes() == _moving_mv8(np.ascontiguousarray(mix[shift:])).tobytes()
    # per-frame form used by callers that compact into their own (pinned) buffer
    cnt, off, recs, _ = ms.synth_host(ms.synth_preset(0, 3), 0, 12)
    out, mcnt = ms.compact_frames(recs, off)
    assert out.tobytes() == _moving_mv8(recs).tobytes()
    assert [int(x) for x in mcnt] == [len(_moving_mv8(np.ascontiguousarray(recs[int(off[f]) : int(off[f + 1])]))) for f in range(12)]
    # misaligned output is refused
    n_out = C.c_uint64(0)
    buf = np.zeros(8 * 4 + 4, np.uint8)
    assert ms.lib().mscan_compact_records(recs.ctypes.data, 4, buf.ctypes.data + 4, C.byref(n_out)) == ms.ERR_INVALID


def test_host_passes_agree_without_avx512():
    """The portable versions of the three host passes (projection, static-elided form, compaction) write the same bytes
    as the AVX-512 ones: a child process with MSCAN_NO_AVX512=1 prints digests that must equal this process's."""
    import hashlib
    import os
    import subprocess
    import sys

    code = r'''
import hashlib, numpy as np, motionscan as ms
cnt, off, recs, pts = ms.synth_host(ms.synth_preset(4, 5), 0, 6)
h = hashlib.sha256()
h.update(ms.pack_records(recs).tobytes())
enc, eoff, te = ms.elide_frames(recs, off)
h.update(np.asarray(enc).tobytes()); h.update(np.asarray(te).tobytes())
out, mc = ms.compact_frames(recs, off)
h.update(out.tobytes()); h.update(mc.tobytes())
print(h.hexdigest())
'''
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(sys.path))
    a = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.strip()
    b = subprocess.run([sys.executable, "-c", code], env=dict(env, MSCAN_NO_AVX512="1"), capture_output=True, text=True, check=True).stdout.strip()
    assert len(a) == 64 and a == b
