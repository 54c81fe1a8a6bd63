"""Runs oracle/_ref/ref_scan — the reference's OWN motion_scanner.cpp / pipeline.cpp compiled against
the fake-libav shim — on an MVS1 file, with the knobs passed through the environment exactly as the
reference reads them (include/motion_trim/config.hpp). Test infrastructure only."""
from __future__ import annotations

import os
import struct
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF_BIN = ROOT / "oracle" / "_ref" / "ref_scan"
RES = struct.Struct("<iiiIIIddddqqqqqqqq")  # RefResult in oracle/ref_harness.cpp


def available() -> bool:
    return REF_BIN.exists() and os.access(REF_BIN, os.X_OK)


def env_for(params, chunk_sec=None, target_fps=None):
    """motionscan.Params → the reference's environment variables."""
    e = {
        "MV_THRESHOLD_SQ": repr(float(params.mv_threshold_sq)),
        "BLOCK_SIZE": str(int(params.block_size)),
        "BLOCK_SHIFT": str(int(params.block_shift)),
        "VECTORS_NEEDED": str(int(params.vectors_needed)),
        "CLUSTERS_NEEDED": str(int(params.clusters_needed)),
        "VERTICAL_MASK": repr(float(np.float32(params.vertical_mask))),
        "MAX_GAP_SEC": repr(float(params.max_gap_sec)),
        "PADDING_SEC": repr(float(params.padding_sec)),
        "MIN_SAVINGS_PCT": repr(float(params.min_savings_pct)),
    }
    if chunk_sec is not None:
        e["CHUNK_DURATION_SEC"] = repr(float(chunk_sec))
    if target_fps is not None:
        e["TARGET_FPS"] = repr(float(target_fps))
    return e


def run(mvs_path, params, threads=2, passes=1, warmup=0, chunk_sec=None, target_fps=None, out_path=None):
    out_path = out_path or (str(mvs_path) + ".ref.bin")
    env = {k: v for k, v in os.environ.items() if k not in env_for(params) and k not in ("CHUNK_DURATION_SEC", "TARGET_FPS")}
    env.update(env_for(params, chunk_sec, target_fps))
    subprocess.run([str(REF_BIN), str(mvs_path), out_path, str(threads), str(passes), str(warmup)], check=True, env=env,
                   stdout=subprocess.DEVNULL)
    raw = Path(out_path).read_bytes()
    (scan_ok, run_rc, decision, n_ts, n_segs, n_pass, duration, removed, pct, fps, analyze_us, decode_us, scan_wall_us,
     run_wall_us, par_sum_us, par_max_us, par_wall_us, par_found) = RES.unpack_from(raw, 0)
    ts = np.frombuffer(raw, dtype="<f8", count=n_ts, offset=RES.size).copy()
    segs = np.frombuffer(raw, dtype="<f8", count=2 * n_segs, offset=RES.size + 8 * n_ts).reshape(-1, 2).copy()
    os.unlink(out_path)
    return dict(scan_ok=scan_ok, run_rc=run_rc, decision=decision, ts=ts, segs=segs, duration=duration,
                time_removed=removed, saved_pct=pct, fps=fps, analyze_us=analyze_us, decode_us=decode_us,
                scan_wall_us=scan_wall_us, run_wall_us=run_wall_us, passes=n_pass, par_analyze_sum_us=par_sum_us,
                par_analyze_max_us=par_max_us, par_wall_us=par_wall_us, par_motion_frames=par_found)
