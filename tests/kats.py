"""Known-answer vectors K1-K20 / S1-S8 of SURVEY.md §4 (derived from the reference's code; the
reference itself ships no tests). Shared by the oracle tests (CPU) and the GPU parity tests."""
from __future__ import annotations

import numpy as np

import motionscan as ms

W, H = 1920, 1080  # → gw=120, gh=68, margin 3, live rows [3,65)


def env_params(**kw):
    """Shipped config/motion_trim.env values: T²=4, VECTORS_NEEDED=4, CLUSTERS_NEEDED=2."""
    p = ms.shipped_env_params()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def code_defaults(**kw):
    p = ms.default_params()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def cell(gx, gy, n=4, dx=2, dy=0):
    """n records whose dst lies in cell (gx,gy) at the 8x8 sub-block centres, src = dst - (dx,dy)."""
    r = np.zeros(n, dtype=ms.MV_DTYPE)
    for k in range(n):
        r["dst_x"][k] = 16 * gx + (12 if (k & 1) else 4)
        r["dst_y"][k] = 16 * gy + (12 if (k & 2) else 4)
    r["src_x"] = r["dst_x"] - dx
    r["src_y"] = r["dst_y"] - dy
    r["source"] = -1
    r["w"] = 8
    r["h"] = 8
    r["motion_scale"] = 4
    return r


def raw(points, dx=2, dy=0):
    r = np.zeros(len(points), dtype=ms.MV_DTYPE)
    for k, (x, y) in enumerate(points):
        r["dst_x"][k], r["dst_y"][k] = x, y
    r["src_x"] = r["dst_x"] - dx
    r["src_y"] = r["dst_y"] - dy
    return r


def cat(*parts):
    # np.concatenate would repack the 40-byte record dtype to 32 bytes; copy into a fresh array instead
    out = np.zeros(sum(len(p) for p in parts), dtype=ms.MV_DTYPE)
    at = 0
    for p in parts:
        out[at : at + len(p)] = p
        at += len(p)
    return out


# name → (params, records or None, expected flag, expected full_count)
def frame_kats():
    E = env_params
    k = {}
    k["K1_empty"] = (E(), None, 0, 0)
    k["K2_isolated"] = (E(), cell(10, 10), 0, 0)
    k["K3_horizontal"] = (E(), cat(cell(10, 10), cell(11, 10)), 1, 2)
    k["K4_vertical"] = (E(), cat(cell(10, 10), cell(10, 11)), 1, 2)
    k["K5_diagonal"] = (E(), cat(cell(10, 10), cell(11, 11)), 0, 0)
    k["K6_three_votes"] = (E(), cat(cell(10, 10, 3), cell(11, 10, 4)), 0, 0)
    k["K7_below_threshold"] = (E(), cat(cell(10, 10, 4, 1, 1), cell(11, 10, 4, 1, 1)), 0, 0)
    k["K8_threshold_equal"] = (E(), cat(cell(10, 10, 4, 2, 0), cell(11, 10, 4, 0, 2)), 1, 2)
    k["K9_left_edge"] = (E(), cat(cell(0, 10), cell(1, 10)), 0, 1)
    k["K10_right_edge"] = (E(), cat(cell(118, 10), cell(119, 10)), 0, 1)
    k["K11_masked_row"] = (E(), cat(cell(10, 2), cell(11, 2)), 0, 0)
    k["K12_first_live_row"] = (E(), cat(cell(10, 3), cell(11, 3)), 1, 2)
    k["K13_row65_masked"] = (E(), cat(cell(10, 64), cell(10, 65)), 0, 0)
    k["K14_three_in_row"] = (E(), cat(cell(10, 10), cell(11, 10), cell(12, 10)), 1, 3)
    k["K15_saturation"] = (E(), cat(cell(10, 10, 300), cell(11, 10, 300)), 1, 2)
    k["K16_out_of_frame"] = (E(), raw([(-5, 100)] * 4 + [(1925, 100)] * 4 + [(100, -3)] * 4 + [(100, 1090)] * 4), 0, 0)
    k["K17_code_defaults"] = (code_defaults(), cat(cell(10, 10, 2, 4, 0), cell(11, 10, 2, 0, 4)), 1, 2)
    k["K18_defaults_below"] = (code_defaults(), cat(cell(10, 10, 2, 3, 2), cell(11, 10, 2, 2, 3)), 0, 0)
    k["K19_edge_need1"] = (E(clusters_needed=1), cat(cell(0, 10), cell(1, 10)), 1, 1)
    k["K20_empty_need0"] = (E(clusters_needed=0), None, 0, 0)
    return k


# name → (timestamps, duration, expected segments, out_dur, time_removed, decision)
def segment_kats():
    s = {}
    s["S1"] = ([1.0, 2.0, 10.0], 60.0, [(0.5, 2.5), (9.5, 10.5)], 3.0, 57.0, ms.CUT)
    s["S2"] = ([0.2, 0.3], 60.0, [(0.0, 0.8)], 0.8, 59.2, ms.CUT)
    s["S3"] = ([59.8], 60.0, [(59.3, 60.0)], 0.70000000000000284, 59.3, ms.CUT)
    s["S4"] = ([1.0, 6.0, 11.0, 16.000000001], 60.0, [(0.5, 11.5), (15.500000001, 16.500000001)], 12.0, 48.0, ms.CUT)
    s["S5"] = ([i / 30 for i in range(1800)], 60.0, [(0.0, 60.0)], 60.0, 0.0, ms.FULL_COPY)
    s["S6"] = ([i / 30 for i in range(1710)], 60.0, [(0.0, 1709 / 30 + 0.5)], None, None, ms.FULL_COPY)
    s["S7"] = ([], 60.0, [], 0.0, 0.0, ms.NO_MOTION)
    s["S8"] = ([10.0, 10.0, 30.0], 60.0, [(9.5, 10.5), (29.5, 30.5)], 2.0, 58.0, ms.CUT)
    return s
