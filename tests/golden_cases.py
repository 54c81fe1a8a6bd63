"""Golden cases: inputs are regenerated deterministically (generator presets, KAT builders, seeded
numpy); the EXPECTED outputs in tests/golden/ref_golden.json were produced by the reference's own
sources (oracle/_ref/ref_scan, see tests/golden/make_golden.py). An input digest in the fixture
detects any drift of the regenerated inputs."""
from __future__ import annotations

import hashlib

import numpy as np

import kats
import motionscan as ms

PARAM_FIELDS = ["mv_threshold_sq", "block_size", "block_shift", "vectors_needed", "clusters_needed", "vertical_mask",
                "max_gap_sec", "padding_sec", "min_savings_pct"]


def params_dict(p):
    return {k: (float(getattr(p, k)) if k in ("mv_threshold_sq", "vertical_mask", "max_gap_sec", "padding_sec", "min_savings_pct")
                else int(getattr(p, k))) for k in PARAM_FIELDS}


def params_from(d):
    p = ms.default_params()
    for k, v in d.items():
        setattr(p, k, v)
    return p


def digest(cnt, recs, ticks):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(cnt, dtype=np.uint32).tobytes())
    h.update(np.ascontiguousarray(recs).tobytes())
    h.update(np.ascontiguousarray(ticks, dtype=np.int64).tobytes())
    return h.hexdigest()


class Case:
    """One video: frames (rec_count, recs), pts ticks + time base, fps, params, reference run options."""

    def __init__(self, name, width, height, fps, tb, ticks, cnt, recs, params, has_mvs=None, chunk_sec=None, threads=2,
                 duration_us=None, target_fps=None):
        self.name, self.width, self.height, self.fps, self.tb = name, width, height, fps, tb
        self.ticks = np.asarray(ticks, dtype=np.int64)
        self.cnt = np.asarray(cnt, dtype=np.uint32)
        self.recs = recs
        self.params = params
        self.has_mvs = np.asarray(has_mvs, dtype=bool) if has_mvs is not None else (self.cnt > 0)
        # same default as mvs_io.write_mvs: frame 0 and every frame without vectors (I-frames) are seek points
        self.key = ~self.has_mvs
        self.key[0:1] = True
        self.chunk_sec, self.threads = chunk_sec, threads
        self.duration_us = duration_us
        self.target_fps = target_fps

    # ---- what the reference's host code does before check_frame (motion_scanner.cpp:303-371) ----
    def selected(self, chunked=True):
        """Frame indices handed to check_frame: by the chunked pipeline (default) or by one
        scan_range(0, duration) call."""
        import oracle_lib as orc

        fps = self.fps[0] / float(self.fps[1])
        chunk = (self.chunk_sec or 30.0) if chunked else self.duration + 1.0
        return orc.select_pipeline(self.ticks, self.key, self.tb[0] / float(self.tb[1]), fps, self.target_fps or 0.0,
                                   self.duration, chunk)

    def subset(self, idx):
        """(rec_count, rec_off, recs, pts) restricted to frames idx, in that order."""
        off = self.off
        cnt = np.where(self.has_mvs[idx], self.cnt[idx], 0).astype(np.uint32)
        parts = [self.recs[int(off[i]) : int(off[i + 1])] for i, h in zip(idx, self.has_mvs[idx]) if h]
        recs = kats.cat(*parts)
        o = np.zeros(len(idx) + 1, dtype=np.uint64)
        np.cumsum(cnt, out=o[1:])
        return cnt, o, recs, self.pts[idx]

    @property
    def off(self):
        o = np.zeros(len(self.cnt) + 1, dtype=np.uint64)
        np.cumsum(self.cnt, out=o[1:])
        return o

    @property
    def pts(self):
        # motion_scanner.cpp:304-305,361: pts = frame->pts * av_q2d(time_base)
        return self.ticks.astype(np.float64) * (self.tb[0] / float(self.tb[1]))

    @property
    def duration(self):
        if self.duration_us is not None:
            return self.duration_us / 1_000_000.0
        n = len(self.cnt)
        return (-(-n * 1_000_000 * self.fps[1] // self.fps[0])) / 1_000_000.0


def synth_case(name, config, seed, n_frames, params, width=None, height=None, **kw):
    spec = ms.synth_preset(config, seed)
    if width:
        spec.width, spec.height = width, height
    cnt, off, recs, _ = ms.synth_host(spec, 0, n_frames)
    fps = (int(spec.fps), 1)
    return Case(name, spec.width, spec.height, fps, (1, int(spec.fps)), np.arange(n_frames), cnt, recs, params, **kw)


def kat_frame_cases():
    """K1-K20 grouped by parameter set: one frame per KAT, 30 fps."""
    groups = {}
    for name, (p, recs, flag, count) in sorted(kats.frame_kats().items()):
        key = (p.mv_threshold_sq, p.vectors_needed, p.clusters_needed)
        groups.setdefault(key, (p, []))[1].append((name, recs))
    out = []
    for gi, (p, items) in enumerate(groups.values()):
        frames = [r for _, r in items]
        cnt = [0 if f is None else len(f) for f in frames]
        recs = kats.cat(*[f for f in frames if f is not None])
        has = np.array([f is not None for f in frames])
        c = Case(f"kat_frames_{gi}", kats.W, kats.H, (30, 1), (1, 30), np.arange(len(frames)), cnt, recs, p, has_mvs=has)
        c.kat_names = [n for n, _ in items]
        out.append(c)
    return out


def kat_segment_cases():
    """S1-S8: frames at the KAT timestamps (time base 1e-9 s), each carrying one horizontal cluster pair."""
    out = []
    active = kats.cat(kats.cell(10, 10), kats.cell(11, 10))
    for name, (ts, duration, *_rest) in sorted(kats.segment_kats().items()):
        p = kats.env_params()
        uniq = sorted(set(ts))  # one decoded frame per pts; S8's duplicate cannot occur in a real decode
        if uniq:
            ticks = [int(round(t * 1e9)) for t in uniq]
            cnt = [len(active)] * len(uniq)
            recs = kats.cat(*([active] * len(uniq)))
            has = np.ones(len(uniq), bool)
        else:  # S7: frames exist but none moves
            ticks, cnt, recs, has = [0, 10 ** 9], [0, 0], kats.cat(), np.zeros(2, bool)
        out.append(Case(f"kat_seg_{name}", kats.W, kats.H, (30, 1), (1, 10 ** 9), ticks, cnt, recs, p, has_mvs=has,
                        duration_us=int(duration * 1e6)))
    return out


def random_param_cases():
    out = []
    rng = np.random.default_rng(20260118)
    shapes = [(1920, 1080), (1280, 720), (352, 288), (640, 360)]
    for i in range(8):
        p = kats.env_params() if i % 2 else kats.code_defaults()
        p.vectors_needed = int(rng.integers(1, 6))
        p.clusters_needed = int(rng.integers(1, 5))
        p.mv_threshold_sq = float(rng.choice([1.0, 4.0, 4.5, 16.0, 25.0]))
        p.vertical_mask = float(rng.choice([0.05, 0.1, 0.2]))  # margin >= 1 on all shapes (margin 0 is UB in the reference)
        p.max_gap_sec = float(rng.choice([5.0, 1.0, 2.5]))
        p.padding_sec = float(rng.choice([0.5, 2.0, 0.0]))
        p.min_savings_pct = float(rng.choice([5.0, 50.0]))
        w, h = shapes[i % 4]
        out.append(synth_case(f"rand_params_{i}", 3 if i % 3 else 0, 40 + i, 600, p, w, h, chunk_sec=float(rng.choice([5.0, 10.0, 30.0])),
                              threads=int(rng.integers(1, 5))))
    return out


def all_cases():
    E = kats.env_params
    cases = []
    cases += kat_frame_cases()
    cases += kat_segment_cases()
    cases.append(synth_case("clip60s_1080p_config0", 0, 1, 1800, E(), threads=4))           # BASELINE configs[0]
    cases.append(synth_case("clip60s_1080p_defaults", 0, 2, 1800, kats.code_defaults(), threads=3))
    cases.append(synth_case("cctv_1080p_600f", 1, 2, 600, E()))
    cases.append(synth_case("batchclip_seed100", 3, 100, 900, E(), chunk_sec=10.0, threads=4))
    cases.append(synth_case("batchclip_seed101", 3, 101, 900, E(), chunk_sec=7.0, threads=3))
    cases.append(synth_case("dense_4k_24f", 2, 3, 24, E()))                                  # configs[2] shape
    cases.append(synth_case("stream_cfg4_720f", 4, 5, 720, E(), chunk_sec=6.0, threads=4))   # configs[4] slice
    cases += random_param_cases()
    # TARGET_FPS frame skipping: the counter restarts at each chunk's seek key frame
    cases.append(synth_case("skip_tfps10_chunk10", 3, 102, 900, E(), chunk_sec=10.0, threads=3, target_fps=10.0))
    cases.append(synth_case("skip_tfps7_chunk2p5", 0, 103, 600, E(), chunk_sec=2.5, threads=4, target_fps=7.0))
    cases.append(synth_case("skip_tfps4_chunk7", 3, 104, 900, kats.code_defaults(), chunk_sec=7.0, threads=2, target_fps=4.0))
    cases.append(synth_case("skip_tfps12p5_720p", 0, 105, 450, E(), 1280, 720, chunk_sec=30.0, threads=1, target_fps=12.5))
    return cases
