#!/usr/bin/env python
"""bench.py — motion-scan hot path on B200: MV records/s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU path on the box's host cores

Workload (config.workload): BASELINE.json configs[4] — a decode-free synthetic AVMotionVector stream
of ~10^9 native 40-byte records per GPU (1080p30 CCTV mix cut into 10-minute videos), generated on the
device by the deterministic generator of include/mvgen_core.h. One step = K-A over every frame of the
stream + K-C over every video. --scaling weak (default): every rank owns a stream of that size
(seed + rank); --scaling strong: ONE stream of --records records, rank g scans frames
[g·F/G, (g+1)·F/G). Videos are independent, so there is no collective on the data path
(torch.distributed only provides the barrier and the max-over-ranks of the timings).

Reported on one JSON line: `value` (device-resident records/s, whole job), `roofline` (K-A's
algorithmic bytes / its CUDA-event duration vs MEASURED_PEAKS.json hbm_gbs), `e2e` (same metric
through the host-facing C ABI with HOST buffers: H2D of every record + D2H of the results inside the
timed region; modes in e2e.modes, the headline is the best mode that starts from native 40-byte host
records), `spec_stream` (K-A on the §8(d) config-5 stream exactly as specified), `cpu_baseline` (the
reference built as oracle/_ref, else the oracle port, on this box's host cores, bounded sample),
`clocks`, `gpu_launches`.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "motion-estimated-video-trimmer_b200"))
sys.path.insert(0, str(ROOT / "tests"))

PRODUCER_MODES = ("producers", "producers_elided", "producers_compact")
METRIC = "mv_records_per_s"
UNIT = "records/s"
REC_BYTES = 40
FRAME_BYTES = 17  # 4 B count + 8 B pts in, 1 B flag + 4 B count out


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(n_rec, n_frames):
    """dram__bytes_read.sum + dram__bytes_write.sum of one K-A launch from the committed `ncu --set full`
    capture (profiles/*ka_traffic.json) — only when that capture was taken on this exact workload."""
    best = None
    for f in sorted((ROOT / "profiles").glob("*ka_traffic.json")):
        try:
            d = json.loads(f.read_text())
            if int(d["records"]) == int(n_rec) and int(d["frames"]) == int(n_frames):
                best = (float(d["traffic"]), f.name)
        except Exception:
            continue
    return best


def rank_cpus(index: int, world: int):
    """The CPUs this rank's producer threads and projection pool may use: the CPUs NVML reports as local to the
    rank's GPU (NUMA affinity), cut into `world` disjoint contiguous shares so that the ranks of one box do not run on
    top of each other (8 ranks x 32 threads on 32 cores was the 12 s `submit` sum of round 1). The process is bound to
    its share, so pinned staging is first-touched there too."""
    cpus = sorted(os.sched_getaffinity(0))
    if world <= 1:
        return cpus
    local = set(cpus)
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        near = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1} & local
        if near:
            local = near
    except Exception:
        pass
    pool = sorted(local)
    # ranks whose GPUs share the same CPU set split it; on a one-node box that is all of them
    per = max(1, len(pool) // world)
    mine = pool[(index % world) * per : (index % world + 1) * per] or pool
    try:
        os.sched_setaffinity(0, set(mine))
    except Exception:
        pass
    return mine


# ------------------------------------------------------------------------------------ clocks ------
class ClockSampler:
    """Samples SM clock + throttle reasons during the timed region (pynvml, else nvidia-smi)."""

    REASONS = {
        0x4: "sw_power_cap",
        0x8: "hw_slowdown",
        0x20: "sw_thermal_slowdown",
        0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown",
    }

    def __init__(self, index: int):
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        self._h = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._h = None

    def _loop(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self._h is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"], "samples": 0}
        return {
            "sm_mhz": float(np.median(self.samples)),
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


# --------------------------------------------------------------------------------- CPU legs -------
def _sample_file(n_bytes):
    """Where the MVS1 sample lives while the reference scans it: /dev/shm when it has room, else the temp dir."""
    import shutil
    import tempfile

    for d in ("/dev/shm", tempfile.gettempdir()):
        try:
            if os.access(d, os.W_OK) and shutil.disk_usage(d).free > n_bytes + (256 << 20):
                return os.path.join(d, f"mscan_bench_{os.getpid()}.mvs")
        except OSError:
            continue
    return os.path.join(tempfile.gettempdir(), f"mscan_bench_{os.getpid()}.mvs")


def cpu_leg(gen, params, spec, n_frames, threads, passes, warmup, repeats=1, sample_data=None, extras=False):
    """The CPU implementation of the path on this box's host cores, on the first n_frames frames of the
    stream. kind "reference": oracle/_ref/ref_scan — the reference's own motion_scanner.cpp / pipeline.cpp
    (built in the build container against the fake-libav shim) — `threads` MotionScanner instances over
    disjoint time ranges through the public scan_range(), timed by the reference's OWN analyze timer
    around check_frame (motion_scanner.cpp:375-380); the step time is the slowest thread's analyze time
    (decode stand-in excluded). `repeats` independent runs of `passes` timed passes each: value = median,
    spread reported. kind "port": the oracle restatement, early-exit semantics, pthreads.
    extras: the same run with the code defaults of config.hpp:57-123 instead of the shipped env, and the
    full-count (no early exit) variant of the oracle port."""
    import oracle_lib as orc
    import ref_runner
    from motionscan import mvgen, mvs_io

    if sample_data is None:
        cnt, off, recs, pts = gen.synth_host(spec, 0, n_frames, n_threads=threads)
    else:
        cnt, off, recs, pts = sample_data
    n_rec = int(off[-1])
    out = {"unit": UNIT, "cores": threads, "repeats": repeats}
    gw, gh, m = orc.geometry(spec.width, spec.height, params.block_size, params.block_shift, params.vertical_mask)

    def port_step(p, th, early):
        cfg = orc.make_cfg(p, gw, gh, m)
        fpv = spec.frames_per_video or n_frames
        flags, _ = orc.scan_frames(cfg, recs, off, early_exit=early, threads=th)
        for a in range(0, n_frames, fpv):
            b = min(n_frames, a + fpv)
            orc.video_tail(pts[a:b], flags[a:b], (b - a) / spec.fps, p.max_gap_sec, p.padding_sec, p.min_savings_pct)
        return flags

    def port_rate(p, th, early, n_pass):
        port_step(p, th, early)
        t0 = time.perf_counter()
        for _ in range(n_pass):
            flags = port_step(p, th, early)
        return n_rec * n_pass / (time.perf_counter() - t0), flags

    if ref_runner.available():
        path = _sample_file(recs.nbytes)
        try:
            mvs_io.write_mvs(path, spec.width, spec.height, int(spec.fps), 1, np.arange(n_frames), cnt, recs)
            dur = n_frames / spec.fps
            chunk = max(1.0, dur / (2 * threads))  # CHUNK_DURATION_SEC: at least 2 chunks per worker thread
            runs = [ref_runner.run(path, params, threads=threads, passes=passes, warmup=warmup, chunk_sec=chunk) for _ in range(max(1, repeats))]
            r1 = ref_runner.run(path, params, threads=1, passes=1, warmup=0, chunk_sec=chunk) if threads > 1 else runs[0]
            rd = ref_runner.run(path, mvgen.code_default_params(), threads=threads, passes=max(1, min(passes, 5)), warmup=1, chunk_sec=chunk) if extras else None
        finally:
            if os.path.exists(path):
                os.unlink(path)
        vals = sorted(n_rec * passes / (r["par_analyze_max_us"] * 1e-6) for r in runs)
        r = runs[0]
        value = float(np.median(vals))
        out.update(
            kind="reference",
            value=value,
            value_min=vals[0],
            value_max=vals[-1],
            values=vals,
            seconds_per_step=n_rec / value,
            value_1thread=n_rec / (r1["analyze_us"] * 1e-6),
            pipeline_value=float(np.median([n_rec * passes / (x["run_wall_us"] * 1e-6) for x in runs])),
            wall_value=float(np.median([n_rec * passes / (x["par_wall_us"] * 1e-6) for x in runs])),
            motion_frames=int(r["par_motion_frames"]),
            sample=f"first {n_frames} frames / {n_rec} records of the stream; reference sources (oracle/_ref): {threads} "
            f"MotionScanner threads over disjoint ranges, {repeats} runs x {passes} timed passes (value = median run), step "
            f"time = slowest thread's own check_frame timer (hot path only, records just copied into that core's cache by "
            f"the decode stand-in); wall_value = the same leg by wall clock, stand-in included; pipeline_value = "
            f"ProcessingPipeline::run() wall incl. mmap + shim demux",
        )
        if rd is not None:
            out["value_code_defaults"] = n_rec * max(1, min(passes, 5)) / (rd["par_analyze_max_us"] * 1e-6)
            out["code_defaults"] = "config.hpp:57-123: MV_THRESHOLD_SQ=16 VECTORS_NEEDED=2 CLUSTERS_NEEDED=2 (second sweep of SURVEY §8(d))"
    else:
        vals = sorted(port_rate(params, threads, True, passes)[0] for _ in range(max(1, repeats)))
        value = float(np.median(vals))
        v1, flags = port_rate(params, 1, True, 1)
        out.update(
            kind="port",
            value=value,
            value_min=vals[0],
            value_max=vals[-1],
            values=vals,
            seconds_per_step=n_rec / value,
            value_1thread=v1,
            motion_frames=int(flags.sum()),
            sample=f"first {n_frames} frames / {n_rec} records of the stream; oracle port (oracle/_ref not built), "
            f"early-exit semantics + tail, {threads} pthreads, {repeats} runs x {passes} timed passes (value = median run)",
        )
        if extras:
            out["value_code_defaults"] = port_rate(mvgen.code_default_params(), threads, True, max(1, min(passes, 5)))[0]
    if extras:  # full cluster count, no early exit (what the CUDA path reports per frame): oracle port, same threads
        out["full_count_value"] = port_rate(params, threads, False, max(1, min(passes, 5)))[0]
        out["full_count_kind"] = "port (the reference cannot run without its early exit, motion_scanner.cpp:288-289)"
    return out, n_rec


def sample_frames(args, spec, gen, threads):
    """Frames of the host-resident sample shared by BOTH arms (e2e of the CUDA path, the reference arm, cpu_baseline):
    the leading frames of the stream holding about --e2e-records records."""
    if args.e2e_frames:
        return args.e2e_frames
    probe = 256
    cnt = np.zeros(probe, np.uint32)
    gen.lib().mscan_synth_host_counts(C.byref(spec), 0, probe, cnt.ctypes.data, threads)
    return max(1, int(np.ceil(args.e2e_records / max(cnt.mean(), 1.0))))


def run_reference(args):
    """--impl reference: the CPU implementation of the path on this box's host cores, all threads. Maps nothing of the
    product: the stream comes from libmvgen.so (the generator alone), the scan runs in oracle/_ref/ref_scan."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from motionscan import mvgen

    threads = os.cpu_count() or 1
    params = mvgen.shipped_env_params()
    global _WORKLOAD_DESC
    preset, seed, _fixed, _WORKLOAD_DESC = WORKLOADS[args.workload]
    spec = mvgen.synth_preset(preset, seed)
    n_frames = args.cpu_frames or sample_frames(args, spec, mvgen, threads)
    if _fixed:
        n_frames = min(n_frames, _fixed)
    cpu, n_rec = cpu_leg(mvgen, params, spec, n_frames, threads, args.steps, args.warmup, repeats=args.ref_repeats)
    # this arm is the reference's CPU code: nothing of the product may be mapped into it
    maps = open("/proc/self/maps").read() if os.path.exists("/proc/self/maps") else ""
    assert "libmotionscan.so" not in maps and "libmscan_feed.so" not in maps, "the reference arm must not load the product library"
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": cpu["value"],
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": cpu["seconds_per_step"] * 1e3,
        "higher_is_better": True,
        "scaling": args.scaling,
        "vs_baseline": None,
        "dtype": "int32",
        "data": "synthetic",
        "config": workload_config(spec, n_frames, n_rec, "host", args),
        "frames_per_s": n_frames / cpu["seconds_per_step"],
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


WORKLOADS = {
    # name: (mvgen preset, seed, frames (None → sized by --records), description)
    "stream1e9": (4, 5, None, "mvstream_1e9: decode-free synthetic AVMotionVector stream (BASELINE.json configs[4]), "
                  "1080p30 CCTV mix cut into 10-min videos, native 40-B records; kept as the headline because it exercises "
                  "the whole path (active frames, clusters, segments, decisions) — the stream exactly as SURVEY §8(d) "
                  "specifies it (uniform dst, 16 320 rec/frame) has almost no active frame and is measured beside it as spec_stream"),
    "stream1e9_spec": (5, 5, None, "mvstream_1e9_spec: SURVEY §8(d) config 5 as specified — 16 320 records per 1080p frame, dst "
                       "uniform over MB sub-centres, 10 % moving with d uniform in [-8,8]^2, 0.1 % out-of-frame dst, native 40-B records"),
    "cctv10min": (1, 2, 18000, "cctv10min: synthetic 10 min 1080p30 CCTV-style clip (BASELINE.json configs[1]) as an MV stream"),
    "dense4k": (2, 3, 3600, "dense4k: synthetic 2 min 4K30 clip with a dense 8x8 MV field, 129 600 records per P-frame "
                "(BASELINE.json configs[2]) as an MV stream"),
    "batch64": (3, 100, 64 * 1800, "batch64: 64 synthetic 1080p30 clips of 30–120 s (uniform; seeds 100…163) in one batch (BASELINE.json "
                "configs[3], SURVEY §8(d) config 4) as MV streams"),
}
_WORKLOAD_DESC = WORKLOADS["stream1e9"][3]


def workload_config(spec, n_frames, n_rec, where, args=None, sample=None):
    cfg = {
        "workload": _WORKLOAD_DESC,
        "resident": where,
        "frames": int(n_frames),
        "records": int(n_rec),
        "frames_per_video": int(spec.frames_per_video),
        "params": "config/motion_trim.env: MV_THRESHOLD_SQ=4 VECTORS_NEEDED=4 CLUSTERS_NEEDED=2 VERTICAL_MASK=0.05 MAX_GAP_SEC=5 PADDING_SEC=0.5 MIN_SAVINGS_PCT=5",
        "l2": "inputs larger than L2 (no flush needed)",
    }
    if args is not None:
        cfg["host_sample"] = (f"the leading frames of the stream holding ~{args.e2e_records:.3g} records: what e2e (CUDA path, host-fed) "
                              "and --impl reference both scan")
    if sample is not None:
        cfg["host_sample_frames"], cfg["host_sample_records"] = int(sample[0]), int(sample[1])
    elif where == "host":
        cfg["host_sample_frames"], cfg["host_sample_records"] = int(n_frames), int(n_rec)
    return cfg


# ---------------------------------------------------------------------------------- GPU arm -------
def build_stream(ctx, ms, spec, frame0, n_frames, sh, stream):
    """Generates frames [frame0, frame0+n_frames) of the stream on the device. Returns the device buffers and the
    host copy of the record offsets."""
    d_cnt = ctx.dev_alloc(4 * max(n_frames, 1))
    d_off = ctx.dev_alloc(8 * (n_frames + 1))
    ctx.synth_counts(spec, frame0, n_frames, d_cnt, sh)
    ctx.offsets_from_counts(d_cnt, n_frames, d_off, sh)
    stream.synchronize()
    off = np.zeros(n_frames + 1, np.uint64)
    ctx.d2h(off, d_off)
    n_rec = int(off[-1])
    d_recs = ctx.dev_alloc(REC_BYTES * n_rec + 256)
    d_pts = ctx.dev_alloc(8 * max(n_frames, 1))
    ctx.synth_fill(spec, frame0, n_frames, d_off, d_recs, d_pts, sh)
    stream.synchronize()
    ctx.dev_free(d_cnt)
    return d_off, d_recs, d_pts, off, n_rec


def batch64_lengths(seed0, n_videos=64, fps=30):
    """durations uniform in 30–120 s (a fixed pseudo-random function of the video's seed)."""
    out = []
    for v in range(n_videos):
        z = (seed0 + v) * 0x9E3779B97F4A7C15 & 0xFFFFFFFFFFFFFFFF
        z ^= z >> 29
        out.append(30 * fps + int(z % (90 * fps + 1)))
    return out


def build_batch(ctx, ms, preset, seed0, lengths, lo, hi, sh, stream):
    """Frames [lo, hi) of the concatenation of len(lengths) videos (video v: preset with seed seed0 + v, lengths[v]
    frames), generated on the device video by video. Returns like build_stream plus the global video start frames."""
    starts = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    n_frames = hi - lo
    d_cnt = ctx.dev_alloc(4 * max(n_frames, 1))
    d_off = ctx.dev_alloc(8 * (n_frames + 1))
    pieces = []
    for v, L in enumerate(lengths):
        a, b = max(lo, int(starts[v])), min(hi, int(starts[v]) + L)
        if a >= b:
            continue
        spec = ms.synth_preset(preset, seed0 + v)
        spec.frames_per_video = L
        pieces.append((spec, a - int(starts[v]), b - a, a - lo))
        ctx.synth_counts(spec, a - int(starts[v]), b - a, d_cnt + 4 * (a - lo), sh)
    ctx.offsets_from_counts(d_cnt, n_frames, d_off, sh)
    stream.synchronize()
    off = np.zeros(n_frames + 1, np.uint64)
    ctx.d2h(off, d_off)
    n_rec = int(off[-1])
    d_recs = ctx.dev_alloc(REC_BYTES * n_rec + 256)
    d_pts = ctx.dev_alloc(8 * max(n_frames, 1))
    for spec, f0, n, at in pieces:
        ctx.synth_fill(spec, f0, n, d_off + 8 * at, d_recs, d_pts + 8 * at, sh)
    stream.synchronize()
    ctx.dev_free(d_cnt)
    return d_off, d_recs, d_pts, off, n_rec, starts


def frames_for_records(ctx, spec, records, sh, stream):
    probe = 4096
    d_probe = ctx.dev_alloc(4 * probe)
    ctx.synth_counts(spec, 0, probe, d_probe, sh)
    stream.synchronize()
    pc = np.zeros(probe, np.uint32)
    ctx.d2h(pc, d_probe)
    ctx.dev_free(d_probe)
    return int(np.ceil(records / max(pc.mean(), 1.0)))


def run_gpu(args):
    import torch

    import motionscan as ms

    from motionscan.dist import Dist, strong_share, throughput, video_pieces

    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the motion-scan path has no CPU fallback")
    torch.cuda.set_device(local)
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    my_cpus = rank_cpus(local, world_env)  # before any pinned allocation: staging is first-touched on these cores
    D = Dist("nccl", torch.device("cuda", local))
    world, rank = D.world, D.rank
    barrier, allmax, allsum = D.barrier, D.allmax, D.allsum

    params = ms.shipped_env_params()
    ctx = ms.Context(local, params, max_log_frames=1 << 20, slab_bytes=args.slab_mb << 20)
    ctx.reserve_staging()  # the pinned ring now: allocated lazily, a slab's 25 ms would land inside a timed host-fed step
    stream = torch.cuda.Stream()
    sh = stream.cuda_stream

    # ---- build the device-resident stream ---------------------------------------------------------
    global _WORKLOAD_DESC
    preset, seed, fixed_frames, _WORKLOAD_DESC = WORKLOADS[args.workload]
    strong = args.scaling == "strong"
    spec = ms.synth_preset(preset, seed if strong else seed + rank)
    batch_lengths = batch64_lengths(seed if strong else seed + 64 * rank) if args.workload == "batch64" else None
    if batch_lengths is not None:
        total_frames_all = int(sum(batch_lengths))
    else:
        total_frames_all = fixed_frames if fixed_frames is not None else frames_for_records(ctx, spec, args.records, sh, stream)
    if strong:  # one stream, rank g scans frames [g·F/G, (g+1)·F/G)  
        frame0, n_frames = strong_share(total_frames_all, world, rank)
    else:
        frame0, n_frames = 0, total_frames_all
    fpv = spec.frames_per_video
    if batch_lengths is not None:  # 64 videos of 30–120 s, seeds 100…163
        d_off, d_recs, d_pts, off, n_rec, vstarts = build_batch(ctx, ms, preset, seed if strong else seed + 64 * rank, batch_lengths,
                                                                frame0, frame0 + n_frames, sh, stream)
    else:
        d_off, d_recs, d_pts, off, n_rec = build_stream(ctx, ms, spec, frame0, n_frames, sh, stream)
        vstarts = np.arange(0, frame0 + n_frames + fpv, fpv, dtype=np.int64) if fpv else np.array([0], np.int64)
    d_flags = ctx.dev_alloc(n_frames)
    d_counts = ctx.dev_alloc(4 * n_frames)
    d_segs = ctx.dev_alloc(16 * n_frames)

    def video_offsets(f0, n):
        """Local frame offsets of the videos (pieces of videos at the ends of a strong-scaling share) in [f0, f0+n)."""
        return np.array(video_pieces(vstarts, f0, n), dtype=np.uint64)

    voff = video_offsets(frame0, n_frames)
    n_videos = len(voff) - 1
    durations = np.diff(voff).astype(np.float64) / spec.fps
    d_res = ctx.dev_alloc(40 * n_videos)
    geom = ms.geometry_from_dims(params, spec.width, spec.height)

    def step():
        ctx.scan_device(d_recs, d_off, None, [geom], n_frames, d_flags, d_counts, sh)
        ctx.segments_device(voff, durations, d_pts, d_flags, d_segs, d_res, sh)

    for _ in range(args.warmup):
        step()
    stream.synchronize()
    ctx.reset_stats()
    ctx.set_profiling(True)
    sampler = ClockSampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    stream.synchronize()
    torch.cuda.synchronize()
    sampler.stop()
    barrier()
    ms_total = allmax(e0.elapsed_time(e1))
    st = ctx.stats()
    ctx.set_profiling(False)
    ka_ms = st.scan_ms / max(st.scan_launches, 1)
    kc_ms = st.segment_ms / max(st.segment_launches, 1)
    launches = int(st.scan_launches + st.segment_launches)
    total_rec = allsum(float(n_rec))
    total_frames = allsum(float(n_frames))
    value = throughput(total_rec, args.steps, ms_total)

    # sanity on the results of the last step (not timed): flags must be a mix, every video decided
    flags = np.zeros(n_frames, np.uint8)
    res = np.zeros(n_videos, ms.RESULT_DTYPE)
    ctx.d2h(flags, d_flags)
    ctx.d2h(res, d_res)

    # ---- PCIe roofline denominator: pinned H2D copy, 1 GiB, best of 8 -----------------
    pin = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
    dev = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    pcie_gbs = 0.0
    with torch.cuda.stream(stream):
        for _ in range(9):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            dev.copy_(pin, non_blocking=True)
            b.record(stream)
            stream.synchronize()
            pcie_gbs = max(pcie_gbs, (1 << 30) / (a.elapsed_time(b) * 1e-3) / 1e9)
    del pin, dev
    pcie_gbs_alone = pcie_gbs
    pcie_gbs = allmax(pcie_gbs)

    # ---- K-A on projected records, device-resident (what the host-fed default launches) ----------------
    packed = None
    if not args.no_packed:
        d_r8 = ctx.dev_alloc(8 * n_rec + 256)
        ctx.pack_records_device(d_recs, n_rec, d_r8, sh)
        stream.synchronize()
        d_flags8 = ctx.dev_alloc(n_frames)
        d_counts8 = ctx.dev_alloc(4 * n_frames)
        for _ in range(3):
            ctx.scan_device_packed(d_r8, d_off, None, [geom], n_frames, d_flags8, d_counts8, sh)
        stream.synchronize()
        ctx.reset_stats()
        ctx.set_profiling(True)
        for _ in range(args.steps):
            ctx.scan_device_packed(d_r8, d_off, None, [geom], n_frames, d_flags8, d_counts8, sh)
        stream.synchronize()
        pst = ctx.stats()
        ctx.set_profiling(False)
        pk_ms = allmax(pst.scan_ms / max(pst.scan_launches, 1))
        flags8 = np.zeros(n_frames, np.uint8)
        ctx.d2h(flags8, d_flags8)
        peak8, _ = peaks()
        packed = {
            "kernel": "ka_scan_kernel<packed>",
            "ms_per_launch": pk_ms,
            "records_per_s": allsum(float(n_rec)) / (pk_ms * 1e-3),
            "bytes_per_record": 8,
            "achieved_gbs": (8 * n_rec + FRAME_BYTES * n_frames) / (pk_ms * 1e-3) / 1e9,
            "frac_of_hbm_peak": (8 * n_rec + FRAME_BYTES * n_frames) / (pk_ms * 1e-3) / 1e9 / peak8,
            "matches_native": bool(np.array_equal(flags8, flags)),
        }
        for d in (d_r8, d_flags8, d_counts8):
            ctx.dev_free(d)

    # ---- e2e: host-fed through the C ABI ------------------------------------------------------------
    # host-resident sample: the leading frames of this rank's stream, bounded by --e2e-records (strong scaling: the
    # sample is cut G ways like the stream). Ways for host records to reach the GPU (include/motionscan.h):
    #   producers       T decode-worker stand-ins (csrc/feed_harness.cpp), one per CPU of the rank: each writes a frame's
    #                   native 40-B records into its own side-data buffer (cache-hot, pageable — what sd->data is right
    #                   after avcodec_receive_frame) and calls mscan_submit per frame, concurrently; the library projects
    #                   the 8 bytes the path reads into its pinned ring outside its mutex and DMAs 8 B/record. WALL CLOCK,
    #                   the stand-in's own work included.
    #   native_inplace  native 40-B records DMA'd straight out of the caller's pinned buffer (no host pass)
    #   projected       whole videos of native records resident in host DRAM, one mscan_submit each: the library's
    #                   worker pool projects them (bound by reading 40 B/record from DRAM)
    #   packed_pinned   the caller hands over records it projected itself beforehand; only DMA + kernels are timed
    #   elided_pinned   same with the caller's own static-elided encoding (mscan_elide_records → mscan_submit_elided)
    #   compact_pinned  same with the caller's own compaction (mscan_compact_records: moving records only → mscan_submit_packed)
    # producers runs three times: `producers` with mscan_mv8 on the wire (STAGING_PACK, 8 B/record), `producers_elided` in the
    # lossless static-elided form (STAGING_ELIDE, ~4.3 B/record) and `producers_compact` as the library does it by default
    # (STAGING_AUTO: T² > 0 here, so a decode thread's frame travels as its moving records only).
    # `e2e.value` is the best of the modes that start from native host records (producers*, native_inplace, projected).
    # the same frame count the reference arm derives (sample_frames): both arms scan the same host sample
    e2e_frames = min(sample_frames(args, spec, ms, max(1, len(my_cpus))), n_frames)
    if strong:
        e2e_frames = max(1, e2e_frames // world)
    e_rec = int(off[e2e_frames])
    h_recs = ctx.pinned_array(e_rec, ms.MV_DTYPE)
    h_r8 = ctx.pinned_array(e_rec, ms.MV8_DTYPE)
    h_pts = ctx.pinned_array(e2e_frames, np.float64)
    e_off = np.ascontiguousarray(off[: e2e_frames + 1])
    h_cnt = np.diff(e_off).astype(np.uint32)
    ctx.d2h(h_recs, d_recs)
    ctx.d2h(h_pts, d_pts)
    ms.pack_records(h_recs, h_r8)
    e_voff = video_offsets(frame0, e2e_frames)
    e_vids = list(range(len(e_voff) - 1))
    e_durs = [float(e_voff[v + 1] - e_voff[v]) / spec.fps for v in e_vids]
    n_prod = args.feed_threads or len(my_cpus)
    pack_threads = max(1, len(my_cpus))
    ctx.set_pack_threads(pack_threads)
    src8 = np.array(h_r8)  # the stand-in decoders' compact input lives in ordinary pageable memory
    # elided_pinned: the caller's own static-elided encoding of the sample in pinned memory (link-saturated run of that form)
    h_enc = z_off = z_te = z_tiles = None
    if "elided_pinned" in args.e2e_modes.split(","):
        z_bound = sum(int(ms.lib().mscan_elide_bound(int(x))) for x in h_cnt)
        h_enc = ctx.pinned_array(z_bound + 16, np.uint8)
        z_enc, z_off, z_te = ms.elide_frames(h_recs, e_off, out=h_enc)
        z_tiles = np.concatenate([[0], np.cumsum((h_cnt.astype(np.int64) + 1023) // 1024)])
    # compact_pinned: the caller's own compaction of the sample (moving records only) in pinned memory
    h_cmp = m_cnt = m_off = None
    if "compact_pinned" in args.e2e_modes.split(","):
        h_cmp = ctx.pinned_array(e_rec + 8, ms.MV8_DTYPE)
        _, m_cnt = ms.compact_frames(h_recs, e_off, out=h_cmp)
        m_off = np.concatenate([[0], np.cumsum(m_cnt.astype(np.int64))])

    def tail(vids):
        out = [ctx.collect(v) for v in vids]
        segs = ctx.segments_batch(vids, e_durs)
        for v in vids:
            ctx.video_close(v)
        return out, segs

    feed_stats = []

    def e2e_step(mode):
        for v in e_vids:
            ctx.video_open(v, spec.width, spec.height)
        if mode in PRODUCER_MODES:
            fr, index = ms.feed_run(ctx, e_vids, e_voff, h_pts, h_cnt, e_off, src8, n_threads=n_prod, cpus=my_cpus[:n_prod] if len(my_cpus) >= n_prod else None,
                                    frames_per_submit=args.feed_batch, submit_kind=0)
            t_tail = time.perf_counter()
            out, segs = tail(e_vids)
            feed_stats.append((fr.wall_s, fr.hot_max_s, fr.hot_sum_s, fr.standin_max_s, fr.standin_sum_s, time.perf_counter() - t_tail, fr.submits))
            # collect() returns submission order; put every video's flags back into frame order
            ordered = []
            for v in e_vids:
                a, b = int(e_voff[v]), int(e_voff[v + 1])
                fl = np.zeros(b - a, np.uint8)
                cn = np.zeros(b - a, np.uint32)
                idx = index[a:b].astype(np.int64)
                fl[:] = out[v][0][idx]
                cn[:] = out[v][1][idx]
                ordered.append((fl, cn))
            return ordered, segs
        for v in e_vids:
            a, b = int(e_voff[v]), int(e_voff[v + 1])
            if mode == "elided_pinned":
                ctx.submit_elided(v, h_pts[a:b], h_cnt[a:b], h_enc, z_off[a : b + 1], z_te[int(z_tiles[a]) : int(z_tiles[b])])
            elif mode == "compact_pinned":
                ctx.submit_packed_raw(v, b - a, h_pts.ctypes.data + 8 * a, m_cnt.ctypes.data + 4 * a, h_cmp.ctypes.data + 8 * int(m_off[a]))
            elif mode == "packed_pinned":
                ctx.submit_packed_raw(v, b - a, h_pts.ctypes.data + 8 * a, h_cnt.ctypes.data + 4 * a, h_r8.ctypes.data + 8 * int(e_off[a]))
            else:
                ctx.submit_raw(v, b - a, h_pts.ctypes.data + 8 * a, h_cnt.ctypes.data + 4 * a, h_recs.ctypes.data + REC_BYTES * int(e_off[a]))
        return tail(e_vids)

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    modes = {}
    seg_ref = None
    all_modes = ("producers", "producers_elided", "producers_compact", "native_inplace", "projected", "packed_pinned", "elided_pinned",
                 "compact_pinned")
    run_modes = [m for m in all_modes if m in args.e2e_modes.split(",")] or list(all_modes)
    for mode in run_modes:
        # producers: mscan_mv8 on the wire (STAGING_PACK); producers_elided: the lossless static-elided form (STAGING_ELIDE);
        # producers_compact: the library's default for a decode thread's pageable frame (STAGING_AUTO → moving records only)
        ctx.set_staging_mode({"projected": ms.STAGING_PACK, "producers": ms.STAGING_PACK, "producers_elided": ms.STAGING_ELIDE}.get(mode, ms.STAGING_AUTO))
        for _ in range(2):
            e2e_out = e2e_step(mode)
        ctx.sync()
        ctx.reset_stats()
        feed_stats.clear()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_out = e2e_step(mode)
        ctx.sync()
        dt = allmax(time.perf_counter() - t0)
        est = ctx.stats()
        e_flags = np.concatenate([o[0] for o in e2e_out[0]])
        seg_bytes = e2e_out[1][0].tobytes() + e2e_out[1][2].tobytes()
        seg_ref = seg_ref or seg_bytes
        modes[mode] = {
            "value": allsum(float(e_rec)) * e2e_steps / dt,
            "h2d_bytes_per_step": int(est.h2d_bytes // e2e_steps),
            "d2h_bytes_per_step": int(est.d2h_bytes // e2e_steps),
            "h2d_gbs": est.h2d_bytes / dt / 1e9,
            "launches": int(est.scan_launches + est.segment_launches),
            # the host-fed results must equal the device-resident ones for the same frames, and every mode's segments agree
            "matches_device_resident": bool(np.array_equal(e_flags, flags[:e2e_frames])) and seg_bytes == seg_ref,
        }
        if mode == "projected" or mode in PRODUCER_MODES:
            modes[mode]["host_threads"] = pack_threads if mode == "projected" else n_prod
            modes[mode]["project_cpu_ms_per_step"] = est.project_ms / e2e_steps
            modes[mode]["records_projected_per_step"] = int(est.records_projected // e2e_steps)
        if mode in ("producers_elided", "producers_compact"):
            modes[mode]["wire_bytes_per_record"] = est.elided_bytes / max(est.records_elided, 1)
        if mode in PRODUCER_MODES:
            fs = np.array(feed_stats)
            wall, hot_max, hot_sum, sd_max, sd_sum, t_tail, submits = fs.sum(axis=0)
            # the reference arm's protocol on this arm: records / (slowest producer's time inside mscan_submit + the tail)
            hot_rate = allsum(float(e_rec)) * e2e_steps / allmax(hot_max + t_tail)
            modes[mode].update(
                timing="wall clock over the producers (decode stand-in + mscan_submit per frame) and the tail (collect, segments_batch, close)",
                frames_per_submit=args.feed_batch,
                submits_per_step=int(submits // e2e_steps),
                standin_share_of_thread_time=float(sd_sum / max(sd_sum + hot_sum, 1e-12)),
                submit_share_of_thread_time=float(hot_sum / max(sd_sum + hot_sum, 1e-12)),
                tail_ms_per_step=float(t_tail / e2e_steps * 1e3),
                hot_path_only_value=hot_rate,
                hot_path_only_how="records / (slowest producer's time inside mscan_submit + tail): the reference arm's own timer protocol "
                "(motion_scanner.cpp:375-380); not the headline — work hidden behind the stand-in is not counted by it",
            )
    ctx.set_staging_mode(ms.STAGING_AUTO)
    # The headline: what the path sustains from cache-hot native records in the decode threads' own buffers — the
    # footing of the reference arm, whose timer runs around check_frame only, on records its decode stand-in has just
    # copied into that core's cache. The producers run is timed by that same protocol (hot_path_only_value); because
    # H2D copies and kernels run asynchronously behind the stand-in, that figure alone could exceed what one PCIe link
    # can carry, so it is capped by the rate measured in this same run with the link saturated (packed_pinned: the same
    # records already projected, DMA + K-A + tail by wall clock). Both terms and the plain wall-clock figure of the
    # producers run (stand-in included) are in the line.
    for name in PRODUCER_MODES:
        if name not in modes:
            continue
        pm = modes[name]
        cap = modes["packed_pinned"]["value"] if "packed_pinned" in modes else world * pcie_gbs * 1e9 / 8.0
        cap_how = "the rate measured in this run with the link saturated (packed_pinned: 8 B/record, DMA + K-A + tail by wall clock)"
        if name == "producers_elided" and "elided_pinned" in modes:
            cap = modes["elided_pinned"]["value"]
            cap_how = ("the rate measured in this run with the link saturated in the same wire form (elided_pinned: the caller's own "
                       "static-elided encoding in pinned memory, DMA + K-A + tail by wall clock)")
        elif name == "producers_compact" and "compact_pinned" in modes:
            cap = modes["compact_pinned"]["value"]
            cap_how = ("the rate measured in this run with the caller's own compaction of the same sample in pinned memory (compact_pinned: "
                       "only the moving records cross the link; DMA + K-A + tail by wall clock — the tail dominates it)")
        elif name != "producers":  # fewer bytes per record on the same link: the measured link-bound rate scales with the wire size
            cap = cap * 8.0 / max(pm["wire_bytes_per_record"], 1e-9)
            cap_how += f", scaled by 8 / {pm['wire_bytes_per_record']:.2f} B per record measured on the wire in this form"
        pm["value_wall_incl_decode_standin"] = pm["value"]
        pm["link_bound_value"] = cap
        pm["value"] = min(pm["hot_path_only_value"], cap)
        pm["value_how"] = ("min(hot_path_only_value, link_bound_value): records / (slowest decode thread's time inside mscan_submit + tail), "
                           "capped by " + cap_how + "; value_wall_incl_decode_standin is the same run by wall clock with the decode "
                           "stand-in's own work (writing 40 B/record into the side-data buffer) included")
    e2e_mode = max([m for m in PRODUCER_MODES + ("native_inplace", "projected") if m in modes] or list(modes),
                   key=lambda m: modes[m]["value"])
    best = modes[e2e_mode]
    e2e_value, e2e_launches, e2e_ok = best["value"], best["launches"], all(m["matches_device_resident"] for m in modes.values())

    # ---- CPU baseline on rank 0, N=1 only ------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle_lib as orc

        threads = os.cpu_count() or 1
        c_frames = e2e_frames  # the same host sample the e2e modes scan
        if args.cpu_frames:
            c_frames = min(args.cpu_frames, e2e_frames)
        c_off = off[: c_frames + 1]
        c_recs = h_recs[: int(c_off[-1])]
        c_pts = h_pts[:c_frames]
        c_cnt = np.diff(c_off).astype(np.uint32)
        cpu, _ = cpu_leg(ms, params, spec, c_frames, threads, 3, 1, repeats=3, sample_data=(c_cnt, c_off, c_recs, c_pts), extras=True)
        # parity of the sample while we are here: oracle flags == GPU flags
        gw, gh, m = orc.geometry(spec.width, spec.height, params.block_size, params.block_shift, params.vertical_mask)
        of, _ = orc.scan_frames(orc.make_cfg(params, gw, gh, m), c_recs, c_off, threads=threads)
        cpu["parity_with_gpu"] = bool(np.array_equal(of, flags[:c_frames])) and cpu["motion_frames"] == int(of.sum())

    # ---- the §8(d) config-5 stream exactly as specified, device-resident (K-A only) ---------------------
    spec_stream = None
    if args.workload == "stream1e9" and not args.no_spec_stream:
        for d in (d_recs, d_off, d_pts, d_flags, d_counts, d_segs):
            ctx.dev_free(d)
        sspec = ms.synth_preset(5, 5 + (0 if strong else rank))
        s_frames_all = frames_for_records(ctx, sspec, args.records, sh, stream)
        s_f0 = s_frames_all * rank // world if strong else 0
        s_n = (s_frames_all * (rank + 1) // world - s_f0) if strong else s_frames_all
        s_off, s_recs, s_pts, s_hoff, s_nrec = build_stream(ctx, ms, sspec, s_f0, s_n, sh, stream)
        s_flags = ctx.dev_alloc(s_n)
        s_counts = ctx.dev_alloc(4 * s_n)
        for _ in range(3):
            ctx.scan_device(s_recs, s_off, None, [geom], s_n, s_flags, s_counts, sh)
        stream.synchronize()
        ctx.reset_stats()
        ctx.set_profiling(True)
        for _ in range(max(3, min(args.steps, 10))):
            ctx.scan_device(s_recs, s_off, None, [geom], s_n, s_flags, s_counts, sh)
        stream.synchronize()
        sst = ctx.stats()
        ctx.set_profiling(False)
        s_ms = allmax(sst.scan_ms / max(sst.scan_launches, 1))
        sf = np.zeros(s_n, np.uint8)
        ctx.d2h(sf, s_flags)
        # parity spot check of the first frames against the oracle
        import oracle_lib as orc

        k = min(64, s_n)
        hk = np.zeros(int(s_hoff[k]), ms.MV_DTYPE)
        ctx.d2h(hk, s_recs)
        gw, gh, m = orc.geometry(sspec.width, sspec.height, params.block_size, params.block_shift, params.vertical_mask)
        of, _ = orc.scan_frames(orc.make_cfg(params, gw, gh, m), hk, s_hoff[: k + 1])
        peak_s, _ = peaks()
        s_bytes = REC_BYTES * s_nrec + FRAME_BYTES * s_n
        spec_stream = {
            "workload": WORKLOADS["stream1e9_spec"][3],
            "frames": int(allsum(float(s_n))),
            "records": int(allsum(float(s_nrec))),
            "ms_per_launch": s_ms,
            "records_per_s": allsum(float(s_nrec)) / (s_ms * 1e-3),
            "achieved_gbs": s_bytes / (s_ms * 1e-3) / 1e9,
            "frac_of_hbm_peak": s_bytes / (s_ms * 1e-3) / 1e9 / peak_s,
            "active_frames": int(sf.sum()),
            "oracle_spot_check": bool(np.array_equal(of, sf[:k])),
        }

    if rank == 0:
        peak, peak_src = peaks()
        ka_bytes = REC_BYTES * n_rec + FRAME_BYTES * n_frames
        achieved = ka_bytes / (ka_ms * 1e-3) / 1e9 if ka_ms > 0 else 0.0
        traffic = measured_traffic(n_rec, n_frames)
        line = {
            "metric": METRIC,
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps,
            "higher_is_better": True,
            "scaling": args.scaling,
            "vs_baseline": None,
            "dtype": "int32",
            "data": "synthetic",
            "config": workload_config(spec, n_frames, n_rec, "hbm", args, sample=(e2e_frames, e_rec)),
            "frames_per_s": total_frames * args.steps / (ms_total * 1e-3),
            "roofline": {
                "bound": "hbm",
                "kernel": "ka_scan_kernel",
                "achieved": achieved,
                "peak": peak,
                "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": traffic[0] if traffic else None,
                "traffic_source": traffic[1] if traffic else None,
                "peak_source": peak_src,
                "bytes_per_launch": ka_bytes,
                "ms_per_launch": ka_ms,
                "kc_ms_per_launch": kc_ms,
            },
            "e2e": {
                "value": e2e_value,
                "unit": UNIT,
                "h2d_bytes_per_step": best["h2d_bytes_per_step"],
                "d2h_bytes_per_step": best["d2h_bytes_per_step"],
                "mode": e2e_mode,
                "steps": e2e_steps,
                "frames": e2e_frames,
                "records": e_rec,
                "h2d_gbs": best["h2d_gbs"],
                "pcie_peak_gbs": pcie_gbs,
                "pcie_peak_gbs_rank0_alone": pcie_gbs_alone,
                "pcie_frac": best["h2d_gbs"] / pcie_gbs if pcie_gbs > 0 else None,
                "pcie_peak_how": "pinned cudaMemcpyAsync H2D, 1 GiB, best of 9, CUDA events (per GPU)",
                "link_limit_records_per_s": world * pcie_gbs * 1e9 / 8.0,
                "host_cpus_per_rank": len(my_cpus),
                "launches": e2e_launches,
                "matches_device_resident": e2e_ok,
                "modes": modes,
                "value_wall_incl_decode_standin": best.get("value_wall_incl_decode_standin"),
                "how": "per step: mscan_video_open, the mode's submits of native 40-B host records, collect, segments_batch, close; max over "
                "ranks; value = best of producers / producers_elided / producers_compact (decode-worker stand-ins submitting cache-hot frames per "
                "frame, concurrently; 8 B/record over PCIe, ~4.3 B/record in the lossless static-elided form, or — the library's default while "
                "MV_THRESHOLD_SQ > 0 — only the moving records, since a record with src == dst cannot pass motion_scanner.cpp:251; timed like the "
                "reference arm — see modes.*.value_how), native_inplace (pinned records DMA'd in place, 40 B/record, "
                "wall clock) and projected (whole videos from host DRAM through the library's pool, wall clock); packed_pinned "
                "(caller-projected records) is reported in modes only. When a mode flattens with more GPUs the saturated resource is the "
                "host (its cores and DRAM): link_limit_records_per_s is what the measured PCIe links could carry at 8 B/record",
            },
            "packed_kernel": packed,
            "spec_stream": spec_stream,
            "cpu_baseline": cpu,
            "clocks": sampler.summary(),
            "gpu_launches": launches,
            "results": {
                "active_frames": int(flags.sum()),
                "videos": int(n_videos),
                "cut": int((res["decision"] == ms.CUT).sum()),
                "full_copy": int((res["decision"] == ms.FULL_COPY).sum()),
                "no_motion": int((res["decision"] == ms.NO_MOTION).sum()),
                "segments": int(res["n_segments"].sum()),
            },
        }
        print(json.dumps(line), flush=True)
    ctx.host_free(h_recs.ctypes.data)
    if h_enc is not None:
        ctx.host_free(h_enc.ctypes.data)
    if h_cmp is not None:
        ctx.host_free(h_cmp.ctypes.data)
    ctx.host_free(h_r8.ctypes.data)
    ctx.host_free(h_pts.ctypes.data)
    ctx.close()
    D.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="stream1e9", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --records per GPU (seed + rank); strong: --records in total, rank g scans frames [g·F/G,(g+1)·F/G)")
    ap.add_argument("--records", type=float, default=1e9, help="records of the device-resident stream (per GPU when weak, in total when strong)")
    ap.add_argument("--e2e-records", type=float, default=6e7, help="records of the host-resident sample both arms scan (~2.4 GB pinned)")
    ap.add_argument("--e2e-frames", type=int, default=0, help="override: frames of the host sample")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--e2e-modes", default="producers,producers_elided,producers_compact,native_inplace,projected,packed_pinned,elided_pinned,compact_pinned", help="experiments: subset of the e2e modes to run")
    ap.add_argument("--feed-batch", type=int, default=1, help="frames per mscan_submit of the producer stand-ins (1 = per frame, like check_frame)")
    ap.add_argument("--feed-threads", type=int, default=0, help="producer stand-in threads per rank (0 = one per CPU of the rank)")
    ap.add_argument("--ref-repeats", type=int, default=5, help="--impl reference: independent runs (value = median)")
    ap.add_argument("--cpu-frames", type=int, default=0, help="override: frames of the CPU sample")
    ap.add_argument("--no-spec-stream", action="store_true", help="skip the SURVEY §8(d) config-5 spec stream measurement")
    ap.add_argument("--slab-mb", type=int, default=64)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-packed", action="store_true", help="skip the device-resident K-A<packed> measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
