#!/usr/bin/env python
"""bench.py — motion-scan hot path on B200: MV records/s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU path on the box's host cores

Workload (config.workload): BASELINE.json configs[4] — a decode-free synthetic AVMotionVector stream
of ~10^9 native 40-byte records per GPU (1080p30 CCTV mix cut into 10-minute videos), generated on the
device by the deterministic generator of include/mvgen_core.h. One step = K-A over every frame of the
stream + K-C over every video. Weak scaling: every rank owns a stream of the same size (seed + rank);
videos are independent, so there is no collective on the data path (torch.distributed only provides
the barrier and the max-over-ranks of the timings).

Reported on one JSON line: `value` (device-resident records/s, whole job), `roofline` (K-A's
algorithmic bytes / its CUDA-event duration vs MEASURED_PEAKS.json hbm_gbs), `e2e` (same metric
through the host-facing C ABI with pinned HOST buffers: H2D of every record + D2H of the results
inside the timed region), `cpu_baseline` (the reference built as oracle/_ref, else the oracle port, on this
box's host cores, bounded sample),
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


def bind_to_gpu_numa(index: int):
    """Best effort: restrict this rank to the CPUs NVML reports as local to its GPU (multi-rank runs only)."""
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


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
def cpu_leg(ms, params, spec, n_frames, threads, passes, warmup, sample_data=None):
    """The CPU implementation of the path on this box's host cores, on the first n_frames frames of the
    stream. kind "reference": oracle/_ref/ref_scan — the reference's own motion_scanner.cpp / pipeline.cpp
    (built in the build container against the fake-libav shim) — `threads` MotionScanner instances over
    disjoint time ranges through the public scan_range(), timed by the reference's OWN analyze timer
    around check_frame (motion_scanner.cpp:375-380); the step time is the slowest thread's analyze time
    (decode stand-in excluded). kind "port": the oracle restatement, early-exit semantics, pthreads."""
    import mvs_io
    import oracle_lib as orc
    import ref_runner

    if sample_data is None:
        cnt, off, recs, pts = ms.synth_host(spec, 0, n_frames, n_threads=threads)
    else:
        cnt, off, recs, pts = sample_data
    n_rec = int(off[-1])
    out = {"unit": UNIT, "cores": threads}
    if ref_runner.available():
        import tempfile

        d = "/dev/shm" if os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
        path = os.path.join(d, f"mscan_bench_{os.getpid()}.mvs")
        try:
            mvs_io.write_mvs(path, spec.width, spec.height, int(spec.fps), 1, np.arange(n_frames), cnt, recs)
            dur = n_frames / spec.fps
            chunk = max(1.0, dur / (2 * threads))  # CHUNK_DURATION_SEC: at least 2 chunks per worker thread
            r = ref_runner.run(path, params, threads=threads, passes=passes, warmup=warmup, chunk_sec=chunk)
            r1 = ref_runner.run(path, params, threads=1, passes=1, warmup=0, chunk_sec=chunk) if threads > 1 else r
        finally:
            if os.path.exists(path):
                os.unlink(path)
        hot_s = r["par_analyze_max_us"] * 1e-6
        out.update(
            kind="reference",
            value=n_rec * passes / hot_s,
            seconds_per_step=hot_s / passes,
            value_1thread=n_rec / (r1["analyze_us"] * 1e-6),
            pipeline_value=n_rec * passes / (r["run_wall_us"] * 1e-6),
            motion_frames=int(r["par_motion_frames"]),
            sample=f"first {n_frames} frames / {n_rec} records of the stream; reference sources (oracle/_ref): {threads} "
            f"MotionScanner threads over disjoint ranges, {passes} timed passes, step time = slowest thread's own "
            f"check_frame timer (hot path only); pipeline_value = ProcessingPipeline::run() wall incl. mmap + shim demux",
        )
    else:
        gw, gh, m = orc.geometry(spec.width, spec.height, params.block_size, params.block_shift, params.vertical_mask)
        cfg = orc.make_cfg(params, gw, gh, m)
        fpv = spec.frames_per_video or n_frames

        def step(th):
            flags, _ = orc.scan_frames(cfg, recs, off, early_exit=True, threads=th)  # reference semantics
            for a in range(0, n_frames, fpv):
                b = min(n_frames, a + fpv)
                orc.video_tail(pts[a:b], flags[a:b], (b - a) / spec.fps, params.max_gap_sec, params.padding_sec, params.min_savings_pct)
            return flags

        for _ in range(warmup):
            step(threads)
        t0 = time.perf_counter()
        for _ in range(passes):
            flags = step(threads)
        dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        step(1)
        dt1 = time.perf_counter() - t0
        out.update(
            kind="port",
            value=n_rec * passes / dt,
            seconds_per_step=dt / passes,
            value_1thread=n_rec / dt1,
            motion_frames=int(flags.sum()),
            sample=f"first {n_frames} frames / {n_rec} records of the stream; oracle port (oracle/_ref not built), "
            f"early-exit semantics + tail, {threads} pthreads, {passes} timed passes",
        )
    return out, n_rec


def run_reference(args):
    """--impl reference: the CPU implementation of the path on this box's host cores, all threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import motionscan as ms

    threads = os.cpu_count() or 1
    params = ms.shipped_env_params()
    global _WORKLOAD_DESC
    preset, seed, _fixed, _WORKLOAD_DESC = WORKLOADS[args.workload]
    spec = ms.synth_preset(preset, seed)
    n_frames = args.cpu_frames
    if not n_frames:  # bounded sample: about --cpu-records records from the head of the stream
        probe = np.zeros(64, np.uint32)
        ms.lib().mscan_synth_host_counts(C.byref(spec), 0, 64, probe.ctypes.data, threads)
        n_frames = max(1, int(np.ceil(args.cpu_records / max(probe.mean(), 1.0))))
    cpu, n_rec = cpu_leg(ms, params, spec, n_frames, threads, args.steps, args.warmup)
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
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "int32",
        "data": "synthetic",
        "config": workload_config(spec, n_frames, n_rec, "host"),
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
                  "1080p30 CCTV mix cut into 10-min videos, native 40-B records"),
    "cctv10min": (1, 2, 18000, "cctv10min: synthetic 10 min 1080p30 CCTV-style clip (BASELINE.json configs[1]) as an MV stream"),
    "dense4k": (2, 3, 3600, "dense4k: synthetic 2 min 4K30 clip with a dense 8x8 MV field, 129 600 records per P-frame "
                "(BASELINE.json configs[2]) as an MV stream"),
    "batch64": (3, 100, 64 * 1800, "batch64: 64 synthetic 60 s 1080p30 clips in one batch (BASELINE.json configs[3]) as MV streams"),
}
_WORKLOAD_DESC = WORKLOADS["stream1e9"][3]


def workload_config(spec, n_frames, n_rec, where):
    return {
        "workload": _WORKLOAD_DESC,
        "resident": where,
        "frames": int(n_frames),
        "records": int(n_rec),
        "frames_per_video": int(spec.frames_per_video),
        "params": "config/motion_trim.env: MV_THRESHOLD_SQ=4 VECTORS_NEEDED=4 CLUSTERS_NEEDED=2 VERTICAL_MASK=0.05 MAX_GAP_SEC=5 PADDING_SEC=0.5 MIN_SAVINGS_PCT=5",
        "l2": "inputs larger than L2 (no flush needed)",
    }


# ---------------------------------------------------------------------------------- GPU arm -------
def run_gpu(args):
    import torch

    import motionscan as ms

    from motionscan.dist import Dist, throughput

    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the motion-scan path has no CPU fallback")
    torch.cuda.set_device(local)
    bind_to_gpu_numa(local)  # pinned staging is then allocated on the GPU's own NUMA node
    D = Dist("nccl", torch.device("cuda", local))
    world, rank = D.world, D.rank
    barrier, allmax, allsum = D.barrier, D.allmax, D.allsum

    params = ms.shipped_env_params()
    ctx = ms.Context(local, params, max_log_frames=1 << 20, slab_bytes=args.slab_mb << 20)
    stream = torch.cuda.Stream()
    sh = stream.cuda_stream

    # ---- build the device-resident stream ---------------------------------------------------------
    global _WORKLOAD_DESC
    preset, seed, fixed_frames, _WORKLOAD_DESC = WORKLOADS[args.workload]
    spec = ms.synth_preset(preset, seed + rank)
    if fixed_frames is None:
        probe = 4096
        d_probe = ctx.dev_alloc(4 * probe)
        ctx.synth_counts(spec, 0, probe, d_probe, sh)
        stream.synchronize()
        pc = np.zeros(probe, np.uint32)
        ctx.d2h(pc, d_probe)
        ctx.dev_free(d_probe)
        n_frames = int(np.ceil(args.records / max(pc.mean(), 1.0)))
    else:
        n_frames = fixed_frames
    d_cnt = ctx.dev_alloc(4 * n_frames)
    d_off = ctx.dev_alloc(8 * (n_frames + 1))
    ctx.synth_counts(spec, 0, n_frames, d_cnt, sh)
    ctx.offsets_from_counts(d_cnt, n_frames, d_off, sh)
    stream.synchronize()
    off = np.zeros(n_frames + 1, np.uint64)
    ctx.d2h(off, d_off)
    n_rec = int(off[-1])
    d_recs = ctx.dev_alloc(REC_BYTES * n_rec + 256)
    d_pts = ctx.dev_alloc(8 * n_frames)
    d_flags = ctx.dev_alloc(n_frames)
    d_counts = ctx.dev_alloc(4 * n_frames)
    d_segs = ctx.dev_alloc(16 * n_frames)
    ctx.synth_fill(spec, 0, n_frames, d_off, d_recs, d_pts, sh)
    stream.synchronize()
    fpv = spec.frames_per_video
    voff = np.array(list(range(0, n_frames, fpv)) + [n_frames], dtype=np.uint64)
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
        packed = {
            "kernel": "ka_scan_kernel<packed>",
            "ms_per_launch": pk_ms,
            "records_per_s": allsum(float(n_rec)) / (pk_ms * 1e-3),
            "bytes_per_record": 8,
            "achieved_gbs": (8 * n_rec + FRAME_BYTES * n_frames) / (pk_ms * 1e-3) / 1e9,
            "matches_native": bool(np.array_equal(flags8, flags)),
        }
        for d in (d_r8, d_flags8, d_counts8):
            ctx.dev_free(d)

    # ---- e2e: host-fed through the C ABI ------------------------------------------------------------
    # host-resident sample: the leading frames of the stream, bounded by --e2e-records (~3.6 GB pinned).
    # Three ways for host records to reach the GPU (include/motionscan.h):
    #   native_inplace  native 40-B records DMA'd straight out of the caller's pinned buffer (no host pass)
    #   projected       the library's staging pass keeps bytes 6..13 of each record (pool of host threads)
    #                   and DMAs 8 B/record — what mscan_submit does by default for pageable memory
    #   packed_pinned   the caller hands over records it projected itself (mscan_pack_records on the
    #                   decoder's cache-hot side data); only the DMA + kernels are inside the timed region
    # `e2e.value` is the best of the two modes that start from native host records.
    e2e_frames = int(min(max(np.searchsorted(off, np.uint64(int(args.e2e_records)), side="right") - 1, 1), n_frames))
    if args.e2e_frames:
        e2e_frames = min(args.e2e_frames, n_frames)
    e_rec = int(off[e2e_frames])
    h_recs = ctx.pinned_array(e_rec, ms.MV_DTYPE)
    h_r8 = ctx.pinned_array(e_rec, ms.MV8_DTYPE)
    h_pts = ctx.pinned_array(e2e_frames, np.float64)
    h_cnt = np.diff(off[: e2e_frames + 1]).astype(np.uint32)
    ctx.d2h(h_recs, d_recs)
    ctx.d2h(h_pts, d_pts)
    ms.pack_records(h_recs, h_r8)
    e_voff = list(range(0, e2e_frames, fpv)) + [e2e_frames]
    pack_threads = max(1, len(os.sched_getaffinity(0)) // world)  # ranks share the box's cores
    ctx.set_pack_threads(pack_threads)

    def e2e_step(mode):
        vids = []
        for v in range(len(e_voff) - 1):
            a, b = e_voff[v], e_voff[v + 1]
            ctx.video_open(v, spec.width, spec.height)
            if mode == "packed_pinned":
                ctx.submit_packed_raw(v, b - a, h_pts.ctypes.data + 8 * a, h_cnt.ctypes.data + 4 * a, h_r8.ctypes.data + 8 * int(off[a]))
            else:
                ctx.submit_raw(v, b - a, h_pts.ctypes.data + 8 * a, h_cnt.ctypes.data + 4 * a, h_recs.ctypes.data + REC_BYTES * int(off[a]))
            vids.append(v)
        out = [ctx.collect(v) for v in vids]
        segs = ctx.segments_batch(vids, [(e_voff[v + 1] - e_voff[v]) / spec.fps for v in vids])
        for v in vids:
            ctx.video_close(v)
        return out, segs

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    modes = {}
    for mode in ("native_inplace", "projected", "packed_pinned"):
        ctx.set_staging_mode(ms.STAGING_PACK if mode == "projected" else ms.STAGING_AUTO)
        for _ in range(2):
            e2e_out = e2e_step(mode)
        ctx.sync()
        ctx.reset_stats()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_out = e2e_step(mode)
        ctx.sync()
        dt = allmax(time.perf_counter() - t0)
        est = ctx.stats()
        e_flags = np.concatenate([o[0] for o in e2e_out[0]])
        modes[mode] = {
            "value": allsum(float(e_rec)) * e2e_steps / dt,
            "h2d_bytes_per_step": int(est.h2d_bytes // e2e_steps),
            "d2h_bytes_per_step": int(est.d2h_bytes // e2e_steps),
            "h2d_gbs": est.h2d_bytes / dt / 1e9,
            "launches": int(est.scan_launches + est.segment_launches),
            # the host-fed results must equal the device-resident ones for the same frames
            "matches_device_resident": bool(np.array_equal(e_flags, flags[:e2e_frames])),
        }
        if mode == "projected":
            modes[mode]["host_threads"] = pack_threads
            modes[mode]["project_ms_per_step"] = est.project_ms / e2e_steps
            modes[mode]["records_projected_per_step"] = int(est.records_projected // e2e_steps)
    ctx.set_staging_mode(ms.STAGING_AUTO)
    e2e_mode = max(("native_inplace", "projected"), key=lambda m: modes[m]["value"])
    best = modes[e2e_mode]
    e2e_value, e2e_launches, e2e_ok = best["value"], best["launches"], all(m["matches_device_resident"] for m in modes.values())

    # ---- CPU baseline on rank 0, N=1 only ------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle_lib as orc

        threads = os.cpu_count() or 1
        c_frames = int(min(max(np.searchsorted(off, np.uint64(int(args.cpu_records)), side="right") - 1, 1), e2e_frames))
        if args.cpu_frames:
            c_frames = min(args.cpu_frames, e2e_frames)
        c_off = off[: c_frames + 1]
        c_recs = h_recs[: int(c_off[-1])]
        c_pts = h_pts[:c_frames]
        c_cnt = np.diff(c_off).astype(np.uint32)
        cpu, _ = cpu_leg(ms, params, spec, c_frames, threads, 3, 1, sample_data=(c_cnt, c_off, c_recs, c_pts))
        # parity of the sample while we are here: oracle flags == GPU flags
        gw, gh, m = orc.geometry(spec.width, spec.height, params.block_size, params.block_shift, params.vertical_mask)
        of, _ = orc.scan_frames(orc.make_cfg(params, gw, gh, m), c_recs, c_off, threads=threads)
        cpu["parity_with_gpu"] = bool(np.array_equal(of, flags[:c_frames])) and cpu["motion_frames"] == int(of.sum())

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
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "int32",
            "data": "synthetic",
            "config": workload_config(spec, n_frames, n_rec, "hbm"),
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
                "pcie_frac": best["h2d_gbs"] / pcie_gbs if pcie_gbs > 0 else None,
                "pcie_peak_how": "pinned cudaMemcpyAsync H2D, 1 GiB, best of 9, CUDA events (per GPU)",
                "launches": e2e_launches,
                "matches_device_resident": e2e_ok,
                "modes": modes,
                "how": "per step: mscan_video_open / mscan_submit of native 40-B host records / collect / segments_batch / close; "
                "value = best of native_inplace (pinned records DMA'd in place, 40 B/record over PCIe) and projected (the "
                "library's staging pass keeps the 8 bytes the path reads, 8 B/record over PCIe); "
                "packed_pinned (caller-projected records) is reported in modes only",
            },
            "packed_kernel": packed,
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
    ap.add_argument("--records", type=float, default=1e9, help="records per GPU in the device-resident stream (stream1e9)")
    ap.add_argument("--e2e-records", type=float, default=9e7, help="records of the host-resident e2e sample (~3.6 GB pinned)")
    ap.add_argument("--e2e-frames", type=int, default=0, help="override: frames of the e2e sample")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-records", type=float, default=3e7, help="records of the CPU-baseline / reference-arm sample")
    ap.add_argument("--cpu-frames", type=int, default=0, help="override: frames of the CPU sample")
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
