// ref_harness.cpp — drives the REFERENCE's own classes (compiled unmodified from /root/reference
// against the fake-libav shim) over an MVS1 stream file. TEST INFRASTRUCTURE ONLY (oracle/_ref).
//
//   ref_scan <in.mvs> <out.bin> <threads> [passes] [warmup]
//
// Knobs come from the environment exactly as in the reference (include/motion_trim/config.hpp), so
// every parameter set is a separate process. Two legs, both pure reference code:
//   1. MotionScanner::initialize + scan_range(0, duration)   → timestamps with motion, analyze_us
//      (src/motion_scanner.cpp:62-202, 297-391; check_frame :217-295 is private and reached this way)
//   2. ProcessingPipeline::run with an FFmpegQueue attached  → the FFmpegJob it would hand to ffmpeg
//      (src/pipeline.cpp:89-404: chunking, worker threads, merge, segments, savings, decision)
// Output (little-endian): RefResult header, then n_ts doubles, then n_segs {start,end} doubles.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "motion_trim/config.hpp"
#include "motion_trim/ffmpeg_queue.hpp"
#include "motion_trim/memory_io.hpp"
#include "motion_trim/motion_scanner.hpp"
#include "motion_trim/pipeline.hpp"

using namespace motion_trim;

struct RefResult {
  int32_t scan_ok;       // initialize() succeeded
  int32_t run_rc;        // ProcessingPipeline::run() return code
  int32_t decision;      // 0 no job, 1 cut, 2 full copy
  uint32_t n_ts;
  uint32_t n_segs;
  uint32_t passes;
  double duration;
  double time_removed;
  double saved_pct;
  double fps;
  int64_t analyze_us;    // the reference's own timer around check_frame (motion_scanner.cpp:375-380)
  int64_t decode_us;
  int64_t scan_wall_us;  // wall time of leg 1
  int64_t run_wall_us;   // wall time of leg 2, summed over the timed passes (warm-up passes excluded)
  // leg 3: `threads` scanners over disjoint time ranges, like the pipeline's workers (pipeline.cpp:186-235),
  // timed passes only
  int64_t par_analyze_sum_us;  // Σ over threads and passes of the reference's analyze timer
  int64_t par_analyze_max_us;  // Σ over passes of the slowest thread's analyze time (the hot path's wall time)
  int64_t par_wall_us;         // Σ over passes of the leg's wall time (includes the shim's demux/"decode" copies)
  int64_t par_motion_frames;   // timestamps found by the last pass (must equal n_ts)
};

int main(int argc, char** argv) {
  if (argc < 4) {
    std::fprintf(stderr, "usage: ref_scan <in.mvs> <out.bin> <threads> [passes] [warmup]\n");
    return 2;
  }
  const char* in = argv[1];
  const char* out = argv[2];
  const int threads = std::atoi(argv[3]);
  const int passes = argc > 4 ? std::max(1, std::atoi(argv[4])) : 1;
  const int warmup = argc > 5 ? std::max(0, std::atoi(argv[5])) : 0;
  RefResult r{};
  std::vector<double> ts;
  std::vector<TimeSegment> segs;

  {  // leg 1: one scanner over the whole file
    MappedFile file;
    if (!MemoryLoader::load_file(in, file)) return 3;
    MotionScanner scanner(file);
    r.scan_ok = scanner.initialize() ? 1 : 0;
    if (r.scan_ok) {
      r.duration = scanner.get_duration();
      r.fps = scanner.get_fps();
      long seek_us = 0, decode_us = 0, analyze_us = 0;
      auto t0 = std::chrono::steady_clock::now();
      ts = scanner.scan_range(0.0, r.duration, seek_us, decode_us, analyze_us);
      auto t1 = std::chrono::steady_clock::now();
      r.analyze_us = analyze_us;
      r.decode_us = decode_us;
      r.scan_wall_us = std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count();
    }
  }

  for (int p = 0; p < warmup + passes; ++p) {  // leg 2: the whole per-video pipeline
    FFmpegQueue queue;
    ProcessingPipeline pipe(in, "/nonexistent/ref_scan_output.mp4", 0, threads, {});
    pipe.set_ffmpeg_queue(&queue);
    auto t0 = std::chrono::steady_clock::now();
    r.run_rc = pipe.run();
    auto t1 = std::chrono::steady_clock::now();
    if (p >= warmup) r.run_wall_us += std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count();
    queue.finish();
    FFmpegJob job;
    segs.clear();
    if (queue.pop(job)) {
      segs = job.segments;
      r.decision = (pipe.get_saved_pct() > Config::min_savings_pct()) ? 1 : 2;
    } else {
      r.decision = 0;
    }
    r.time_removed = pipe.get_time_removed();
    r.saved_pct = pipe.get_saved_pct();
    r.passes = (uint32_t)(p >= warmup ? p - warmup + 1 : 0);
  }

  if (r.scan_ok) {  // leg 3
    MappedFile file;
    if (!MemoryLoader::load_file(in, file)) return 3;
    const int T = threads < 1 ? 1 : threads;
    for (int p = 0; p < warmup + passes; ++p) {
      std::vector<long> an(T, 0), found(T, 0);
      std::vector<std::thread> th;
      auto t0 = std::chrono::steady_clock::now();
      for (int t = 0; t < T; ++t)
        th.emplace_back([&, t] {
          MotionScanner sc(file);
          if (!sc.initialize()) return;
          long seek_us = 0, decode_us = 0, analyze_us = 0;
          const double a = r.duration * t / T, b = (t + 1 == T) ? r.duration : r.duration * (t + 1) / T;
          found[t] = (long)sc.scan_range(a, b, seek_us, decode_us, analyze_us).size();
          an[t] = analyze_us;
        });
      for (auto& x : th) x.join();
      auto t1 = std::chrono::steady_clock::now();
      if (p >= warmup) {
        long mx = 0, tot = 0;
        r.par_motion_frames = 0;
        for (int t = 0; t < T; ++t) {
          mx = std::max(mx, an[t]);
          tot += an[t];
          r.par_motion_frames += found[t];
        }
        r.par_analyze_sum_us += tot;
        r.par_analyze_max_us += mx;
        r.par_wall_us += std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count();
      }
    }
  }

  r.n_ts = (uint32_t)ts.size();
  r.n_segs = (uint32_t)segs.size();
  std::FILE* f = std::fopen(out, "wb");
  if (!f) return 4;
  std::fwrite(&r, sizeof r, 1, f);
  if (!ts.empty()) std::fwrite(ts.data(), sizeof(double), ts.size(), f);
  for (const auto& s : segs) {
    const double pair[2] = {s.start, s.end};
    std::fwrite(pair, sizeof(double), 2, f);
  }
  std::fclose(f);
  return 0;
}
