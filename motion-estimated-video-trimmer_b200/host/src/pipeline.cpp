// pipeline.cpp — host mirror of ProcessingPipeline::run (reference src/pipeline.cpp:89-415).
// Phases kept: map file (:95), probe (:108-125), chunk queue (:141-167), worker threads each with a
// private scanner (:186-235), decision + FFmpegJob (:358-404). Phases moved to the GPU: check_frame
// inside the workers, and merge/segment/savings (:297-356) as one mscan_segments call.
#include "motion_trim/pipeline.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <mutex>
#include <thread>

#include "motion_trim/config.hpp"
#include "motion_trim/gpu_pool.hpp"
#include "motion_trim/motion_scanner.hpp"
#include "motion_trim/system.hpp"

namespace motion_trim {

namespace {
std::mutex g_log_mu;
void logf(int stream_id, const char* level, const std::string& msg) {
  std::lock_guard<std::mutex> lk(g_log_mu);
  if (stream_id >= 0) std::printf("%s[Stream %d] %s\n", level, stream_id, msg.c_str());
  else std::printf("%s%s\n", level, msg.c_str());
  std::fflush(stdout);
}
}  // namespace

ProcessingPipeline::ProcessingPipeline(std::string in, std::string out, int stream_id, int num_threads,
                                       std::vector<int> cpu_set)
    : input_path_(std::move(in)), output_path_(std::move(out)), stream_id_(stream_id), num_threads_(num_threads),
      cpu_set_(std::move(cpu_set)) {}

void ProcessingPipeline::set_gpu(GpuPool* pool, int gpu_index) { set_gpus(pool, {gpu_index}); }

void ProcessingPipeline::set_gpus(GpuPool* pool, std::vector<int> gpu_indices) {
  pool_ = pool;
  gpus_ = std::move(gpu_indices);
  gpu_index_ = gpus_.empty() ? -1 : gpus_.front();
}

int ProcessingPipeline::run() {
  bool gpus_ok = pool_ && !gpus_.empty();
  for (int g : gpus_) gpus_ok = gpus_ok && g >= 0 && g < pool_->size();
  if (!gpus_ok) {
    logf(stream_id_, "[ERROR] ", "no GPU context (the motion scan has no CPU fallback)");
    return 1;
  }
  mscan_ctx* gpu = pool_->ctx(gpu_index_);
  const size_t n_gpus = gpus_.size();
  using clk = std::chrono::steady_clock;
  auto lap = [t = clk::now()](double& into) mutable {
    const auto n = clk::now();
    into += std::chrono::duration<double>(n - t).count();
    t = n;
  };
  if (!MemoryLoader::load_file(input_path_, file_buffer_)) {
    logf(stream_id_, "[ERROR] ", "Failed to map file: " + input_path_);
    return 1;
  }
  lap(phases_.map);
  const uint32_t video_id = pool_->next_video_id();  // the same id on every GPU that scans this video
  double fps = 0;
  int width = 0, height = 0;
  bool decode_fed = false;
  {
    MotionScanner probe(file_buffer_, gpu, video_id);
    if (!probe.initialize()) {
#ifdef MT_WITH_FFMPEG
      logf(stream_id_, "[ERROR] ", "Failed to initialize probe (neither an MVS1 stream nor media FFmpeg can open)");
#else
      logf(stream_id_, "[ERROR] ", "Failed to initialize probe (not an MVS1 motion-vector stream; media files need a -DMT_WITH_FFMPEG build)");
#endif
      return 1;
    }
    duration_ = probe.get_duration();
    fps = probe.get_fps();
    width = probe.width();
    height = probe.height();
    decode_fed = probe.uses_ffmpeg();
  }
  if (decode_fed) file_buffer_.will_need();
  if (!(duration_ > 0)) {  // the reference divides by zero here (pipeline.cpp:141-143,259); refuse instead
    logf(stream_id_, "[ERROR] ", "stream has no duration");
    return 1;
  }
  lap(phases_.probe);
  // Pin the mapped stream so mscan_submit DMAs records straight out of the page cache; if the platform
  // refuses (or MOTION_TRIM_NO_PIN is set) submits fall back to the library's pinned staging copy.
  bool registered = false;
  // (decode-fed runs stage projected records instead: nothing is DMA'd out of the media file)
  if (!decode_fed && !std::getenv("MOTION_TRIM_NO_PIN")) {
    registered = mscan_host_register(gpu, const_cast<uint8_t*>(file_buffer_.data()), file_buffer_.size(), 1) == MSCAN_OK;
    static std::atomic<bool> told{false};
    if (!registered && !told.exchange(true))
      logf(stream_id_, "[INFO] ", std::string("input mapping not pinned (") + mscan_last_error(gpu) + "): records are projected into the library's pinned ring instead");
  }
  lap(phases_.pin);
  std::vector<mscan_ctx*> readers;  // every context that may DMA out of the mapping
  for (int g : gpus_) readers.push_back(pool_->ctx(g));
  struct Unpin {
    std::vector<mscan_ctx*> gs;
    const uint8_t* p;
    bool on;
    double* seconds;
    ~Unpin() {
      if (!on) return;
      const auto t0 = clk::now();
      for (mscan_ctx* g : gs) mscan_host_fence(g);  // no DMA may still be reading the mapping
      mscan_host_unregister(gs.front(), const_cast<uint8_t*>(p));
      *seconds += std::chrono::duration<double>(clk::now() - t0).count();
    }
  } unpin{readers, file_buffer_.data(), registered, &phases_.unpin};
  for (size_t k = 0; k < n_gpus; ++k) {
    mscan_ctx* g = pool_->ctx(gpus_[k]);
    if (mscan_video_open(g, video_id, width, height) != MSCAN_OK) {
      logf(stream_id_, "[ERROR] ", std::string("mscan_video_open: ") + mscan_last_error(g));
      for (size_t j = 0; j < k; ++j) mscan_video_close(pool_->ctx(gpus_[j]), video_id);
      return 1;
    }
  }
  auto close_all = [&] {
    for (int g : gpus_) mscan_video_close(pool_->ctx(g), video_id);
  };
  char buf[160];
  if (n_gpus > 1) {
    std::snprintf(buf, sizeof buf, "Splitting the video over %zu GPUs (stitched on GPU %d)", n_gpus, gpu_index_);
    logf(stream_id_, "[INFO] ", buf);
  }
  std::snprintf(buf, sizeof buf, "Duration: %.2fs (%.0f frames @ %.1ffps) on GPU %d%s", duration_, duration_ * fps, fps, gpu_index_,
                decode_fed ? ", FFmpeg export_mvs front-end" : registered ? ", input pinned for in-place DMA" : "");
  logf(stream_id_, "[INFO] ", buf);

  // ---- chunk queue + workers (the workers are the decode front-end; here they walk the MVS index)
  const double chunk = Config::chunk_duration_sec();
  std::vector<ScanTask> tasks;
  int id = 0;
  for (double t = 0; t < duration_; t += chunk) tasks.push_back(ScanTask{t, std::min(t + chunk, duration_), id++});
  int n_threads = num_threads_ > 0 ? num_threads_ : std::max(2u, std::thread::hardware_concurrency());
  n_threads = std::max(1, std::min<int>(n_threads, (int)tasks.size()));
  n_threads = std::max<int>(n_threads, (int)std::min(n_gpus, tasks.size()));  // every GPU gets a worker
  std::atomic<size_t> next{0};
  std::atomic<long> frames{0};
  std::atomic<bool> failed{false};
  std::vector<std::thread> workers;
  for (int w = 0; w < n_threads; ++w)
    workers.emplace_back([&, w] {
      if (!cpu_set_.empty()) pin_thread_to_cpus(cpu_set_);  // src/pipeline.cpp:191-193
      MotionScanner scanner(file_buffer_, pool_->ctx(gpus_[(size_t)w % n_gpus]), video_id);
      if (!scanner.initialize()) {
        failed = true;
        return;
      }
      long seek_us = 0, decode_us = 0;
      for (size_t k = next.fetch_add(1); k < tasks.size(); k = next.fetch_add(1)) {
        const long n = scanner.scan_range_async(tasks[k].start, tasks[k].end, seek_us, decode_us);
        if (n < 0) {
          failed = true;
          return;
        }
        frames += n;
      }
    });
  for (auto& t : workers) t.join();
  lap(phases_.scan);
  frames_scanned_ = (uint64_t)frames.load();
  if (failed) {
    logf(stream_id_, "[ERROR] ", std::string("scan failed: ") + mscan_last_error(gpu));
    close_all();
    return 1;
  }
  // ---- cross-GPU stitch: the other GPUs' per-frame results join the first GPU's log (NVLink peer copy)
  for (size_t k = 1; k < n_gpus; ++k)
    if (mscan_video_append_from(gpu, video_id, pool_->ctx(gpus_[k]), video_id) != MSCAN_OK) {
      logf(stream_id_, "[ERROR] ", std::string("mscan_video_append_from: ") + mscan_last_error(gpu));
      close_all();
      return 1;
    }

  // ---- merge + segments + savings + decision on the GPU (pipeline.cpp:297-358)
  segments_.assign(64, TimeSegment{0, 0});
  uint32_t n_seg = 0;
  mscan_video_result res{};
  int rc = mscan_segments(gpu, video_id, duration_, reinterpret_cast<mscan_segment*>(segments_.data()),
                          (uint32_t)segments_.size(), &n_seg, &res);
  if (rc == MSCAN_ERR_CAPACITY) {
    segments_.assign(n_seg, TimeSegment{0, 0});
    rc = mscan_segments(gpu, video_id, duration_, reinterpret_cast<mscan_segment*>(segments_.data()),
                        (uint32_t)segments_.size(), &n_seg, &res);
  }
  close_all();
  lap(phases_.segments);
  if (rc != MSCAN_OK) {
    logf(stream_id_, "[ERROR] ", std::string("mscan_segments: ") + mscan_last_error(gpu));
    return 1;
  }
  segments_.resize(n_seg);
  decision_ = res.decision;
  time_removed_ = res.time_removed;
  saved_pct_ = res.saved_pct;
  std::snprintf(buf, sizeof buf, "Scanned %llu frames, found %u motion frames", (unsigned long long)frames_scanned_, res.n_motion_frames);
  logf(stream_id_, "[INFO] ", buf);

  if (decision_ == MSCAN_NO_MOTION) {  // :308-319 — warn, return 0, no job, no output file
    logf(stream_id_, "[WARN] ", "No motion found.");
    return 0;
  }
  if (decision_ == MSCAN_FULL_COPY) {
    std::snprintf(buf, sizeof buf, "Savings too low (%d%%). Min required: %d%%. Copying full stream.", (int)saved_pct_,
                  (int)Config::min_savings_pct());
    logf(stream_id_, "[WARN] ", buf);
  }
  if (ffmpeg_queue_) {
    FFmpegJob job;
    job.stream_id = stream_id_;
    job.input_path = std::filesystem::absolute(input_path_).string();
    job.output_path = output_path_;
    job.segments = segments_;
    job.cpu_set = cpu_set_;
    ffmpeg_queue_->push(std::move(job));
    logf(stream_id_, "[INFO] ", decision_ == MSCAN_CUT ? "Pushed FFmpeg job to queue" : "Pushed full-copy job to queue");
    return 0;
  }
  return execute_ffmpeg_cut(input_path_, output_path_, segments_, cpu_set_, stream_id_) == 0 ? 0 : 1;
}

}  // namespace motion_trim
