// batch_processor.cpp — host mirror of BatchProcessor::process (reference src/batch_processor.cpp:48-213):
// skip inputs whose output exists (:60-72), one FFmpeg worker thread draining the job queue
// (:138-150), stream threads pulling files from a shared queue (:215-235,307-382), failure count as
// the return value (:204-212). New: the stream threads are spread over the GPUs of the box.
#include "motion_trim/batch_processor.hpp"

#include <chrono>
#include <cstdio>
#include <filesystem>
#include <thread>

#include "motion_trim/config.hpp"
#include "motion_trim/pipeline.hpp"

namespace fs = std::filesystem;

namespace motion_trim {

BatchProcessor::BatchProcessor(int parallel_streams) : streams_per_gpu_(parallel_streams > 0 ? parallel_streams : 2) {}

bool BatchProcessor::next_file(std::string& out) {
  std::lock_guard<std::mutex> lk(queue_mu_);
  if (work_.empty()) return false;
  out = std::move(work_.front());
  work_.pop();
  return true;
}

void BatchProcessor::stream_worker(int stream_id, int gpu, const std::string& output_dir) {
  int threads = Config::threads_per_stream();
  if (threads <= 0) threads = std::max(1u, std::thread::hardware_concurrency() / (unsigned)(streams_per_gpu_ * pool_.size()));
  std::string file;
  while (next_file(file)) {
    const std::string out = (fs::path(output_dir) / fs::path(file).filename()).string();
    const auto t0 = std::chrono::steady_clock::now();
    ProcessingPipeline p(file, out, stream_id, threads);
    p.set_gpu(&pool_, gpu);
    p.set_ffmpeg_queue(&ffmpeg_queue_);
    StreamResult r;
    r.file = file;
    r.gpu = gpu;
    r.rc = p.run();
    r.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    r.decision = p.get_decision();
    r.duration = p.get_duration();
    r.time_removed = p.get_time_removed();
    r.saved_pct = p.get_saved_pct();
    r.frames = p.frames_scanned();
    r.n_segments = p.get_segments().size();
    if (r.rc != 0) failures_++;
    std::lock_guard<std::mutex> lk(results_mu_);
    results_.push_back(std::move(r));
  }
}

int BatchProcessor::process(const std::vector<std::string>& input_files, const std::string& output_dir,
                            const std::string&) {
  if (input_files.empty()) {
    std::printf("[WARN] No input files to process\n");
    return 0;
  }
  for (const std::string& f : input_files) {
    const std::string out = (fs::path(output_dir) / fs::path(f).filename()).string();
    if (fs::exists(out)) {
      std::printf("[INFO] Skipping existing output: %s\n", out.c_str());
      continue;
    }
    work_.push(f);
  }
  if (!pool_.open(Config::gpus())) {
    std::printf("[ERROR] %s\n", pool_.error().c_str());
    return (int)work_.size();
  }
  const int n_gpus = pool_.size();
  std::printf("[INFO] %zu files, %d GPU(s), %d stream(s) per GPU\n", work_.size(), n_gpus, streams_per_gpu_);
  const auto t0 = std::chrono::steady_clock::now();

  std::atomic<int> mux_failures{0};
  std::thread ffmpeg_worker([&] {
    FFmpegJob job;
    while (ffmpeg_queue_.pop(job))
      if (execute_ffmpeg_cut(job.input_path, job.output_path, job.segments, job.cpu_set, job.stream_id) != 0) mux_failures++;
  });
  std::vector<std::thread> streams;
  for (int g = 0; g < n_gpus; ++g)
    for (int s = 0; s < streams_per_gpu_; ++s)
      streams.emplace_back(&BatchProcessor::stream_worker, this, g * streams_per_gpu_ + s, g, output_dir);
  for (auto& t : streams) t.join();
  ffmpeg_queue_.finish();
  ffmpeg_worker.join();

  const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  double sum = 0;
  uint64_t frames = 0;
  for (const auto& r : results_) {
    sum += r.seconds;
    frames += r.frames;
  }
  std::printf("========== BATCH SUMMARY ==========\n");
  std::printf("files %zu  failed %d  frames %llu  wall %.3fs  speedup %.2fx (sum of file times / wall)\n", results_.size(),
              failures_.load() + mux_failures.load(), (unsigned long long)frames, wall, wall > 0 ? sum / wall : 0.0);
  return failures_.load() + mux_failures.load();
}

}  // namespace motion_trim
