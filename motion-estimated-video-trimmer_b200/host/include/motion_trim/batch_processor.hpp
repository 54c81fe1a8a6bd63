// batch_processor.hpp — host mirror of BatchProcessor (reference include/motion_trim/batch_processor.hpp:77-147):
// BatchProcessor(parallel_streams), process(files, output_dir, input_dir) → number of failures.
// The reference deals files from one queue to PARALLEL_STREAMS CPU stream threads
// (src/batch_processor.cpp:152-157,215-235,307-382); here the same queue feeds PARALLEL_STREAMS stream
// threads PER GPU, each bound to its GPU's context — videos are independent, so GPUs never exchange data.
#pragma once

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <queue>
#include <set>
#include <string>
#include <vector>

#include "ffmpeg_queue.hpp"
#include "gpu_pool.hpp"
#include "memory_io.hpp"

namespace motion_trim {

struct StreamResult {
  std::string file;
  int gpu = 0;
  int rc = 0;
  int decision = 0;
  double seconds = 0, duration = 0, time_removed = 0, saved_pct = 0;
  uint64_t frames = 0, records = 0;
  size_t n_segments = 0;
  double t_map = 0, t_probe = 0, t_pin = 0, t_scan = 0, t_segments = 0, t_unpin = 0;
};

class BatchProcessor {
 public:
  explicit BatchProcessor(int parallel_streams);
  int process(const std::vector<std::string>& input_files, const std::string& output_dir,
              const std::string& input_dir = "");
  const std::vector<StreamResult>& results() const { return results_; }
  // Which directory entries are inputs (main.cpp passes its extension filter; used by watch mode).
  void set_input_filter(std::function<bool(const std::string& ext)> f) { accept_ = std::move(f); }
  // Ends watch mode (the reference has no way to: its stop flag is never set, batch_processor.cpp:180-183,240).
  void stop_watch();

 private:
  bool next_file(std::string& out);
  // WATCH_MODE=1 (reference src/batch_processor.cpp:237-305): poll the input directory every 2 s, enqueue
  // files that are new, have no output yet and whose size was stable for 500 ms.
  void monitor_directory(const std::string& input_dir, const std::string& output_dir);
  void stream_worker(int stream_id, int gpu, std::vector<int> cpu_set, int threads, const std::string& output_dir);

  int streams_per_gpu_;
  GpuPool pool_;
  FFmpegQueue ffmpeg_queue_;
  std::mutex queue_mu_, results_mu_;
  std::condition_variable queue_cv_;
  std::queue<std::string> work_;
  std::set<std::string> seen_;  // watch mode: paths already enqueued or skipped
  std::function<bool(const std::string&)> accept_;
  std::atomic<bool> watching_{false}, stop_watch_{false};
  std::atomic<int> in_progress_{0};
  // finished inputs are unmapped off the stream threads
  void reaper_loop();
  std::mutex reap_mu_;
  std::condition_variable reap_cv_;
  std::queue<MappedFile> reap_;
  bool reap_done_ = false;
  std::vector<StreamResult> results_;
  std::atomic<int> failures_{0};
};

}  // namespace motion_trim
