// batch_processor.cpp — host mirror of BatchProcessor::process (reference src/batch_processor.cpp:48-213):
// skip inputs whose output exists (:60-72), one FFmpeg worker thread draining the job queue
// (:138-150), stream threads pulling files from a shared queue (:215-235,307-382), failure count as
// the return value (:204-212). New: the stream threads are spread over the GPUs of the box.
#include "motion_trim/batch_processor.hpp"

#include <cctype>
#include <chrono>
#include <cstdio>
#include <filesystem>
#include <system_error>
#include <thread>

#include "motion_trim/config.hpp"
#include "motion_trim/pipeline.hpp"
#include "motion_trim/system.hpp"

namespace fs = std::filesystem;

namespace motion_trim {

BatchProcessor::BatchProcessor(int parallel_streams) : streams_per_gpu_(parallel_streams > 0 ? parallel_streams : 2) {}

bool BatchProcessor::next_file(std::string& out) {
  std::unique_lock<std::mutex> lk(queue_mu_);
  if (watching_) queue_cv_.wait(lk, [this] { return !work_.empty() || stop_watch_; });  // :218-227
  if (work_.empty()) return false;
  out = std::move(work_.front());
  work_.pop();
  in_progress_++;
  return true;
}

void BatchProcessor::reaper_loop() {
  for (;;) {
    MappedFile f;
    {
      std::unique_lock<std::mutex> lk(reap_mu_);
      reap_cv_.wait(lk, [this] { return !reap_.empty() || reap_done_; });
      if (reap_.empty()) return;
      f = std::move(reap_.front());
      reap_.pop();
    }
  }  // ~MappedFile unmaps here, outside the lock
}

void BatchProcessor::stop_watch() {
  stop_watch_ = true;
  queue_cv_.notify_all();
}

void BatchProcessor::monitor_directory(const std::string& input_dir, const std::string& output_dir) {
  using namespace std::chrono;
  const double idle_exit = Config::watch_idle_exit_sec();
  auto last_activity = steady_clock::now();
  for (int poll = 0; !stop_watch_; ++poll) {
    if (poll % 15 == 0) std::printf("[INFO] [Watch] Monitoring directory: %s (Waiting for new files...)\n", input_dir.c_str());
    std::error_code ec;
    for (fs::directory_iterator it(input_dir, ec), end; !ec && it != end; it.increment(ec)) {
      if (!it->is_regular_file(ec)) continue;
      const std::string path = it->path().string();
      std::string ext = it->path().extension().string();
      for (char& ch : ext) ch = (char)std::tolower((unsigned char)ch);
      if ((accept_ && !accept_(ext)) || seen_.count(path)) continue;
      if (fs::exists(fs::path(output_dir) / it->path().filename())) {  // persistence across restarts (:262-269)
        std::printf("[INFO] [Watch] Skipping file (already processed): %s\n", it->path().filename().c_str());
        seen_.insert(path);
        continue;
      }
      const auto size1 = fs::file_size(path, ec);  // still being written? (:273-279)
      std::this_thread::sleep_for(milliseconds(500));
      const auto size2 = fs::file_size(path, ec);
      if (ec || size1 != size2) continue;
      std::printf("[INFO] [Watch] New file detected: %s\n", it->path().filename().c_str());
      {
        std::lock_guard<std::mutex> lk(queue_mu_);
        work_.push(path);
        seen_.insert(path);
      }
      queue_cv_.notify_one();
      last_activity = steady_clock::now();
    }
    bool busy;
    {
      std::lock_guard<std::mutex> lk(queue_mu_);
      busy = !work_.empty() || in_progress_.load() > 0;
    }
    if (busy) last_activity = steady_clock::now();
    else if (idle_exit > 0 && duration<double>(steady_clock::now() - last_activity).count() >= idle_exit) stop_watch();
    for (int k = 0; k < 20 && !stop_watch_; ++k) std::this_thread::sleep_for(milliseconds(100));  // 2 s poll (:297)
  }
  queue_cv_.notify_all();
}

void BatchProcessor::stream_worker(int stream_id, int gpu, std::vector<int> cpu_set, int threads, const std::string& output_dir) {
  if (pin_thread_to_cpus(cpu_set))  // like src/batch_processor.cpp:307-320: the stream thread lives on its own CPUs
    std::printf("[INFO] [Stream %d] GPU %d, pinned to CPUs [%s]\n", stream_id, gpu, cpu_list_string(cpu_set).c_str());
  std::string file;
  while (next_file(file)) {
    const std::string out = (fs::path(output_dir) / fs::path(file).filename()).string();
    const auto t0 = std::chrono::steady_clock::now();
    ProcessingPipeline p(file, out, stream_id, threads, cpu_set);
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
    r.t_map = p.phases().map;
    r.t_probe = p.phases().probe;
    r.t_pin = p.phases().pin;
    r.t_scan = p.phases().scan;
    r.t_segments = p.phases().segments;
    {
      std::lock_guard<std::mutex> lk(reap_mu_);
      reap_.push(p.release_input());
    }
    reap_cv_.notify_one();
    if (r.rc != 0) failures_++;
    {
      std::lock_guard<std::mutex> lk(results_mu_);
      results_.push_back(std::move(r));
    }
    in_progress_--;
  }
}

int BatchProcessor::process(const std::vector<std::string>& input_files, const std::string& output_dir,
                            const std::string& input_dir_arg) {
  watching_ = Config::watch_mode();
  if (input_files.empty() && !watching_) {
    std::printf("[WARN] No input files to process\n");
    return 0;
  }
  for (const std::string& f : input_files) {
    const std::string out = (fs::path(output_dir) / fs::path(f).filename()).string();
    if (fs::exists(out)) {
      std::printf("[INFO] Skipping existing output: %s\n", out.c_str());
      seen_.insert(f);
      continue;
    }
    work_.push(f);
    seen_.insert(f);
  }
  if (!pool_.open(Config::gpus())) {
    std::printf("[ERROR] %s\n", pool_.error().c_str());
    return (int)work_.size();
  }
  const int n_gpus = pool_.size();
  // CPU plan (reference src/batch_processor.cpp:78-110): the CPUs the process may use are dealt to the streams in
  // disjoint sets of THREADS_PER_STREAM (auto: CPUs / streams); a stream's chunk workers, its projection work and its
  // FFmpeg job stay on that set. New: a GPU's streams take CPUs local to that GPU first (sysfs local_cpulist), so the
  // pinned staging they fill is on the GPU's NUMA node; with one set of CPUs for all GPUs this is the reference's plan.
  const std::vector<int> cpus = get_available_cpus();
  const int n_streams = n_gpus * streams_per_gpu_;
  int threads = Config::threads_per_stream();
  if (threads <= 0) threads = std::max(1, (int)cpus.size() / n_streams);
  std::vector<std::vector<int>> stream_cpus((size_t)n_streams);
  {
    std::set<int> free_cpus(cpus.begin(), cpus.end());
    for (int g = 0; g < n_gpus; ++g) {
      char bus[32] = {0};
      std::vector<int> local;
      if (mscan_device_pci_bus_id(g, bus, (int)sizeof bus) == MSCAN_OK) local = pci_local_cpus(bus);
      for (int s = 0; s < streams_per_gpu_; ++s) {
        std::vector<int>& mine = stream_cpus[(size_t)(g * streams_per_gpu_ + s)];
        for (int c : local)
          if ((int)mine.size() < threads && free_cpus.erase(c)) mine.push_back(c);
        while ((int)mine.size() < threads && !free_cpus.empty()) {  // not enough local ones left: any free CPU
          mine.push_back(*free_cpus.begin());
          free_cpus.erase(free_cpus.begin());
        }
      }
    }
  }
  // When the streams' chunk workers already occupy every CPU, each worker projects its own chunk on its own CPU; the
  // library's shared pool (all CPUs per large submit) is for boxes with fewer workers than CPUs.
  if (n_streams * threads >= (int)cpus.size())
    for (int g = 0; g < n_gpus; ++g) mscan_set_pack_threads(pool_.ctx(g), 1);
  std::printf("[INFO] %zu files, %d GPU(s), %d stream(s) per GPU, %d thread(s)/CPU(s) per stream, %zu CPUs available\n", work_.size(), n_gpus,
              streams_per_gpu_, threads, cpus.size());
  const auto t0 = std::chrono::steady_clock::now();

  std::atomic<int> mux_failures{0};
  std::thread ffmpeg_worker([&] {
    FFmpegJob job;
    while (ffmpeg_queue_.pop(job))
      if (execute_ffmpeg_cut(job.input_path, job.output_path, job.segments, job.cpu_set, job.stream_id) != 0) mux_failures++;
  });
  std::thread reaper(&BatchProcessor::reaper_loop, this);
  std::vector<std::thread> streams;
  for (int g = 0; g < n_gpus; ++g)
    for (int s = 0; s < streams_per_gpu_; ++s)
      streams.emplace_back(&BatchProcessor::stream_worker, this, g * streams_per_gpu_ + s, g, stream_cpus[(size_t)(g * streams_per_gpu_ + s)],
                           threads, output_dir);
  if (watching_) {  // :161-183 — the monitor feeds the queue until stopped
    std::string input_dir = input_dir_arg;
    if (input_dir.empty() && !input_files.empty()) input_dir = fs::path(input_files[0]).parent_path().string();
    if (input_dir.empty()) input_dir = ".";
    std::printf("[INFO] Starting Watch Mode on directory: %s\n", input_dir.c_str());
    std::thread monitor(&BatchProcessor::monitor_directory, this, input_dir, output_dir);
    monitor.join();
  }
  for (auto& t : streams) t.join();
  const double scan_wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  {
    std::lock_guard<std::mutex> lk(reap_mu_);
    reap_done_ = true;
  }
  reap_cv_.notify_all();
  reaper.join();
  ffmpeg_queue_.finish();
  ffmpeg_worker.join();

  const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  double sum = 0, ph[5] = {0, 0, 0, 0, 0};
  uint64_t frames = 0;
  for (const auto& r : results_) {
    sum += r.seconds;
    frames += r.frames;
    ph[0] += r.t_map;
    ph[1] += r.t_probe;
    ph[2] += r.t_pin;
    ph[3] += r.t_scan;
    ph[4] += r.t_segments;
  }
  std::printf("========== BATCH SUMMARY ==========\n");
  std::printf("files %zu  failed %d  frames %llu  wall %.3fs  (scan %.3fs)  speedup %.2fx (sum of file times / wall)\n", results_.size(),
              failures_.load() + mux_failures.load(), (unsigned long long)frames, wall, scan_wall, wall > 0 ? sum / wall : 0.0);
  std::printf("phases (sum over files, s): map %.3f  probe %.3f  pin %.3f  submit %.3f  segments+close %.3f  other %.3f\n", ph[0], ph[1],
              ph[2], ph[3], ph[4], sum - ph[0] - ph[1] - ph[2] - ph[3] - ph[4]);
  return failures_.load() + mux_failures.load();
}

}  // namespace motion_trim
