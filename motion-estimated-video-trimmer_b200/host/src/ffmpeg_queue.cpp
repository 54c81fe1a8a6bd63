// ffmpeg_queue.cpp — job queue + muxing step (contract of reference src/ffmpeg_queue.cpp and
// src/ffmpeg_executor.cpp:24-118; see ffmpeg_queue.hpp).
#include "motion_trim/ffmpeg_queue.hpp"

#include <spawn.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <fstream>

namespace motion_trim {

void FFmpegQueue::push(FFmpegJob job) {
  {
    std::lock_guard<std::mutex> lk(mutex_);
    jobs_.push(std::move(job));
  }
  cv_.notify_one();
}

bool FFmpegQueue::pop(FFmpegJob& job) {
  std::unique_lock<std::mutex> lk(mutex_);
  cv_.wait(lk, [this] { return !jobs_.empty() || done_.load(); });
  if (jobs_.empty()) return false;
  job = std::move(jobs_.front());
  jobs_.pop();
  return true;
}

void FFmpegQueue::finish() {
  done_.store(true);
  cv_.notify_all();
}

bool FFmpegQueue::empty() const {
  std::lock_guard<std::mutex> lk(mutex_);
  return jobs_.empty();
}

std::string build_concat_list(const std::string& input_path, const std::vector<TimeSegment>& segments) {
  const std::string abs = std::filesystem::absolute(input_path).string();
  std::string out;
  char line[64];
  for (const TimeSegment& s : segments) {
    if (s.end <= s.start) continue;
    out += "file '";
    for (char ch : abs) {  // the concat demuxer's quoting: a single quote inside '...' is written '\''
      if (ch == '\'') out += "'\\''";
      else out += ch;
    }
    out += "'\n";
    std::snprintf(line, sizeof line, "inpoint %.2f\n", s.start);
    out += line;
    std::snprintf(line, sizeof line, "outpoint %.2f\n", s.end);
    out += line;
  }
  return out;
}

static std::string find_ffmpeg() {
  if (const char* e = std::getenv("MOTION_TRIM_FFMPEG")) return e;
  const char* dflt = "/usr/local/bin/ffmpeg";
  return access(dflt, X_OK) == 0 ? dflt : "";
}

int execute_ffmpeg_cut(const std::string& input_path, const std::string& output_path,
                       const std::vector<TimeSegment>& segments, const std::vector<int>& cpu_set, int stream_id) {
  if (segments.empty()) {
    std::printf("[WARN] %sNo segments to cut\n", stream_id >= 0 ? ("[Stream " + std::to_string(stream_id) + "] ").c_str() : "");
    return 0;
  }
  const std::string list = build_concat_list(input_path, segments);
  const std::string ffmpeg = find_ffmpeg();
  const std::string list_path = output_path + ".concat.txt";
  {
    std::ofstream f(list_path, std::ios::binary);
    if (!f) return -1;
    f << list;
  }
  if (ffmpeg.empty()) {
    std::printf("[WARN] no ffmpeg binary in this image: wrote %s instead of cutting\n", list_path.c_str());
    return 0;
  }
  // Same command line as the reference (src/ffmpeg_executor.cpp:60-100: taskset -c <cpus> ffmpeg -f concat … -c copy),
  // but spawned with an argv vector instead of through a shell: file names are data (watch mode feeds arbitrary
  // directory entries), so quotes, $(…) or backticks in them must never be interpreted.
  std::vector<std::string> args;
  if (!cpu_set.empty()) {
    std::string cpus;
    for (size_t i = 0; i < cpu_set.size(); ++i) cpus += (i ? "," : "") + std::to_string(cpu_set[i]);
    args = {"taskset", "-c", cpus};
  }
  args.push_back(ffmpeg);
  for (const char* a : {"-y", "-hide_banner", "-loglevel", "error", "-f", "concat", "-safe", "0", "-i"}) args.push_back(a);
  args.push_back(list_path);
  for (const char* a : {"-c", "copy", "-fflags", "+genpts", "-avoid_negative_ts", "make_zero", "-movflags", "+faststart"}) args.push_back(a);
  args.push_back(output_path);
  std::vector<char*> argv;
  for (std::string& a : args) argv.push_back(a.data());
  argv.push_back(nullptr);
  pid_t pid = 0;
  int status = -1;
  if (posix_spawnp(&pid, argv[0], nullptr, nullptr, argv.data(), environ) == 0) {
    int st = 0;
    while (waitpid(pid, &st, 0) < 0 && errno == EINTR) {
    }
    status = WIFEXITED(st) ? WEXITSTATUS(st) : -1;
  }
  std::remove(list_path.c_str());
  if (status != 0) std::printf("[ERROR] FFmpeg failed with status %d\n", status);
  return status;
}

}  // namespace motion_trim
