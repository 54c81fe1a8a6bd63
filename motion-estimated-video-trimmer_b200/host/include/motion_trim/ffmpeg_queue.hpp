// ffmpeg_queue.hpp — producer/consumer hand-off of FFmpegJob to the single muxing worker, same
// contract as the reference (include/motion_trim/ffmpeg_queue.hpp:40-90): push / blocking pop /
// finish. execute_ffmpeg_cut keeps the reference's concat-list format (src/ffmpeg_executor.cpp:38-50).
#pragma once

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <queue>
#include <string>
#include <vector>

#include "types.hpp"

namespace motion_trim {

class FFmpegQueue {
 public:
  void push(FFmpegJob job);
  bool pop(FFmpegJob& job);  // false once finished and drained
  void finish();
  bool empty() const;
  bool is_done() const { return done_.load() && empty(); }

 private:
  mutable std::mutex mutex_;
  std::condition_variable cv_;
  std::queue<FFmpegJob> jobs_;
  std::atomic<bool> done_{false};
};

// "file '<abs>'\ninpoint %.2f\noutpoint %.2f\n" per segment, segments with end <= start dropped.
std::string build_concat_list(const std::string& input_path, const std::vector<TimeSegment>& segments);

// Runs `ffmpeg -f concat … -c copy` like the reference when an ffmpeg binary exists
// ($MOTION_TRIM_FFMPEG, else /usr/local/bin/ffmpeg); otherwise writes the concat list to
// "<output>.concat.txt" so the cut decision is still observable. Returns 0 on success.
int execute_ffmpeg_cut(const std::string& input_path, const std::string& output_path,
                       const std::vector<TimeSegment>& segments, const std::vector<int>& cpu_set, int stream_id);

}  // namespace motion_trim
