// pipeline.hpp — host mirror of ProcessingPipeline (reference include/motion_trim/pipeline.hpp:66-147):
// same constructor arguments, run(), set_ffmpeg_queue() and getters. run() keeps the reference's
// phases — map file, probe, chunk queue, worker threads, decision, FFmpegJob — with the scan and the
// merge/segment/decision block (src/pipeline.cpp:297-404) done by the GPU context it is given.
#pragma once

#include <string>
#include <vector>

#include "ffmpeg_queue.hpp"
#include "memory_io.hpp"
#include "motionscan.h"
#include "types.hpp"

namespace motion_trim {

class GpuPool;

class ProcessingPipeline {
 public:
  ProcessingPipeline(std::string in, std::string out, int stream_id = -1, int num_threads = 0,
                     std::vector<int> cpu_set = {});
  // which GPU context scans this video (required before run())
  void set_gpu(GpuPool* pool, int gpu_index);
  // Split this one video over several GPUs of the pool: chunk workers are bound to the
  // GPUs round-robin, every GPU scans the chunks its workers pull, and the per-frame results are stitched
  // on the first GPU (mscan_video_append_from, NVLink peer copy) before the segment step.
  void set_gpus(GpuPool* pool, std::vector<int> gpu_indices);
  void set_ffmpeg_queue(FFmpegQueue* q) { ffmpeg_queue_ = q; }

  int run();  // 0 ok (including "no motion"), 1 failure — as the reference

  double get_duration() const { return duration_; }
  double get_time_removed() const { return time_removed_; }
  double get_saved_pct() const { return saved_pct_; }
  int get_decision() const { return decision_; }
  const std::vector<TimeSegment>& get_segments() const { return segments_; }
  uint64_t frames_scanned() const { return frames_scanned_; }
  // wall seconds per phase of run() (the reference's TIMER_START/END records, src/pipeline.cpp:95-292)
  struct Phases {
    double map = 0, probe = 0, pin = 0, scan = 0, segments = 0, unpin = 0;
  };
  const Phases& phases() const { return phases_; }
  // Hands the input mapping to the caller (the batch processor unmaps finished inputs on a side thread:
  // tearing down a 720 MB mapping costs ≈ 9 ms, more than projecting its records).
  MappedFile release_input() { return std::move(file_buffer_); }
  uint64_t records_scanned() const { return records_scanned_; }

 private:
  MappedFile file_buffer_;
  std::string input_path_, output_path_;
  double duration_ = 0, time_removed_ = 0, saved_pct_ = 0;
  int decision_ = MSCAN_NO_MOTION;
  std::vector<TimeSegment> segments_;
  uint64_t frames_scanned_ = 0, records_scanned_ = 0;
  Phases phases_;
  int stream_id_;
  int num_threads_;
  std::vector<int> cpu_set_;
  GpuPool* pool_ = nullptr;
  int gpu_index_ = 0;              // the GPU that runs the segment step (first of gpus_)
  std::vector<int> gpus_;          // all GPUs scanning this video
  FFmpegQueue* ffmpeg_queue_ = nullptr;
};

}  // namespace motion_trim
