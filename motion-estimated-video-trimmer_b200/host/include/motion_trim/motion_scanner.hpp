// motion_scanner.hpp — host mirror of the reference's MotionScanner
// (include/motion_trim/motion_scanner.hpp:61-153): same public surface — initialize(),
// get_duration(), get_fps(), scan_range(start, end, &seek_us, &decode_us, &analyze_us) →
// timestamps with motion — but check_frame (private, :106) is gone: selected frames are handed to
// the GPU context in one batch per scan_range call.
#pragma once

#include <cstdint>
#include <memory>
#include <vector>

#include "memory_io.hpp"
#include "motionscan.h"
#ifdef MT_WITH_FFMPEG
#include "ffmpeg_frontend.hpp"
#endif

namespace motion_trim {

class MotionScanner {
 public:
  // One scanner per worker thread, like the reference; all scanners of a video share video_id.
  MotionScanner(const MappedFile& data, mscan_ctx* gpu, uint32_t video_id);
  MotionScanner(const MotionScanner&) = delete;
  MotionScanner& operator=(const MotionScanner&) = delete;

  bool initialize();
  double get_duration();  // fmt_ctx->duration / AV_TIME_BASE, or 0.0 (src/motion_scanner.cpp:204-208)
  double get_fps();       // avg_frame_rate, 25.0 when unknown (:210-215)
  int width() const;
  int height() const;
  // Which producer feeds the GPU: the zero-copy MVS1 view, or (MT_WITH_FFMPEG builds) FFmpeg demux +
  // export_mvs decode. MOTION_TRIM_FRONTEND=ffmpeg|mvs forces one; by default MVS1 files take the view
  // and everything else goes to FFmpeg.
  bool uses_ffmpeg() const { return use_ffmpeg_; }

  // Contract-compatible: submits the range, waits for the GPU and returns the pts with motion.
  std::vector<double> scan_range(double start, double end, long& seek_us, long& decode_us, long& analyze_us);
  // Pipeline variant: submit only (asynchronous); K-C later reads the device-side log. Returns the
  // number of frames handed to the GPU, or -1 on error.
  long scan_range_async(double start, double end, long& seek_us, long& decode_us, uint64_t* first_frame = nullptr);

 private:
  const MappedFile& file_;
  mscan_ctx* gpu_;
  uint32_t video_id_;
  MvsView view_;
  bool ready_ = false;
  bool use_ffmpeg_ = false;
#ifdef MT_WITH_FFMPEG
  std::unique_ptr<FFmpegFrontEnd> ff_;
  long scan_range_ffmpeg(double start, double end, long& seek_us, long& decode_us, uint64_t* first_frame);
#endif
  struct Run {
    uint64_t first;  // index in the video's submission order
    uint32_t n;
    size_t at;       // offset into pts_
  };
  std::vector<double> pts_;
  std::vector<uint32_t> counts_;
  std::vector<uint32_t> sel_;  // file frame index of every selected frame
  std::vector<Run> runs_;
};

}  // namespace motion_trim
