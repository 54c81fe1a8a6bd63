// ffmpeg_frontend.hpp — decode front-end of the host mirror, compiled only with
// -DMT_WITH_FFMPEG (libavformat / libavcodec / libavutil headers + libraries needed).
//
// It is the producer side of the drop-in boundary: FFmpeg demux + decode with flags2=+export_mvs exactly
// as the reference configures it (src/motion_scanner.cpp:62-202), the reference's frame selection
// (:303-371), and — where the reference calls check_frame(frame) (:376) — the frame's motion-vector
// side data is projected to mscan_mv8 while it is still hot in the decoding core's cache and staged for
// mscan_submit_packed. No analysis happens on the host.
#pragma once
#ifdef MT_WITH_FFMPEG

#include <cstdint>
#include <functional>
#include <vector>

#include "memory_io.hpp"
#include "motionscan.h"

struct AVFormatContext;
struct AVCodecContext;
struct AVIOContext;
struct AVFrame;
struct AVPacket;

namespace motion_trim {

// Selected frames of one scan, staged for the GPU: records already projected.
struct StagedFrames {
  std::vector<double> pts;
  std::vector<uint32_t> counts;  // 0 ⇔ the frame had no MV side data (src/motion_scanner.cpp:219-221)
  std::vector<mscan_mv8> recs;
  void clear() {
    pts.clear();
    counts.clear();
    recs.clear();
  }
};

class FFmpegFrontEnd {
 public:
  explicit FFmpegFrontEnd(const MappedFile& file);
  ~FFmpegFrontEnd();
  FFmpegFrontEnd(const FFmpegFrontEnd&) = delete;
  FFmpegFrontEnd& operator=(const FFmpegFrontEnd&) = delete;

  bool open();
  double duration() const;  // src/motion_scanner.cpp:204-208
  double fps() const;       // :210-215
  int width() const;        // decoder (display) dimensions, the ones the grid is derived from (:189-192)
  int height() const;

  // Decodes [start, end) and hands the selected frames to `sink` in batches of about `batch_records`
  // projected records (and once more at the end of the range). Returns the number of selected frames,
  // or -1 if the sink refused a batch.
  long scan(double start, double end, long& seek_us, long& decode_us, long& stage_us, size_t batch_records,
            const std::function<bool(const StagedFrames&)>& sink);

 private:
  struct Cursor {  // read position of the AVIO callbacks inside the mapping
    const uint8_t* base;
    size_t size, pos;
  };
  static int io_read(void* opaque, uint8_t* buf, int n);
  static int64_t io_seek(void* opaque, int64_t off, int whence);
  void close();

  const MappedFile& file_;
  Cursor cur_{nullptr, 0, 0};
  AVIOContext* io_ = nullptr;
  AVFormatContext* fmt_ = nullptr;
  AVCodecContext* dec_ = nullptr;
  AVFrame* frame_ = nullptr;
  AVPacket* pkt_ = nullptr;
  int stream_ = -1;
  StagedFrames staged_;
};

}  // namespace motion_trim
#endif  // MT_WITH_FFMPEG
