// ffmpeg_frontend.cpp — see ffmpeg_frontend.hpp. Behaviour follows the reference's MotionScanner
// (src/motion_scanner.cpp), cited per step; the analysis it performs in-line is replaced by staging.
#ifdef MT_WITH_FFMPEG
#include "motion_trim/ffmpeg_frontend.hpp"

extern "C" {
#include <libavcodec/avcodec.h>
#include <libavformat/avformat.h>
#include <libavformat/avio.h>
#include <libavutil/error.h>
#include <libavutil/motion_vector.h>
}

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>

#include "motion_trim/config.hpp"

namespace motion_trim {

namespace {
constexpr int kIoBuffer = 256 * 1024;  // AVIO_BUFFER_SIZE (include/motion_trim/types.hpp:28-33)
static_assert(sizeof(AVMotionVector) == sizeof(mscan_mv), "mscan_mv mirrors AVMotionVector (40 bytes)");

using clk = std::chrono::steady_clock;
struct Lap {  // adds the time since construction (or the last lap) to a microsecond counter
  clk::time_point t = clk::now();
  void into(long& us) {
    const auto n = clk::now();
    us += (long)std::chrono::duration_cast<std::chrono::microseconds>(n - t).count();
    t = n;
  }
};
}  // namespace

FFmpegFrontEnd::FFmpegFrontEnd(const MappedFile& file) : file_(file) {}
FFmpegFrontEnd::~FFmpegFrontEnd() { close(); }

// AVIO over the mapping (the role of MemoryLoader::read/seek, src/memory_io.cpp:134-166)
int FFmpegFrontEnd::io_read(void* opaque, uint8_t* buf, int n) {
  Cursor* c = static_cast<Cursor*>(opaque);
  if (c->pos >= c->size) return AVERROR_EOF;
  const size_t take = std::min(c->size - c->pos, (size_t)std::max(n, 0));
  std::memcpy(buf, c->base + c->pos, take);
  c->pos += take;
  return (int)take;
}

int64_t FFmpegFrontEnd::io_seek(void* opaque, int64_t off, int whence) {
  Cursor* c = static_cast<Cursor*>(opaque);
  if (whence == AVSEEK_SIZE) return (int64_t)c->size;
  int64_t to;
  switch (whence) {
    case SEEK_SET: to = off; break;
    case SEEK_CUR: to = (int64_t)c->pos + off; break;
    case SEEK_END: to = (int64_t)c->size + off; break;
    default: to = (int64_t)c->pos; break;
  }
  c->pos = (size_t)std::clamp<int64_t>(to, 0, (int64_t)c->size);
  return (int64_t)c->pos;
}

bool FFmpegFrontEnd::open() {
  close();
  if (!file_.is_valid()) return false;
  cur_ = Cursor{file_.data(), file_.size(), 0};
  frame_ = av_frame_alloc();
  pkt_ = av_packet_alloc();
  unsigned char* buf = static_cast<unsigned char*>(av_malloc(kIoBuffer));
  if (!frame_ || !pkt_ || !buf) {
    av_free(buf);
    return false;
  }
  io_ = avio_alloc_context(buf, kIoBuffer, 0, &cur_, &FFmpegFrontEnd::io_read, nullptr, &FFmpegFrontEnd::io_seek);
  if (!io_) {
    av_free(buf);
    return false;
  }
  fmt_ = avformat_alloc_context();
  if (!fmt_) return false;
  fmt_->pb = io_;
  fmt_->flags |= AVFMT_FLAG_CUSTOM_IO;  // :96
  if (avformat_open_input(&fmt_, "RAM", nullptr, nullptr) < 0) {  // frees fmt_ itself on failure
    fmt_ = nullptr;
    return false;
  }
  if (avformat_find_stream_info(fmt_, nullptr) < 0) return false;
  stream_ = av_find_best_stream(fmt_, AVMEDIA_TYPE_VIDEO, -1, -1, nullptr, 0);
  if (stream_ < 0) return false;
  for (unsigned i = 0; i < fmt_->nb_streams; ++i)  // :119-123 — audio/subtitle packets are dropped by the demuxer
    if ((int)i != stream_) fmt_->streams[i]->discard = AVDISCARD_ALL;

  const AVCodecParameters* par = fmt_->streams[stream_]->codecpar;
  const AVCodec* codec = avcodec_find_decoder(par->codec_id);
  if (!codec) codec = avcodec_find_decoder_by_name(par->codec_id == AV_CODEC_ID_HEVC ? "hevc" : "h264");  // :128-133
  if (!codec) return false;
  dec_ = avcodec_alloc_context3(codec);
  if (!dec_ || avcodec_parameters_to_context(dec_, par) < 0) return false;
  // Pixels are never looked at: same decoder shortcuts as the reference (:145-165)
  dec_->skip_loop_filter = AVDISCARD_ALL;
  dec_->skip_idct = AVDISCARD_ALL;
  dec_->skip_frame = AVDISCARD_BIDIR;  // B-frames never reach the scan
  dec_->flags2 |= AV_CODEC_FLAG2_FAST;
  dec_->flags |= AV_CODEC_FLAG_GRAY;
  dec_->thread_count = 1;              // parallelism is per chunk, one decoder per worker
  dec_->thread_type = FF_THREAD_SLICE;
  AVDictionary* opts = nullptr;
  av_dict_set(&opts, "flags2", "+export_mvs", 0);  // :168-172 — without it there is no MV side data
  const int rc = avcodec_open2(dec_, codec, &opts);
  av_dict_free(&opts);
  return rc >= 0;
}

void FFmpegFrontEnd::close() {
  if (dec_) avcodec_free_context(&dec_);
  if (fmt_) {
    fmt_->pb = nullptr;  // custom IO stays ours: released below whatever avformat_close_input does with pb
    avformat_close_input(&fmt_);
  }
  if (io_) {
    av_free(io_->buffer);  // libavformat may have replaced the buffer it was given
    io_->buffer = nullptr;
    avio_context_free(&io_);
  }
  if (frame_) av_frame_free(&frame_);
  if (pkt_) av_packet_free(&pkt_);
  stream_ = -1;
}

double FFmpegFrontEnd::duration() const {
  return (fmt_ && fmt_->duration != AV_NOPTS_VALUE) ? fmt_->duration / (double)AV_TIME_BASE : 0.0;
}

double FFmpegFrontEnd::fps() const {
  if (!fmt_ || stream_ < 0) return 25.0;
  const AVRational r = fmt_->streams[stream_]->avg_frame_rate;
  return r.den > 0 ? av_q2d(r) : 25.0;
}

int FFmpegFrontEnd::width() const { return dec_ ? dec_->width : 0; }
int FFmpegFrontEnd::height() const { return dec_ ? dec_->height : 0; }

long FFmpegFrontEnd::scan(double start, double end, long& seek_us, long& decode_us, long& stage_us, size_t batch_records,
                          const std::function<bool(const StagedFrames&)>& sink) {
  if (!fmt_ || !dec_) return -1;
  const double time_base = av_q2d(fmt_->streams[stream_]->time_base);  // :304-305
  const double target = Config::target_fps(), video_fps = fps();
  const int every = (target > 0 && target < video_fps) ? (int)(video_fps / target) : 1;  // :310-313
  int decoded = 0;  // counts from the key frame the seek lands on, pre-range frames included (:314,357)
  long selected = 0;
  staged_.clear();

  Lap lap;
  if (start > 0) {  // :319-325
    av_seek_frame(fmt_, stream_, (int64_t)(start / time_base), AVSEEK_FLAG_BACKWARD);
    avcodec_flush_buffers(dec_);
  }
  lap.into(seek_us);

  auto flush = [&]() {
    if (staged_.pts.empty()) return true;
    const bool ok = sink(staged_);
    staged_.clear();
    return ok;
  };

  bool in_range = true;
  while (in_range && av_read_frame(fmt_, pkt_) >= 0) {
    if (pkt_->stream_index == stream_) {
      lap = Lap{};
      const int sent = avcodec_send_packet(dec_, pkt_);
      lap.into(decode_us);
      while (sent >= 0 && in_range) {
        const int got = avcodec_receive_frame(dec_, frame_);
        lap.into(decode_us);
        if (got < 0) break;
        if (++decoded % every != 0) continue;             // :357 TARGET_FPS skip
        const double pts = (double)frame_->pts * time_base;  // :361
        if (pts < start) continue;                        // :364
        if (pts >= end) {                                 // :368 — the range is over
          in_range = false;
          break;
        }
        // ---- where the reference analyses (:376): project the side data while it is cache-hot ----
        const AVFrameSideData* sd = av_frame_get_side_data(frame_, AV_FRAME_DATA_MOTION_VECTORS);
        const size_t n = sd ? sd->size / sizeof(AVMotionVector) : 0;  // :226
        const size_t at = staged_.recs.size();
        staged_.recs.resize(at + n);
        if (n) mscan_pack_records(reinterpret_cast<const mscan_mv*>(sd->data), n, staged_.recs.data() + at);
        staged_.pts.push_back(pts);
        staged_.counts.push_back((uint32_t)n);
        ++selected;
        lap.into(stage_us);
        if (staged_.recs.size() >= batch_records && !flush()) {
          av_packet_unref(pkt_);
          return -1;
        }
        lap = Lap{};
      }
    }
    av_packet_unref(pkt_);
  }
  return flush() ? selected : -1;
}

}  // namespace motion_trim
#endif  // MT_WITH_FFMPEG
