// motion_scanner.cpp — host mirror of MotionScanner. Frame selection follows the reference's
// scan_range exactly (src/motion_scanner.cpp:303-371): backward seek to the key frame at or before
// `start`, a frame-skip counter that starts at that key frame, pts = ticks * time_base, frames with
// pts < start dropped, the loop ends at the first pts >= end. What used to be check_frame per frame
// (:376) is one mscan_submit per call.
#include "motion_trim/motion_scanner.hpp"

#include <chrono>
#include <cstdlib>
#include <cstring>

#include "motion_trim/config.hpp"

namespace motion_trim {

namespace {
using clk = std::chrono::steady_clock;
long us_since(clk::time_point t0) { return (long)std::chrono::duration_cast<std::chrono::microseconds>(clk::now() - t0).count(); }
constexpr uint32_t kKey = 1u, kHasMvs = 2u;
}  // namespace

MotionScanner::MotionScanner(const MappedFile& data, mscan_ctx* gpu, uint32_t video_id)
    : file_(data), gpu_(gpu), video_id_(video_id) {}

bool MotionScanner::initialize() {
  ready_ = false;
  if (!gpu_) return false;
  const char* want = std::getenv("MOTION_TRIM_FRONTEND");
  const bool force_ff = want && !std::strcmp(want, "ffmpeg"), force_mvs = want && !std::strcmp(want, "mvs");
  if (!force_ff && view_.parse(file_)) {
    use_ffmpeg_ = false;
    ready_ = true;
    return true;
  }
  if (force_mvs) return false;
#ifdef MT_WITH_FFMPEG
  ff_ = std::make_unique<FFmpegFrontEnd>(file_);
  use_ffmpeg_ = ready_ = ff_->open();
  if (!ready_) ff_.reset();
  return ready_;
#else
  return false;  // media files need a build with -DMT_WITH_FFMPEG
#endif
}

double MotionScanner::get_duration() {
  if (!ready_) return 0.0;
#ifdef MT_WITH_FFMPEG
  if (use_ffmpeg_) return ff_->duration();
#endif
  return view_.duration_us / 1000000.0;
}

double MotionScanner::get_fps() {
#ifdef MT_WITH_FFMPEG
  if (ready_ && use_ffmpeg_) return ff_->fps();
#endif
  return (ready_ && view_.fps_den > 0 && view_.fps_num > 0) ? view_.fps_num / (double)view_.fps_den : 25.0;
}

int MotionScanner::width() const {
#ifdef MT_WITH_FFMPEG
  if (use_ffmpeg_ && ff_) return ff_->width();
#endif
  return view_.width;
}

int MotionScanner::height() const {
#ifdef MT_WITH_FFMPEG
  if (use_ffmpeg_ && ff_) return ff_->height();
#endif
  return view_.height;
}

#ifdef MT_WITH_FFMPEG
// Decode-fed variant: every batch of projected frames the front-end stages becomes one mscan_submit_packed.
long MotionScanner::scan_range_ffmpeg(double start, double end, long& seek_us, long& decode_us, uint64_t* first_frame) {
  pts_.clear();
  runs_.clear();
  long stage_us = 0;
  bool failed = false;
  const long n = ff_->scan(start, end, seek_us, decode_us, stage_us, size_t(1) << 20, [&](const StagedFrames& b) {
    uint64_t idx = 0;
    const int rc = mscan_submit_packed(gpu_, video_id_, (uint32_t)b.pts.size(), b.pts.data(), b.counts.data(),
                                       b.recs.empty() ? nullptr : b.recs.data(), &idx);
    if (rc != MSCAN_OK) {
      failed = true;
      return false;
    }
    runs_.push_back(Run{idx, (uint32_t)b.pts.size(), pts_.size()});
    pts_.insert(pts_.end(), b.pts.begin(), b.pts.end());
    return true;
  });
  decode_us += stage_us;  // staging is part of what the worker does per decoded frame
  if (failed || n < 0) return -1;
  if (first_frame) *first_frame = runs_.empty() ? 0 : runs_.front().first;
  return n;
}
#endif

long MotionScanner::scan_range_async(double start, double end, long& seek_us, long& decode_us, uint64_t* first_frame) {
  if (!ready_) return -1;
#ifdef MT_WITH_FFMPEG
  if (use_ffmpeg_) return scan_range_ffmpeg(start, end, seek_us, decode_us, first_frame);
#endif
  const double time_base = view_.tb_num / (double)view_.tb_den;  // av_q2d (:304-305)
  const double video_fps = get_fps();
  const double target = Config::target_fps();
  const int frame_skip = (target > 0 && target < video_fps) ? (int)(video_fps / target) : 1;  // :310-313
  int frame_count = 0;

  auto t0 = clk::now();
  uint32_t i = 0;
  if (start > 0) {  // :319-325 — AVSEEK_FLAG_BACKWARD: last key frame with pts <= seek_ts
    const int64_t seek_ts = (int64_t)(start / time_base);
    uint32_t best = 0;
    for (uint32_t k = 0; k < view_.n_frames; ++k) {
      if (!(view_.frames[k].flags & kKey)) continue;
      if (view_.frames[k].pts <= seek_ts) best = k;
      else break;
    }
    i = best;
  }
  seek_us += us_since(t0);

  t0 = clk::now();
  pts_.clear();
  counts_.clear();
  sel_.clear();
  for (; i < view_.n_frames; ++i) {
    if (++frame_count % frame_skip != 0) continue;       // :357
    const MvsFrame& f = view_.frames[i];
    const double pts = (double)f.pts * time_base;        // :361
    if (pts < start) continue;                           // :364
    if (pts >= end) break;                               // :368
    sel_.push_back(i);
    pts_.push_back(pts);
    counts_.push_back((f.flags & kHasMvs) ? f.n_records : 0u);  // no side data ⇒ 0 records (:219-221)
  }
  decode_us += us_since(t0);
  runs_.clear();
  if (first_frame) *first_frame = 0;
  if (pts_.empty()) return 0;
  // One submit per run of selected frames that are adjacent in the file (their records are then one
  // contiguous block). Without frame skipping the whole range is a single run.
  size_t k = 0;
  while (k < sel_.size()) {
    size_t run = 1;
    while (k + run < sel_.size() && sel_[k + run] == sel_[k + run - 1] + 1 &&
           view_.frames[sel_[k + run]].first_record ==
               view_.frames[sel_[k + run - 1]].first_record + view_.frames[sel_[k + run - 1]].n_records)
      ++run;
    uint64_t idx = 0;
    const int rc = mscan_submit(gpu_, video_id_, (uint32_t)run, pts_.data() + k, counts_.data() + k,
                                view_.records + view_.frames[sel_[k]].first_record, &idx);
    if (rc != MSCAN_OK) return -1;
    runs_.push_back(Run{idx, (uint32_t)run, k});
    k += run;
  }
  if (first_frame) *first_frame = runs_.front().first;
  return (long)pts_.size();
}

std::vector<double> MotionScanner::scan_range(double start, double end, long& seek_us, long& decode_us, long& analyze_us) {
  std::vector<double> ts;
  const long n = scan_range_async(start, end, seek_us, decode_us, nullptr);
  if (n <= 0) return ts;
  auto t0 = clk::now();
  std::vector<uint8_t> flags;
  for (const Run& r : runs_) {  // other workers of the same video may have submitted in between: read run by run
    flags.resize(r.n);
    if (mscan_collect_range(gpu_, video_id_, r.first, r.n, flags.data(), nullptr) != MSCAN_OK) break;
    for (uint32_t i = 0; i < r.n; ++i)
      if (flags[i]) ts.push_back(pts_[r.at + i]);  // :382-383
  }
  analyze_us += us_since(t0);
  return ts;
}

}  // namespace motion_trim
