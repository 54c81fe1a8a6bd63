// synth_host.cu — host side of the synthetic-stream harness (include/mvgen_core.h): the same
// generator the device kernels in synth.cu run, for tests, the CPU baseline and host-fed (e2e)
// benchmark legs. Pure host code; needs no GPU.
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/motionscan.h"
#include "../../include/mvgen_core.h"

static_assert(sizeof(mscan_mv) == 40, "mscan_mv must match AVMotionVector (40 bytes)");
static_assert(offsetof(mscan_mv, src_x) == 6 && offsetof(mscan_mv, src_y) == 8, "src_x@6 src_y@8");
static_assert(offsetof(mscan_mv, dst_x) == 10 && offsetof(mscan_mv, dst_y) == 12, "dst_x@10 dst_y@12");
static_assert(offsetof(mscan_mv, flags) == 16 && offsetof(mscan_mv, motion_x) == 24, "flags@16 motion_x@24");
static_assert(offsetof(mscan_mv, motion_scale) == 32, "motion_scale@32");

namespace {

uint32_t frame_count(const mvgen_spec& spec, uint64_t gframe) {
  mvgen_frame fr;
  mvgen_frame_init(&spec, gframe, &fr);
  if (fr.iframe) return 0;
  uint32_t acc = 0;
  for (int32_t my = 0; my < fr.mbh; ++my)
    for (int32_t mx = 0; mx < fr.mbw; ++mx) {
      mvgen_mb m;
      mvgen_mb_eval(&spec, &fr, mx, my, &m);
      acc += (uint32_t)m.nrec;
    }
  return acc;
}

void frame_fill(const mvgen_spec& spec, uint64_t gframe, mscan_mv* out) {
  mvgen_frame fr;
  mvgen_frame_init(&spec, gframe, &fr);
  if (fr.iframe) return;
  for (int32_t my = 0; my < fr.mbh; ++my)
    for (int32_t mx = 0; mx < fr.mbw; ++mx) {
      mvgen_mb m;
      mvgen_mb_eval(&spec, &fr, mx, my, &m);
      for (int32_t k = 0; k < m.nrec; ++k) {
        mvgen_rec r;
        mvgen_record(&spec, &m, mx, my, k, &r);
        mscan_mv v;
        std::memset(&v, 0, sizeof v);  // padding bytes are zero, like the device generator
        v.source = r.source;
        v.w = (uint8_t)r.w;
        v.h = (uint8_t)r.h;
        v.src_x = (int16_t)r.src_x;
        v.src_y = (int16_t)r.src_y;
        v.dst_x = (int16_t)r.dst_x;
        v.dst_y = (int16_t)r.dst_y;
        v.flags = 0;
        v.motion_x = r.motion_x;
        v.motion_y = r.motion_y;
        v.motion_scale = 4;
        std::memcpy(out++, &v, sizeof v);
      }
    }
}

template <typename F>
void parallel_frames(uint32_t n, int n_threads, F&& body) {
  if (n_threads < 1) n_threads = 1;
  if ((uint32_t)n_threads > n) n_threads = n ? (int)n : 1;
  if (n_threads == 1) {
    for (uint32_t f = 0; f < n; ++f) body(f);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t)
    th.emplace_back([&, t] {
      for (uint32_t f = (uint32_t)t; f < n; f += (uint32_t)n_threads) body(f);
    });
  for (auto& x : th) x.join();
}

}  // namespace

extern "C" {

int mscan_synth_preset(mvgen_spec* spec, int config, uint64_t seed) {
  if (!spec) return MSCAN_ERR_INVALID;
  std::memset(spec, 0, sizeof *spec);
  mvgen_preset(spec, config, seed);
  return MSCAN_OK;
}

int mscan_synth_host_counts(const mvgen_spec* spec, uint64_t frame0, uint32_t n_frames, uint32_t* rec_count,
                            int n_threads) {
  if (!spec || (n_frames && !rec_count)) return MSCAN_ERR_INVALID;
  parallel_frames(n_frames, n_threads, [&](uint32_t f) { rec_count[f] = frame_count(*spec, frame0 + f); });
  return MSCAN_OK;
}

int mscan_synth_host_fill(const mvgen_spec* spec, uint64_t frame0, uint32_t n_frames, const uint64_t* rec_off,
                          mscan_mv* recs, double* pts, int n_threads) {
  if (!spec || (n_frames && (!rec_off || !recs))) return MSCAN_ERR_INVALID;
  parallel_frames(n_frames, n_threads, [&](uint32_t f) {
    frame_fill(*spec, frame0 + f, recs + rec_off[f]);
    if (pts) pts[f] = mvgen_pts(spec, frame0 + f);
  });
  return MSCAN_OK;
}

}  // extern "C"
