// feed_harness.cpp — measurement harness (libmscan_feed.so): T producer threads play the decode workers of the
// reference (one MotionScanner per worker thread, include/motion_trim/motion_scanner.hpp:8-13; worker loop
// src/pipeline.cpp:186-235) in front of libmotionscan's C ABI.
//
// Each thread owns a contiguous share of the frames and, per frame (or small batch), does what a decode worker does:
//   1. "decode" — a stand-in for avcodec_receive_frame with flags2=+export_mvs (src/motion_scanner.cpp:168-172,347):
//      it reads a compact per-record input (8 B/record, as a decoder reads a compressed bitstream) and WRITES the
//      frame's full native 40-byte AVMotionVector records into its own side-data buffer, which is therefore
//      cache-hot and pageable, exactly like sd->data (src/motion_scanner.cpp:219-226);
//   2. the call that replaces check_frame(frame) (src/motion_scanner.cpp:376): mscan_submit of those native records.
// Timers follow the reference: a per-thread timer around the hot-path call only (its analyze timer,
// src/motion_scanner.cpp:375-380) and one around the stand-in, plus the wall clock of the whole run.
// Nothing of the scan itself happens here: this file only calls the public ABI (include/motionscan.h).
#include <pthread.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/motionscan.h"

extern "C" {

typedef struct mscan_feed_spec {
  int32_t n_threads;              /* producer ("decode worker") threads                                   */
  const int32_t* cpus;            /* [n_threads] CPU to pin each thread to, or NULL                       */
  uint32_t n_videos;
  const uint32_t* video_ids;      /* [n_videos] videos already open in the context                        */
  const uint64_t* video_frame_off;/* [n_videos+1] each video's frame range in the arrays below            */
  const double* pts;              /* [F]                                                                  */
  const uint32_t* rec_count;      /* [F]                                                                  */
  const uint64_t* rec_off;        /* [F+1] record index of each frame's first record                      */
  const void* source;             /* decoder stand-in input: mscan_mv8[N] (source_kind 1) or mscan_mv[N] (0) */
  int32_t source_kind;
  uint32_t frames_per_submit;     /* 1 = one call per frame, like check_frame(frame)                      */
  int32_t submit_kind;            /* 0: mscan_submit(native records); 1: mscan_pack_records + mscan_submit_packed;
                                     diagnosis only: 2 = stand-in alone (no call), 3 = stand-in + mscan_pack_records into a
                                     private buffer (no submit) */
  uint64_t* frame_index_out;      /* [F] index of every frame in its video's submission order, or NULL    */
} mscan_feed_spec;

typedef struct mscan_feed_result {
  double wall_s;         /* first thread start → last thread done (stand-in + submits)                    */
  double hot_max_s;      /* slowest thread's Σ time inside mscan_submit* (the reference's analyze timer)   */
  double hot_sum_s;
  double standin_max_s;  /* slowest thread's Σ decode stand-in time                                        */
  double standin_sum_s;
  uint64_t frames, records, submits;
  int32_t rc;            /* first non-zero status of any ABI call                                          */
  int32_t avx512;        /* stand-in ran its AVX-512 record writer                                         */
} mscan_feed_result;

int mscan_feed_run(mscan_ctx* ctx, const mscan_feed_spec* spec, mscan_feed_result* out);
/* the stand-in's record writer alone (tests): out[i] = native record expanded from in[i] */
int mscan_feed_expand(const mscan_mv8* in, uint64_t n, mscan_mv* out);

}  // extern "C"

namespace {

using clk = std::chrono::steady_clock;
inline double secs(clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); }

// One native record from its four coordinates — the fields export_mvs fills for a past-reference H.264 vector
// (SURVEY Appendix E): source -1, w = h = 16, flags 0, motion = (src - dst) * 4 quarter-pel, motion_scale 4.
inline void expand_scalar(const mscan_mv8* in, uint64_t n, mscan_mv* out) {
  for (uint64_t i = 0; i < n; ++i) {
    const mscan_mv8 r = in[i];
    uint64_t w[5];
    const uint64_t head = 0xFFFFFFFFull | (16ull << 32) | (16ull << 40) | ((uint64_t)(uint16_t)r.src_x << 48);
    w[0] = head;
    w[1] = (uint64_t)(uint16_t)r.src_y | ((uint64_t)(uint16_t)r.dst_x << 16) | ((uint64_t)(uint16_t)r.dst_y << 32);
    w[2] = 0;
    const int32_t mx = ((int32_t)r.src_x - (int32_t)r.dst_x) * 4, my = ((int32_t)r.src_y - (int32_t)r.dst_y) * 4;
    w[3] = (uint64_t)(uint32_t)mx | ((uint64_t)(uint32_t)my << 32);
    w[4] = 4;
    std::memcpy(reinterpret_cast<unsigned char*>(out) + 40 * i, w, 40);
  }
}

#if defined(__x86_64__)
// Same bytes, 8 records (one 64-byte load) → 320 bytes (five 64-byte stores) per iteration: every output byte is a
// coordinate byte (from the input vector), a motion byte (from the vector of 8 x {motion_x, motion_y}) or a constant,
// so one two-source byte permute (AVX-512 VBMI) plus an OR builds each output vector.
struct ExpandTables {
  alignas(64) uint8_t idx[5][64], konst[5][64];
  uint64_t mask[5];
};
const ExpandTables& expand_tables() {
  static const ExpandTables t = [] {
    ExpandTables x{};
    for (int k = 0; k < 5; ++k) {
      x.mask[k] = 0;
      for (int p = 0; p < 64; ++p) {
        const int b = 64 * k + p, i = b / 40, r = b % 40;
        x.idx[k][p] = 0;
        x.konst[k][p] = 0;
        if (r < 4) x.konst[k][p] = 0xFF;             // source = -1
        else if (r < 6) x.konst[k][p] = 16;          // w = h = 16
        else if (r < 14) {                           // src_x, src_y, dst_x, dst_y: the 8 input bytes of record i
          x.idx[k][p] = (uint8_t)(8 * i + (r - 6));
          x.mask[k] |= 1ull << p;
        } else if (r >= 24 && r < 32) {              // motion_x, motion_y: second source of vpermi2b
          x.idx[k][p] = (uint8_t)(64 + 8 * i + (r - 24));
          x.mask[k] |= 1ull << p;
        } else if (r == 32) x.konst[k][p] = 4;       // motion_scale = 4; flags and padding stay 0
      }
    }
    return x;
  }();
  return t;
}

__attribute__((target("avx512f,avx512bw,avx512vl,avx512vbmi"))) void expand_avx512(const mscan_mv8* in, uint64_t n, mscan_mv* out) {
  const ExpandTables& t = expand_tables();
  unsigned char* o = reinterpret_cast<unsigned char*>(out);
  __m512i idx[5], kon[5];
  for (int k = 0; k < 5; ++k) {
    idx[k] = _mm512_load_si512(t.idx[k]);
    kon[k] = _mm512_load_si512(t.konst[k]);
  }
  const uint64_t blocks = n / 8;
  for (uint64_t g = 0; g < blocks; ++g) {
    const __m512i q = _mm512_loadu_si512(in + 8 * g);  // per record: src_x | src_y<<16 | dst_x<<32 | dst_y<<48
    // motion = (src - dst) * 4: 16-bit differences (exact for every |d| < 2^15), widened to two int32 per record
    const __m512i d16 = _mm512_sub_epi16(q, _mm512_srli_epi64(q, 32));
    const __m512i v3 = _mm512_slli_epi32(_mm512_cvtepi16_epi32(_mm512_cvtepi64_epi32(d16)), 2);
#pragma GCC unroll 5
    for (int k = 0; k < 5; ++k) {
      const __m512i w = _mm512_or_si512(_mm512_maskz_permutex2var_epi8(t.mask[k], q, idx[k], v3), kon[k]);
      _mm512_storeu_si512(o + 320 * g + 64 * k, w);
    }
  }
  if (n > 8 * blocks) expand_scalar(in + 8 * blocks, n - 8 * blocks, reinterpret_cast<mscan_mv*>(o + 320 * blocks));
}
#endif

bool have_avx512() {
#if defined(__x86_64__)
  return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
         __builtin_cpu_supports("avx512vbmi") && !(std::getenv("MSCAN_FEED_NO_AVX512") && std::getenv("MSCAN_FEED_NO_AVX512")[0] == '1');
#else
  return false;
#endif
}

inline void expand(const mscan_mv8* in, uint64_t n, mscan_mv* out, bool avx512) {
#if defined(__x86_64__)
  if (avx512) return expand_avx512(in, n, out);
#endif
  (void)avx512;
  expand_scalar(in, n, out);
}

struct ThreadOut {
  double hot = 0, standin = 0;
  uint64_t frames = 0, records = 0, submits = 0;
  int rc = 0;
  clk::time_point t_begin, t_end;
};

}  // namespace

extern "C" int mscan_feed_expand(const mscan_mv8* in, uint64_t n, mscan_mv* out) {
  if (n && (!in || !out)) return MSCAN_ERR_INVALID;
  expand(in, n, out, have_avx512());
  return MSCAN_OK;
}

extern "C" int mscan_feed_run(mscan_ctx* ctx, const mscan_feed_spec* sp, mscan_feed_result* out) {
  if (!ctx || !sp || !out || sp->n_threads < 1 || !sp->video_frame_off || !sp->video_ids || !sp->pts || !sp->rec_count ||
      !sp->rec_off || !sp->source)
    return MSCAN_ERR_INVALID;
  const int T = sp->n_threads;
  const uint64_t F = sp->video_frame_off[sp->n_videos];
  const uint64_t N = sp->rec_off[F];
  const uint32_t batch = std::max(1u, sp->frames_per_submit);
  const bool avx512 = have_avx512();
  // contiguous shares with about the same number of records each (the reference cuts the file into time ranges)
  std::vector<uint64_t> cut((size_t)T + 1, F);
  cut[0] = 0;
  for (int t = 1; t < T; ++t) {
    const uint64_t want = N / (uint64_t)T * (uint64_t)t;
    cut[(size_t)t] = (uint64_t)(std::lower_bound(sp->rec_off, sp->rec_off + F + 1, want) - sp->rec_off);
    cut[(size_t)t] = std::min(std::max(cut[(size_t)t], cut[(size_t)t - 1]), F);
  }
  uint32_t max_frame = 0;
  for (uint64_t f = 0; f < F; ++f) max_frame = std::max(max_frame, sp->rec_count[f]);

  std::vector<ThreadOut> res((size_t)T);
  std::atomic<int> ready{0};
  std::atomic<bool> go{false};
  std::vector<std::thread> th;
  for (int t = 0; t < T; ++t)
    th.emplace_back([&, t] {
      if (sp->cpus) {
        cpu_set_t set;
        CPU_ZERO(&set);
        CPU_SET(sp->cpus[t], &set);
        pthread_setaffinity_np(pthread_self(), sizeof set, &set);  // like system.cpp:211-225
      }
      ThreadOut& r = res[(size_t)t];
      // the decoder-owned side-data buffer (pageable, reused for every batch → stays in this core's cache)
      const size_t cap = (size_t)max_frame * batch;
      mscan_mv* sd = static_cast<mscan_mv*>(std::aligned_alloc(64, std::max<size_t>(64, (cap * sizeof(mscan_mv) + 127) & ~size_t(63))));
      mscan_mv8* packed = (sp->submit_kind == 1 || sp->submit_kind == 3) ? static_cast<mscan_mv8*>(std::aligned_alloc(64, std::max<size_t>(64, (cap * 8 + 63) & ~size_t(63)))) : nullptr;
      if (!sd || ((sp->submit_kind == 1 || sp->submit_kind == 3) && !packed)) r.rc = MSCAN_ERR_NOMEM;
      if (sd) std::memset(sd, 0, cap * sizeof(mscan_mv));
      ready.fetch_add(1);
      while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
      r.t_begin = clk::now();
      uint64_t f = cut[(size_t)t];
      const uint64_t f_end = cut[(size_t)t + 1];
      // which video does f belong to?
      uint32_t v = 0;
      while (v + 1 < sp->n_videos && sp->video_frame_off[v + 1] <= f) ++v;
      while (f < f_end && r.rc == 0) {
        while (sp->video_frame_off[v + 1] <= f) ++v;
        const uint64_t e = std::min<uint64_t>({f + batch, f_end, sp->video_frame_off[v + 1]});
        const uint64_t r0 = sp->rec_off[f], nrec = sp->rec_off[e] - r0;
        // 1. decode stand-in: write the batch's native records into the side-data buffer
        const auto t0 = clk::now();
        if (sp->source_kind == 1) expand(static_cast<const mscan_mv8*>(sp->source) + r0, nrec, sd, avx512);
        else std::memcpy(sd, static_cast<const mscan_mv*>(sp->source) + r0, nrec * sizeof(mscan_mv));
        const auto t1 = clk::now();
        // 2. the hot-path call
        uint64_t first = 0;
        int rc;
        if (sp->submit_kind == 1) {
          rc = mscan_pack_records(sd, nrec, packed);
          if (rc == MSCAN_OK) rc = mscan_submit_packed(ctx, sp->video_ids[v], (uint32_t)(e - f), sp->pts + f, sp->rec_count + f, packed, &first);
        } else if (sp->submit_kind == 2) {
          rc = MSCAN_OK;
        } else if (sp->submit_kind == 3) {
          rc = mscan_pack_records(sd, nrec, packed);
        } else {
          rc = mscan_submit(ctx, sp->video_ids[v], (uint32_t)(e - f), sp->pts + f, sp->rec_count + f, sd, &first);
        }
        const auto t2 = clk::now();
        r.standin += secs(t0, t1);
        r.hot += secs(t1, t2);
        if (rc != MSCAN_OK) r.rc = rc;
        if (sp->frame_index_out)
          for (uint64_t k = f; k < e; ++k) sp->frame_index_out[k] = first + (k - f);
        r.frames += e - f;
        r.records += nrec;
        r.submits += 1;
        f = e;
      }
      r.t_end = clk::now();
      std::free(sd);
      std::free(packed);
    });
  while (ready.load() < T) std::this_thread::yield();
  go.store(true, std::memory_order_release);
  for (auto& x : th) x.join();

  mscan_feed_result o{};
  clk::time_point b = res[0].t_begin, e = res[0].t_end;
  for (const ThreadOut& r : res) {
    b = std::min(b, r.t_begin);
    e = std::max(e, r.t_end);
    o.hot_max_s = std::max(o.hot_max_s, r.hot);
    o.hot_sum_s += r.hot;
    o.standin_max_s = std::max(o.standin_max_s, r.standin);
    o.standin_sum_s += r.standin;
    o.frames += r.frames;
    o.records += r.records;
    o.submits += r.submits;
    if (r.rc && !o.rc) o.rc = r.rc;
  }
  o.wall_s = secs(b, e);
  o.avx512 = avx512 ? 1 : 0;
  *out = o;
  return o.rc;
}
