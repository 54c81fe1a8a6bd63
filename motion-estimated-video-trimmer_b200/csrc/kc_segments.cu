// kc_segments.cu — K-C: per-video timestamp merge, gap-merged segments, savings and decision.
//
// Replaces the inline block of ProcessingPipeline::run (reference src/pipeline.cpp:297-404):
//   timestamps of flagged frames (motion_scanner.cpp:382-383)  → ordered warp-ballot compaction
//   std::sort + std::unique (:302-304)                         → skipped when already strictly
//                                                                 increasing, else bitonic + unique
//   segment builder (:325-344): `last_act` is always ts[i-1], so a segment starts wherever
//     ts[i]-ts[i-1] > MAX_GAP_SEC — an adjacent difference + prefix count, not a serial scan
//   clamp + savings (:349-356): clamps are element-wise; only the left-to-right f64 sum is serial
//   decision (:308-319, :358-404)
// One CTA per video; many videos per launch. All f64 arithmetic uses explicit round-to-nearest
// intrinsics (never contracted) in the reference's order, so results are bit-identical.
#include "common.cuh"
#include "kernels.cuh"

namespace mscan {

namespace {

constexpr int kSegThreads = 1024;
constexpr int kSegWarps = kSegThreads / 32;

struct BlockScan {
  uint32_t warp_tot[kSegWarps];
  uint32_t total;
};

// exclusive prefix of v over the CTA (thread order); *total_out = CTA sum. Two barriers.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, BlockScan& s, uint32_t* total_out) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (uint32_t)o) inc += t;
  }
  if (lane == 31) s.warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = s.warp_tot[lane];
    uint32_t winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= (uint32_t)o) winc += t;
    }
    s.warp_tot[lane] = winc - w;
    if (lane == 31) s.total = winc;
  }
  __syncthreads();
  const uint32_t r = s.warp_tot[warp] + inc - v;
  *total_out = s.total;
  return r;
}

// std::max(a,b) / std::min(a,b) exactly as libstdc++ defines them (matters for -0.0 / NaN)
__device__ __forceinline__ double max_like_std(double a, double b) { return (a < b) ? b : a; }
__device__ __forceinline__ double min_like_std(double a, double b) { return (b < a) ? b : a; }

__global__ void __launch_bounds__(kSegThreads, 1) kc_segments_kernel(const __grid_constant__ SegArgs a) {
  __shared__ BlockScan scan;
  __shared__ double lens[kSegThreads];
  __shared__ double s_sum;

  const SegJob job = a.jobs[blockIdx.x];
  const uint32_t tid = threadIdx.x;
  double* ts = a.ts_a + job.ts_base;
  double* ts2 = a.ts_b + job.ts_base;
  mscan_segment* segs = a.segs + job.seg_base;

  // ---- 1. ordered compaction of flagged pts over the video's extents --------------------------
  uint32_t n = 0;  // uniform across the CTA
  for (uint32_t e = job.ext_begin; e < job.ext_end; ++e) {
    const SegExtent ex = a.extents[e];
    for (uint64_t i0 = 0; i0 < ex.n; i0 += kSegThreads) {
      const uint64_t i = i0 + tid;
      const bool act = (i < ex.n) && (a.flags[ex.start + i] != 0);
      uint32_t tot;
      const uint32_t pos = block_excl_scan(act ? 1u : 0u, scan, &tot);
      if (act) ts[n + pos] = a.pts[ex.start + i];
      n += tot;
      __syncthreads();
    }
  }
  __syncthreads();

  // ---- 2. sort + unique unless already strictly increasing ------------------------------------
  bool disorder = false;
  for (uint32_t i = tid + 1; i < n; i += kSegThreads) disorder |= !(ts[i - 1] < ts[i]);
  if (__syncthreads_or(disorder)) {
    uint64_t P = 1;
    while (P < n) P <<= 1;  // <= job.ts_cap
    for (uint64_t i = n + tid; i < P; i += kSegThreads) ts[i] = __longlong_as_double(0x7FF0000000000000ll);
    __syncthreads();
    for (uint64_t k = 2; k <= P; k <<= 1) {
      for (uint64_t j = k >> 1; j > 0; j >>= 1) {
        for (uint64_t i = tid; i < P; i += kSegThreads) {
          const uint64_t ixj = i ^ j;
          if (ixj > i) {
            const double x = ts[i], y = ts[ixj];
            const bool up = (i & k) == 0;
            if (up ? (y < x) : (x < y)) {
              ts[i] = y;
              ts[ixj] = x;
            }
          }
        }
        __syncthreads();
      }
    }
    // unique: keep the first of every run of == values (:303-304)
    uint32_t m = 0;
    for (uint32_t i0 = 0; i0 < n; i0 += kSegThreads) {
      const uint32_t i = i0 + tid;
      const bool keep = (i < n) && (i == 0 || !(ts[i - 1] == ts[i]));
      uint32_t tot;
      const uint32_t pos = block_excl_scan(keep ? 1u : 0u, scan, &tot);
      if (keep) ts2[m + pos] = ts[i];
      m += tot;
      __syncthreads();
    }
    n = m;
    ts = ts2;
    __syncthreads();
  }

  mscan_video_result res;
  res.decision = MSCAN_NO_MOTION;
  res.n_motion_frames = n;
  res.n_segments = 0;
  res.reserved = 0;
  res.out_dur = 0.0;
  res.time_removed = 0.0;
  res.saved_pct = 0.0;
  if (n == 0) {  // :308-319 — no job, no output
    if (tid == 0) a.results[blockIdx.x] = res;
    return;
  }

  // ---- 3. segment heads by adjacent difference (:328-344) -------------------------------------
  uint32_t nseg = 0;
  for (uint32_t i0 = 0; i0 < n; i0 += kSegThreads) {
    const uint32_t i = i0 + tid;
    bool head = false;
    double cur = 0.0, prv = 0.0;
    if (i < n) {
      cur = ts[i];
      if (i == 0) head = true;
      else {
        prv = ts[i - 1];
        head = __dsub_rn(cur, prv) > a.max_gap;  // :329-330, strict
      }
    }
    uint32_t tot;
    const uint32_t pos = block_excl_scan(head ? 1u : 0u, scan, &tot);
    if (head) {
      const uint32_t k = nseg + pos;
      segs[k].start = max_like_std(0.0, __dsub_rn(cur, a.padding));   // :337 / :343
      if (i > 0) segs[k - 1].end = __dadd_rn(prv, a.padding);         // :338
    }
    nseg += tot;
    __syncthreads();
  }
  if (tid == 0) segs[nseg - 1].end = __dadd_rn(ts[n - 1], a.padding);  // :344
  __syncthreads();

  // ---- 4. clamp (element-wise) + out_dur (serial, left to right) (:349-354) -------------------
  if (tid == 0) s_sum = 0.0;
  for (uint32_t i0 = 0; i0 < nseg; i0 += kSegThreads) {
    const uint32_t i = i0 + tid;
    if (i < nseg) {
      mscan_segment s = segs[i];
      s.end = min_like_std(s.end, job.duration);
      s.start = min_like_std(s.start, s.end);
      segs[i] = s;
      lens[tid] = __dsub_rn(s.end, s.start);
    }
    __syncthreads();
    if (tid == 0) {
      double acc = s_sum;
      const uint32_t cnt = min((uint32_t)kSegThreads, nseg - i0);
      for (uint32_t j = 0; j < cnt; ++j) acc = __dadd_rn(acc, lens[j]);
      s_sum = acc;
    }
    __syncthreads();
  }
  if (tid == 0) {
    res.n_segments = nseg;
    res.out_dur = s_sum;
    res.time_removed = __dsub_rn(job.duration, s_sum);                                          // :355
    res.saved_pct = (job.duration > 0.0) ? __dmul_rn(__ddiv_rn(res.time_removed, job.duration), 100.0) : 0.0;  // :356
    res.decision = (res.saved_pct > a.min_savings_pct) ? MSCAN_CUT : MSCAN_FULL_COPY;           // :358
    a.results[blockIdx.x] = res;
  }
}

}  // namespace

cudaError_t segments_launch(const SegArgs& a, uint32_t n_videos, cudaStream_t st) {
  if (n_videos == 0) return cudaSuccess;
  kc_segments_kernel<<<n_videos, kSegThreads, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace mscan
