// synth.cu — auxiliary kernels: exclusive scan of per-frame record counts (frame → record offsets,
// the CSR index K-A consumes) and device-side generation of the synthetic AVMotionVector streams
// defined in include/mvgen_core.h (measurement harness; lets the 10^9-record stream of
// BASELINE.json configs[4] exist without a 40 GB host buffer).
#include "common.cuh"
#include "kernels.cuh"

namespace mscan {

namespace {

constexpr int kAuxThreads = 1024;

__device__ __forceinline__ uint64_t block_excl_scan64(uint64_t v, uint64_t* warp_tot, uint64_t* total) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (uint32_t)o) inc += t;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const uint64_t w = warp_tot[lane];
    uint64_t winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint64_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= (uint32_t)o) winc += t;
    }
    warp_tot[lane] = winc - w;
    if (lane == 31) warp_tot[32] = winc;
  }
  __syncthreads();
  const uint64_t r = warp_tot[warp] + inc - v;
  *total = warp_tot[32];
  __syncthreads();
  return r;
}

// Single CTA, running carry. n is at most a few 10^5 frames per launch; this is not on the hot path.
__global__ void __launch_bounds__(kAuxThreads, 1) offsets_kernel(const uint32_t* __restrict__ counts, uint32_t n,
                                                                   uint64_t* __restrict__ off) {
  __shared__ uint64_t warp_tot[33];
  uint64_t carry = 0;
  for (uint32_t i0 = 0; i0 < n; i0 += kAuxThreads) {
    const uint32_t i = i0 + threadIdx.x;
    const uint64_t v = (i < n) ? counts[i] : 0;
    uint64_t tot;
    const uint64_t ex = block_excl_scan64(v, warp_tot, &tot);
    if (i < n) off[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) off[n] = carry;
}

// One CTA per frame: Σ records over the frame's macroblocks.
__global__ void __launch_bounds__(256) synth_counts_kernel(const __grid_constant__ mvgen_spec spec, uint64_t frame0,
                                                            uint32_t n_frames, uint32_t* __restrict__ counts) {
  __shared__ mvgen_frame fr;
  __shared__ uint32_t wsum[8];
  for (uint32_t f = blockIdx.x; f < n_frames; f += gridDim.x) {
    if (threadIdx.x == 0) mvgen_frame_init(&spec, frame0 + f, &fr);
    __syncthreads();
    uint32_t acc = 0;
    const int32_t n_mb = fr.mbw * fr.mbh;
    if (!fr.iframe) {
      for (int32_t mb = threadIdx.x; mb < n_mb; mb += blockDim.x) {
        mvgen_mb m;
        mvgen_mb_eval(&spec, &fr, mb % fr.mbw, mb / fr.mbw, &m);
        acc += (uint32_t)m.nrec;
      }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t t = 0;
      for (int w = 0; w < 8; ++w) t += wsum[w];
      counts[f] = t;
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void store_rec(mscan_mv* dst, const mvgen_rec& r) {
  // 40-byte record as five 8-byte stores (records are 8-byte aligned); padding bytes written as 0
  uint64_t* q = reinterpret_cast<uint64_t*>(dst);
  const uint64_t w0 = (uint64_t)(uint32_t)r.source | ((uint64_t)(uint8_t)r.w << 32) | ((uint64_t)(uint8_t)r.h << 40) |
                      ((uint64_t)(uint16_t)(int16_t)r.src_x << 48);
  const uint64_t w1 = (uint64_t)(uint16_t)(int16_t)r.src_y | ((uint64_t)(uint16_t)(int16_t)r.dst_x << 16) |
                      ((uint64_t)(uint16_t)(int16_t)r.dst_y << 32);
  const uint64_t w3 = (uint64_t)(uint32_t)r.motion_x | ((uint64_t)(uint32_t)r.motion_y << 32);
  q[0] = w0;
  q[1] = w1;
  q[2] = 0;  // flags
  q[3] = w3;
  q[4] = 4;  // motion_scale = 4 (quarter-pel), padding 0
}

// One CTA per frame: block scan of per-MB record counts, then every thread writes its MB's records.
__global__ void __launch_bounds__(256) synth_fill_kernel(const __grid_constant__ mvgen_spec spec, uint64_t frame0,
                                                          uint32_t n_frames, const uint64_t* __restrict__ rec_off,
                                                          mscan_mv* __restrict__ recs, double* __restrict__ pts) {
  __shared__ mvgen_frame fr;
  __shared__ uint32_t wtot[9];
  for (uint32_t f = blockIdx.x; f < n_frames; f += gridDim.x) {
    if (threadIdx.x == 0) {
      mvgen_frame_init(&spec, frame0 + f, &fr);
      if (pts) pts[f] = mvgen_pts(&spec, frame0 + f);
    }
    __syncthreads();
    const int32_t n_mb = fr.mbw * fr.mbh;
    uint64_t base = rec_off[f];
    if (!fr.iframe) {
      for (int32_t mb0 = 0; mb0 < n_mb; mb0 += blockDim.x) {
        const int32_t mb = mb0 + (int32_t)threadIdx.x;
        mvgen_mb m;
        m.nrec = 0;
        int32_t mx = 0, my = 0;
        if (mb < n_mb) {
          mx = mb % fr.mbw;
          my = mb / fr.mbw;
          mvgen_mb_eval(&spec, &fr, mx, my, &m);
        }
        // block exclusive scan of nrec
        const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        uint32_t inc = (uint32_t)m.nrec;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= (uint32_t)o) inc += t;
        }
        if (lane == 31) wtot[warp] = inc;
        __syncthreads();
        if (threadIdx.x == 0) {
          uint32_t run = 0;
          for (int w = 0; w < 8; ++w) {
            const uint32_t t = wtot[w];
            wtot[w] = run;
            run += t;
          }
          wtot[8] = run;
        }
        __syncthreads();
        const uint64_t at = base + wtot[warp] + inc - (uint32_t)m.nrec;
        for (int32_t k = 0; k < m.nrec; ++k) {
          mvgen_rec r;
          mvgen_record(&spec, &m, mx, my, k, &r);
          store_rec(recs + at + k, r);
        }
        base += wtot[8];
        __syncthreads();
      }
    }
    __syncthreads();
  }
}

// Device-side projection AVMotionVector → mscan_mv8 (bytes 6..13). Records are 8-byte aligned, so the
// range is the top 2 bytes of word 0 and the low 6 bytes of word 1.
__global__ void __launch_bounds__(256) project_kernel(const uint64_t* __restrict__ recs, uint64_t n,
                                                       uint64_t* __restrict__ out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t w0 = recs[5 * i], w1 = recs[5 * i + 1];
    out[i] = (w0 >> 48) | (w1 << 16);
  }
}

}  // namespace

cudaError_t project_launch(const mscan_mv* recs, uint64_t n, mscan_mv8* out, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  project_kernel<<<148 * 8, 256, 0, st>>>(reinterpret_cast<const uint64_t*>(recs), n, reinterpret_cast<uint64_t*>(out));
  return cudaGetLastError();
}

uint32_t offsets_scratch_elems(uint32_t) { return 0; }

cudaError_t offsets_launch(const uint32_t* counts, uint32_t n, uint64_t* off, uint64_t*, cudaStream_t st) {
  offsets_kernel<<<1, kAuxThreads, 0, st>>>(counts, n, off);
  return cudaGetLastError();
}

cudaError_t synth_counts_launch(const mvgen_spec& spec, uint64_t frame0, uint32_t n_frames, uint32_t* counts,
                                cudaStream_t st) {
  if (n_frames == 0) return cudaSuccess;
  const uint32_t grid = n_frames < 148u * 16u ? n_frames : 148u * 16u;
  synth_counts_kernel<<<grid, 256, 0, st>>>(spec, frame0, n_frames, counts);
  return cudaGetLastError();
}

cudaError_t synth_fill_launch(const mvgen_spec& spec, uint64_t frame0, uint32_t n_frames, const uint64_t* rec_off,
                              mscan_mv* recs, double* pts, cudaStream_t st) {
  if (n_frames == 0) return cudaSuccess;
  const uint32_t grid = n_frames < 148u * 16u ? n_frames : 148u * 16u;
  synth_fill_kernel<<<grid, 256, 0, st>>>(spec, frame0, n_frames, rec_off, recs, pts);
  return cudaGetLastError();
}

}  // namespace mscan
