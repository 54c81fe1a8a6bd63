// common.cuh — sm_100a device helpers shared by the motion-scan kernels (mbarrier, bulk async copy,
// named barriers, warp reductions). Inline PTX only; no CUTLASS/CUB.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace mscan {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait returns true once the phase with the given parity has completed. kHint: pass a suspend-time hint — the thread
// may be suspended up to that long while the phase is incomplete, which suits the projected layouts (many resident warps,
// issue-bound); the HBM-bound native layout polls without it (one CTA per SM at 4K: nothing hides a late wake-up).
constexpr uint32_t kTryWaitHintNs = 20000;
template <bool kHint = false>
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  if (kHint) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(kTryWaitHintNs)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
  return ok != 0;
}
// kSleep: poll with a short sleep, so that a starved warp does not burn issue slots the busy warps of the co-resident
// CTAs could use (the try_wait loop was 18 % of all issued instructions of K-A<packed>, which is bound by issue
// slots). The HBM-bound native layout keeps the plain spin: there the sleep only adds latency (8K: 7.1 → 6.8 TB/s).
template <bool kSleep = false>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait<kSleep>(bar, parity)) return;
  while (!mbar_try_wait<kSleep>(bar, parity)) {
    if (kSleep) __nanosleep(40);  // (an exponential back-off to 256 ns was tried: no gain at 1080p, −18 % on the dense 4K field)
  }
}

// ---- 1-D bulk async copy global → shared (TMA engine, SASS UBLKCP) -----------------------------
// dst/src 16-byte aligned, bytes a multiple of 16; completion is signalled on `bar` as tx bytes.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                         uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}

// ---- named barriers over a subset of the CTA's warps -------------------------------------------
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}
__device__ __forceinline__ bool named_bar_or(uint32_t id, uint32_t n_threads, bool pred) {
  uint32_t out;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.u32 q, %3, 0;\n\t"
      "bar.red.or.pred p, %1, %2, q;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(out)
      : "r"(id), "r"(n_threads), "r"((uint32_t)pred)
      : "memory");
  return out != 0;
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace mscan
