// ka_scan_cluster.cu — K-A for block grids that do not fit one CTA's shared memory (8K: 480x270 = 129 600
// cells, 16K: 518 400): a thread-block cluster of 2/4/8/16 CTAs scans one frame together.
//
// Same computation as ka_scan.cu (check_frame, reference src/motion_scanner.cpp:217-295); what changes is
// where the vote grid lives and who reads which records:
//   * the grid is distributed over the cluster's shared memory by row bands: CTA r owns rows
//     [r*rpr, (r+1)*rpr), rpr = ceil(gh / C), as 16-bit counters with ka_scan.cu's carry guard: a warp adds at
//     most 32 per trip, so at most 32 x 16 warps x 16 CTAs = 8 192 can slip past the 0x7FFF threshold, far from
//     the 0x8000 that would carry into the word-mate;
//   * the frame's records are cut into C contiguous slices, CTA r streams slice r through its own
//     bulk-copy ring (HBM is read once); export_mvs order is raster order, so most votes of slice r land in
//     band r — the rest go to the owning CTA with a DSMEM atomic (red.shared::cluster);
//   * Phase 2 runs per band; the bit-row just outside a band comes from the neighbour CTA over DSMEM, the
//     per-band cluster counts are summed in CTA 0's shared memory;
//   * three cluster barriers per frame order votes → bit-rows → counts; frames are handed out by an atomic
//     queue popped by CTA 0 and published to the cluster one frame ahead.
// Before this kernel such grids used per-CTA counters in global memory (L2 atomics): 2.3 TB/s at 8K against
// 7.5 TB/s for grids that fit (tools/ka_bigframe.py). That path remains for grids beyond 16 CTAs of shared memory.
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"

namespace mscan {

namespace {

constexpr int kTileRec = 512;
constexpr int kTileBytes = kTileRec * kRecBytes;  // 20480; holds 2560 projected records
constexpr int kConsWarps = 16;
constexpr int kCons = kConsWarps * 32;
constexpr int kThreads = kCons + 32;

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ uint32_t ld_cluster(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_cluster(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_cluster(uint32_t addr, uint32_t v) {
  asm volatile("red.shared::cluster.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

struct __align__(16) SliceTile {
  uint32_t n_rec;     // 0 = end of this CTA's slice
  uint32_t byte_off;  // first record's offset inside the 16-byte aligned copy
  uint32_t pad0, pad1;
};

template <bool kPacked>
__global__ void __launch_bounds__(kThreads, 1) ka_scan_cluster_kernel(const __grid_constant__ ScanArgs a) {
  constexpr uint32_t kStride = kPacked ? kPackedBytes : kRecBytes;
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t stages = a.stages;
  unsigned char* ring = smem;
  SliceTile* desc = reinterpret_cast<SliceTile*>(ring + (size_t)stages * kTileBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(desc + stages);  // full[stages], empty[stages]
  uint32_t* mailbox = reinterpret_cast<uint32_t*>(bars + 2 * stages);  // [2] next frame index, [2] = count accumulator (CTA 0)
  uint32_t* bits = mailbox + 4;                                        // [max_bit_words] this band's active bit-rows
  uint32_t* cnt = bits + a.max_bit_words;                              // [(max_cells+1)/2] 16-bit counters of this band

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_rank(), C = cluster_size();
  const uint32_t bar_full0 = smem_u32(bars), bar_empty0 = smem_u32(bars + stages);

  if (tid == 0) {
    for (uint32_t s = 0; s < stages; ++s) {
      mbar_init(bar_full0 + 8 * s, 1);
      mbar_init(bar_empty0 + 8 * s, kConsWarps);
    }
    mbar_fence_init();
    mailbox[2] = 0;
  }
  for (uint32_t i = tid; i < (a.max_cells + 1) / 2; i += kThreads) cnt[i] = 0;
  __syncthreads();
  if (rank == 0 && tid == 0) {  // first frame of this cluster, published to every CTA
    const uint32_t f0 = atomicAdd(a.work, 1u);
    for (uint32_t p = 0; p < C; ++p) st_cluster(map_to_cta(smem_u32(&mailbox[0]), p), f0);
  }
  cluster_sync_all();

  uint32_t stage = 0, phase = 0;  // ring position: the producer lane and the consumers advance in step
  for (uint32_t par = 0;; par ^= 1u) {
    const uint32_t f = mailbox[par];
    if (f >= a.n_frames) break;
    const uint64_t o0 = __ldg(a.rec_off + f), o1 = __ldg(a.rec_off + f + 1);
    const uint32_t gi = a.frame_geom ? __ldg(a.frame_geom + f) : 0u;
    const int4 gq = __ldg(reinterpret_cast<const int4*>(a.geoms) + gi);
    const int32_t gw = gq.x, gh = gq.y, y_min = gq.z, y_max = gq.w;
    if (rank == 0 && tid == 0) {  // the frame after this one; visible cluster-wide after the next barrier
      const uint32_t fn = atomicAdd(a.work, 1u);
      for (uint32_t p = 0; p < C; ++p) st_cluster(map_to_cta(smem_u32(&mailbox[par ^ 1u]), p), fn);
    }
    const uint64_t n64 = o1 - o0;
    if (n64 == 0) {  // no MV side data ⇒ false (motion_scanner.cpp:219-221)
      if (rank == 0 && tid == 0) {
        a.flags[f] = 0;
        a.counts[f] = 0;
      }
      cluster_sync_all();
      continue;
    }
    const uint32_t rpr = ((uint32_t)gh + C - 1) / C;                 // rows per band
    // exact gy / rpr for gy < 2^16 and rpr >= 2; rpr == 1 (a grid with no more rows than the cluster has CTAs, e.g. a
    // QVGA video sharing a context with a 16K one) would wrap the reciprocal to 1: the owner is then gy itself
    const uint32_t rpr_inv = (uint32_t)(0x100000000ull / rpr) + 1u;
    const uint32_t row0 = rank * rpr;
    const uint32_t rows = row0 < (uint32_t)gh ? min(rpr, (uint32_t)gh - row0) : 0u;
    const uint32_t wpr = ((uint32_t)gw + 31u) >> 5;
    // this CTA's slice of the frame's records
    const uint64_t s0 = o0 + n64 * rank / C, s1 = o0 + n64 * (rank + 1) / C;
    const uint32_t n = (uint32_t)(s1 - s0);

    if (warp == 0) {
      // ================================ producer ================================================
      if (lane == 0) {
        const uint64_t policy = l2_policy_evict_first();
        if (n) {
          const uint64_t byte0 = s0 * (uint64_t)kStride;
          const uint32_t d = (uint32_t)(byte0 & 15u);
          const unsigned char* src = a.recs + (byte0 - d);
          const uint32_t n_tiles = kPacked ? (uint32_t)((d + (uint64_t)kPackedBytes * n + kTileBytes - 1) / kTileBytes)
                                           : (n + kTileRec - 1) / kTileRec;
          for (uint32_t t = 0; t < n_tiles; ++t) {
            uint32_t nr, bytes, boff;
            if (kPacked) {
              const uint32_t r_lo = t ? (uint32_t)(((uint64_t)t * kTileBytes - d) / kPackedBytes) : 0u;
              const uint32_t r_hi = (uint32_t)min((uint64_t)n, ((uint64_t)(t + 1) * kTileBytes - d) / kPackedBytes);
              nr = r_hi - r_lo;
              boff = t ? 0u : d;
              bytes = (t + 1 < n_tiles) ? (uint32_t)kTileBytes : ((boff + kPackedBytes * nr + 15u) & ~15u);
            } else {
              nr = min((uint32_t)kTileRec, n - t * kTileRec);
              boff = d;
              bytes = (t + 1 < n_tiles) ? (uint32_t)kTileBytes : ((d + kRecBytes * (nr - 1) + 16u + 15u) & ~15u);
            }
            mbar_wait<kPacked>(bar_empty0 + 8 * stage, phase ^ 1u);
            desc[stage] = SliceTile{nr, boff, 0u, 0u};
            mbar_arrive_expect_tx(bar_full0 + 8 * stage, bytes);
            bulk_g2s(smem_u32(ring + (size_t)stage * kTileBytes), src + (size_t)t * kTileBytes, bytes, bar_full0 + 8 * stage, policy);
            if (++stage == stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        mbar_wait<kPacked>(bar_empty0 + 8 * stage, phase ^ 1u);  // end-of-slice marker
        desc[stage] = SliceTile{0u, 0u, 0u, 0u};
        mbar_arrive(bar_full0 + 8 * stage);
        if (++stage == stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      stage = __shfl_sync(0xffffffffu, stage, 0);
      phase = __shfl_sync(0xffffffffu, phase, 0);
    } else {
      // ================================ consumers: Phase 1 (:242-268) ============================
      const uint32_t cwarp = (tid - 32) >> 5;
      const int32_t ithr = a.ithr, shift = a.shift;
      const bool keep_any = a.keep_none == 0;
      const uint32_t live_rows = (uint32_t)(y_max - y_min);
      const uint32_t cnt_addr = smem_u32(cnt);
      while (true) {
        mbar_wait<kPacked>(bar_full0 + 8 * stage, phase);
        const SliceTile td = desc[stage];
        if (td.n_rec) {
          const unsigned char* base = ring + (size_t)stage * kTileBytes + td.byte_off;
          for (uint32_t r0 = cwarp * 32; r0 < td.n_rec; r0 += kCons) {
            const uint32_t r = r0 + lane;
            int32_t key = -1, gx = 0, gy = 0;
            if (r < td.n_rec) {
              int32_t sx, sy, tx, ty;
              if (kPacked) {
                const uint2 w = *reinterpret_cast<const uint2*>(base + (size_t)r * kPackedBytes);
                sx = (int32_t)(int16_t)(w.x & 0xFFFFu);
                sy = (int32_t)w.x >> 16;
                tx = (int32_t)(int16_t)(w.y & 0xFFFFu);
                ty = (int32_t)w.y >> 16;
              } else {
                const unsigned char* p = base + (size_t)r * kRecBytes;
                const uint32_t w1 = *reinterpret_cast<const uint32_t*>(p + 4);
                const uint2 w23 = *reinterpret_cast<const uint2*>(p + 8);
                sx = (int32_t)w1 >> 16;
                sy = (int32_t)(int16_t)(w23.x & 0xFFFFu);
                tx = (int32_t)w23.x >> 16;
                ty = (int32_t)(int16_t)(w23.y & 0xFFFFu);
              }
              const int32_t dx = tx - sx, dy = ty - sy;                                                   // :246-247
              const int32_t mag = (int32_t)((uint32_t)dx * (uint32_t)dx + (uint32_t)dy * (uint32_t)dy);   // :248
              gx = tx >> shift;                                                                           // :255-256
              gy = ty >> shift;
              const bool in = ((uint32_t)gx < (uint32_t)gw) && ((uint32_t)(gy - y_min) < live_rows);      // :262
              if (keep_any && mag >= ithr && in) key = gy * gw + gx;                                      // :251
            }
            const int32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
            const bool head = (lane == 0) || (key != prev);
            const uint32_t heads = __ballot_sync(0xffffffffu, head);
            if (head && key >= 0) {
              const uint32_t above = (lane == 31) ? 0u : (heads & (0xFFFFFFFEu << lane));
              const uint32_t next = above ? (uint32_t)(__ffs(above) - 1) : 32u;
              const uint32_t owner = rpr == 1u ? (uint32_t)gy : __umulhi((uint32_t)gy, rpr_inv);
              const uint32_t lkey = ((uint32_t)gy - owner * rpr) * (uint32_t)gw + (uint32_t)gx;
              const uint32_t sh = (lkey & 1u) * 16u;
              const uint32_t word = cnt_addr + 4u * (lkey >> 1);
              if (owner == rank) {
                uint32_t* w = &cnt[lkey >> 1];
                if (((*reinterpret_cast<volatile uint32_t*>(w) >> sh) & 0xFFFFu) < 0x7FFFu) atomicAdd(w, (next - lane) << sh);
              } else {
                const uint32_t remote = map_to_cta(word, owner);
                if (((ld_cluster(remote) >> sh) & 0xFFFFu) < 0x7FFFu) red_add_cluster(remote, (next - lane) << sh);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty0 + 8 * stage);
        if (++stage == stages) {
          stage = 0;
          phase ^= 1u;
        }
        if (td.n_rec == 0) break;
      }
    }
    cluster_sync_all();  // #1: every vote of the frame, local and remote, has landed

    // ---- Phase 2 (:272-294), pass 1: this band's counters → bit-rows, counters re-zeroed (:229)
    if (warp != 0) {
      const uint32_t cwarp = (tid - 32) >> 5;
      for (uint32_t y = cwarp; y < rows; y += kConsWarps)
        for (uint32_t w = 0; w < wpr; ++w) {
          const uint32_t x = w * 32 + lane;
          uint32_t c = 0;
          const bool valid = x < (uint32_t)gw;
          if (valid) {
            uint16_t* h = reinterpret_cast<uint16_t*>(cnt) + (y * (uint32_t)gw + x);
            c = *h;
            *h = 0;
          }
          const uint32_t word = __ballot_sync(0xffffffffu, valid && c >= a.vec_need);  // :282
          if (lane == 0) bits[y * wpr + w] = word;
        }
    }
    cluster_sync_all();  // #2: bit-rows visible to the neighbours; counters are clean for the next frame

    // ---- pass 2: centre cells of this band with an active 4-neighbour; rows just outside come over DSMEM
    if (warp != 0) {
      const uint32_t ctid = tid - 32;
      uint32_t total = 0;
      const uint32_t bits_addr = smem_u32(bits);
      const uint32_t up_rows = rank ? rpr : 0u;  // the band above is always full
      const uint32_t dn_row0 = row0 + rows;      // first row of the band below
      for (uint32_t i = ctid; i < rows * wpr; i += kCons) {
        const uint32_t ly = i / wpr, w = i % wpr;
        const uint32_t y = row0 + ly;
        if (y < (uint32_t)y_min || y >= (uint32_t)y_max) continue;
        const uint32_t A = bits[ly * wpr + w];
        if (!A) continue;
        const uint32_t Lw = w ? bits[ly * wpr + w - 1] : 0u;
        const uint32_t Rw = (w + 1 < wpr) ? bits[ly * wpr + w + 1] : 0u;
        uint32_t U = 0, D = 0, UL = 0, UR = 0, DL = 0, DR = 0;
        auto row_word = [&](bool above, uint32_t ww) -> uint32_t {  // word ww of row y-1 / y+1
          if (above) {
            if (ly) return bits[(ly - 1) * wpr + ww];
            if (!up_rows) return 0u;  // y == 0: out-of-grid neighbours are inactive
            return ld_cluster(map_to_cta(bits_addr + 4u * ((up_rows - 1) * wpr + ww), rank - 1));
          }
          if (ly + 1 < rows) return bits[(ly + 1) * wpr + ww];
          if (dn_row0 >= (uint32_t)gh) return 0u;
          return ld_cluster(map_to_cta(bits_addr + 4u * ww, rank + 1));
        };
        U = row_word(true, w);   // :286 idx-gw
        D = row_word(false, w);  // :286 idx+gw
        uint32_t nb = (A << 1) | (Lw >> 31) | (A >> 1) | (Rw << 31) | U | D;  // :284-286
        if (a.adj8) {  // extension (not in the reference): diagonal neighbours too
          UL = w ? row_word(true, w - 1) : 0u;
          UR = (w + 1 < wpr) ? row_word(true, w + 1) : 0u;
          DL = w ? row_word(false, w - 1) : 0u;
          DR = (w + 1 < wpr) ? row_word(false, w + 1) : 0u;
          nb |= (U << 1) | (UL >> 31) | (U >> 1) | (UR << 31) | (D << 1) | (DL >> 31) | (D >> 1) | (DR << 31);
        }
        uint32_t mask = 0xFFFFFFFFu;  // centre columns are 1 .. gw-2 (:280)
        if (w == 0) mask &= ~1u;
        const int32_t hi_bit = gw - 2 - (int32_t)(w * 32);
        if (hi_bit < 31) mask &= (hi_bit < 0) ? 0u : ((2u << hi_bit) - 1u);
        total += (uint32_t)__popc(A & nb & mask);
      }
      total = warp_sum(total);
      if (lane == 0 && total) red_add_cluster(map_to_cta(smem_u32(&mailbox[2]), 0), total);
    }
    cluster_sync_all();  // #3: every band's count is in CTA 0
    if (rank == 0 && tid == 0) {
      const uint32_t total = mailbox[2];
      mailbox[2] = 0;
      a.counts[f] = total;
      a.flags[f] = (total >= a.clust_need) ? 1 : 0;  // closed form of :288-289
    }
  }

  // self-resetting work queue: the last CTA to finish re-arms it for the next launch
  cluster_sync_all();  // no CTA may exit while a neighbour can still address its shared memory
  if (tid == 0) {
    __threadfence();
    const uint32_t done = atomicAdd(a.work + 1, 1u);
    if (done == gridDim.x - 1) {
      a.work[0] = 0;
      a.work[1] = 0;
    }
  }
}

}  // namespace

uint32_t scan_cluster_smem(uint32_t stages, uint32_t band_cells, uint32_t band_bit_words) {
  return stages * (uint32_t)kTileBytes + stages * (uint32_t)sizeof(SliceTile) + 2 * stages * 8u + 16u + band_bit_words * 4u +
         ((((band_cells + 1u) / 2u) * 4u + 15u) & ~15u) + 128u;
}

cudaError_t scan_cluster_configure(uint32_t smem_optin) {
  cudaError_t e = cudaFuncSetAttribute(ka_scan_cluster_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(ka_scan_cluster_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin);
  // clusters of 16 CTAs (16K grids with a deep ring) are beyond the portable size of 8
  if (e == cudaSuccess) e = cudaFuncSetAttribute(ka_scan_cluster_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(ka_scan_cluster_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  return e;
}

cudaError_t scan_cluster_launch(const ScanArgs& a, const ScanPlan& plan, int num_sms, cudaStream_t st) {
  if (a.n_frames == 0) return cudaSuccess;
  if (std::getenv("MSCAN_KA_FAIL_CLUSTER_LAUNCH")) return cudaErrorLaunchOutOfResources;  // exercises the caller's fallback
  if (a.packed == kLayoutMvz) return cudaErrorInvalidValue;  // the cluster kernel reads native and mv8 records only
  const uint32_t C = plan.cluster;
  uint32_t clusters = (uint32_t)num_sms / C;
  if (clusters > a.n_frames) clusters = a.n_frames;
  if (clusters == 0) clusters = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * C);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = plan.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (a.packed) return cudaLaunchKernelEx(&cfg, ka_scan_cluster_kernel<true>, a);
  return cudaLaunchKernelEx(&cfg, ka_scan_cluster_kernel<false>, a);
}

}  // namespace mscan
