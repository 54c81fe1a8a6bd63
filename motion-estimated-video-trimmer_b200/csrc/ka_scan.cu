// ka_scan.cu — K-A: per-frame vote scatter + 4-neighbour cluster count + activity flag.
//
// Replaces MotionScanner::check_frame (reference src/motion_scanner.cpp:217-295) for a batch of
// frames from many videos:
//   Phase 0  :229      memset(grid)                 → counters live in shared memory, re-zeroed by
//                                                      the epilogue pass that reads them
//   Phase 1  :242-268  per-record vote              → consumer warps, run-length pre-aggregated
//                                                      shared-memory atomicAdd
//   Phase 2  :272-294  cluster scan w/ early exit   → ballot to bit-rows + popc; reports the FULL
//                                                      count, flag = count >= max(1,CLUSTERS_NEEDED)
//
// Shape of the kernel (HBM-bound byte/integer work, no tensor cores):
//   * persistent CTAs (ctas_per_sm × #SM), frames handed out by an atomic queue (frames vary from
//     0 to 129 600 records);
//   * warp 0 is the producer: one lane streams the frame's native 40-byte records HBM → shared
//     memory with 1-D bulk async copies (cp.async.bulk → UBLKCP) into an mbarrier ring, running
//     ahead across frame boundaries so the epilogue of frame n overlaps the loads of frame n+1;
//   * 8 consumer warps read the 12 useful bytes of each record with LDS.32 + LDS.64 (stride 40 B is
//     bank-conflict free for 64-bit accesses), apply the integer threshold and bounds tests and
//     vote; consecutive records of one cell (8x8-split macroblocks export 4 in a row) are merged
//     with shfl/ballot so one lane issues one atomic for the run;
//   * counters are u32 (saturation at 255, :265-266, is unobservable for VECTORS_NEEDED <= 255);
//   * the epilogue is skipped for frames in which nothing voted (the common CCTV case).
#include "common.cuh"
#include "kernels.cuh"

namespace mscan {

namespace {

constexpr int kTileRec = 512;                      // records per ring stage
constexpr int kTileBytes = kTileRec * kRecBytes;   // 20480, multiple of lcm(16,40)=80
constexpr int kConsWarps = 8;
constexpr int kCons = kConsWarps * 32;
constexpr int kThreads = kCons + 32;
constexpr uint32_t kEndFrame = 0xFFFFFFFFu;
constexpr uint32_t kBarCons = 1;  // named barrier id of the consumer warps

struct __align__(16) TileDesc {
  uint32_t n_rec;     // records in this tile
  uint32_t byte_off;  // 0 or 8: first record's offset inside the 16-byte aligned copy
  uint32_t frame;     // frame index, kEndFrame = no more work
  uint32_t last;      // last tile of the frame
  int32_t gw, gh, y_min, y_max;
};

struct FrameMeta {
  uint64_t o0, o1;
  DevGeom g;
};

__device__ __forceinline__ FrameMeta load_meta(const ScanArgs& a, uint32_t f) {
  FrameMeta m;
  m.o0 = 0;
  m.o1 = 0;
  m.g = DevGeom{0, 0, 0, 0};
  if (f < a.n_frames) {
    m.o0 = __ldg(a.rec_off + f);
    m.o1 = __ldg(a.rec_off + f + 1);
    const uint32_t gi = a.frame_geom ? __ldg(a.frame_geom + f) : 0u;
    const int4 g = __ldg(reinterpret_cast<const int4*>(a.geoms) + gi);
    m.g = DevGeom{g.x, g.y, g.z, g.w};
  }
  return m;
}

// kGlobalCnt: the vote counters of grids too large for shared memory (8K/16K video) live in a
// per-CTA slice of a zero-initialised global scratch (L2-resident); everything else is identical.
template <bool kGlobalCnt>
__global__ void __launch_bounds__(kThreads, 2) ka_scan_kernel(const __grid_constant__ ScanArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t stages = a.stages;
  unsigned char* ring = smem;
  TileDesc* desc = reinterpret_cast<TileDesc*>(ring + (size_t)stages * kTileBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(desc + stages);  // full[stages], empty[stages]
  uint32_t* bits = reinterpret_cast<uint32_t*>(bars + 2 * stages);
  uint32_t* cnt = kGlobalCnt ? a.cnt_scratch + (size_t)blockIdx.x * a.max_cells : bits + 2 * a.max_bit_words;

  const uint32_t tid = threadIdx.x;
  const uint32_t warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_full0 = smem_u32(bars);
  const uint32_t bar_empty0 = smem_u32(bars + stages);

  if (tid == 0) {
    for (uint32_t s = 0; s < stages; ++s) {
      mbar_init(bar_full0 + 8 * s, 1);
      mbar_init(bar_empty0 + 8 * s, kConsWarps);
    }
    mbar_fence_init();
  }
  if (!kGlobalCnt)  // the global scratch is zero on entry and every epilogue leaves it zero
    for (uint32_t i = tid; i < a.max_cells; i += kThreads) cnt[i] = 0;
  __syncthreads();

  if (warp == 0) {
    // ================================ producer ================================================
    if (lane == 0) {
      const uint64_t policy = l2_policy_evict_first();
      uint32_t stage = 0, phase = 0;
      uint32_t f_cur = atomicAdd(a.work, 1u);
      FrameMeta m_cur = load_meta(a, f_cur);
      uint32_t f_next = atomicAdd(a.work, 1u);
      while (f_cur < a.n_frames) {
        // issue next frame's metadata loads and the queue pop after it now; they complete while
        // this frame's tiles stream
        const FrameMeta m_next = load_meta(a, f_next);
        const uint32_t f_next2 = (f_next < a.n_frames) ? atomicAdd(a.work, 1u) : f_next;
        const uint64_t n64 = m_cur.o1 - m_cur.o0;
        if (n64 == 0) {
          // rec_count == 0 ⇔ no MV side data ⇒ false (motion_scanner.cpp:219-221)
          a.flags[f_cur] = 0;
          a.counts[f_cur] = 0;
        } else {
          const uint32_t n = (uint32_t)n64;
          const uint64_t byte0 = m_cur.o0 * (uint64_t)kRecBytes;
          const uint32_t d = (uint32_t)(byte0 & 15u);
          const unsigned char* src = a.recs + (byte0 - d);
          const uint32_t n_tiles = (n + kTileRec - 1) / kTileRec;
          for (uint32_t t = 0; t < n_tiles; ++t) {
            const uint32_t nr = min((uint32_t)kTileRec, n - t * kTileRec);
            // full tiles copy kTileBytes; the last one stops at the end of the last record's
            // 16 useful bytes, rounded up to 16 (never past the record's own 40 bytes)
            const uint32_t bytes = (t + 1 < n_tiles) ? (uint32_t)kTileBytes : ((d + kRecBytes * (nr - 1) + 16u + 15u) & ~15u);
            mbar_wait(bar_empty0 + 8 * stage, phase ^ 1u);
            TileDesc td;
            td.n_rec = nr;
            td.byte_off = d;
            td.frame = f_cur;
            td.last = (t + 1 == n_tiles) ? 1u : 0u;
            td.gw = m_cur.g.gw;
            td.gh = m_cur.g.gh;
            td.y_min = m_cur.g.y_min;
            td.y_max = m_cur.g.y_max;
            desc[stage] = td;
            mbar_arrive_expect_tx(bar_full0 + 8 * stage, bytes);
            bulk_g2s(smem_u32(ring + (size_t)stage * kTileBytes), src + (size_t)t * kTileBytes, bytes,
                     bar_full0 + 8 * stage, policy);
            if (++stage == stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        f_cur = f_next;
        m_cur = m_next;
        f_next = f_next2;
      }
      // terminal descriptor
      mbar_wait(bar_empty0 + 8 * stage, phase ^ 1u);
      desc[stage].frame = kEndFrame;
      desc[stage].n_rec = 0;
      mbar_arrive(bar_full0 + 8 * stage);
    }
  } else {
    // ================================ consumers ===============================================
    const uint32_t ctid = tid - 32;
    const uint32_t cwarp = ctid >> 5;
    const int32_t ithr = a.ithr;
    const int32_t shift = a.shift;
    const bool keep_any = a.keep_none == 0;
    const uint32_t vec_need = a.vec_need;
    uint32_t stage = 0, phase = 0, fseq = 0;
    bool voted = false;
    while (true) {
      mbar_wait(bar_full0 + 8 * stage, phase);
      const TileDesc td = desc[stage];
      if (td.frame == kEndFrame) break;
      const unsigned char* base = ring + (size_t)stage * kTileBytes + td.byte_off;
      const int32_t gw = td.gw;
      const uint32_t live_rows = (uint32_t)(td.y_max - td.y_min);
      for (uint32_t r0 = cwarp * 32; r0 < td.n_rec; r0 += kCons) {
        const uint32_t r = r0 + lane;
        int32_t key = -1;
        if (r < td.n_rec) {
          const unsigned char* p = base + (size_t)r * kRecBytes;
          const uint32_t w1 = *reinterpret_cast<const uint32_t*>(p + 4);  // w | h<<8 | src_x<<16
          const uint2 w23 = *reinterpret_cast<const uint2*>(p + 8);       // src_y | dst_x<<16, dst_y | pad<<16
          const int32_t sx = (int32_t)w1 >> 16;
          const int32_t sy = (int32_t)(int16_t)(w23.x & 0xFFFFu);
          const int32_t tx = (int32_t)w23.x >> 16;
          const int32_t ty = (int32_t)(int16_t)(w23.y & 0xFFFFu);
          const int32_t dx = tx - sx, dy = ty - sy;                        // :246-247
          const int32_t mag = (int32_t)((uint32_t)dx * (uint32_t)dx + (uint32_t)dy * (uint32_t)dy);  // :248
          const int32_t gx = tx >> shift, gy = ty >> shift;                // :255-256
          const bool in = ((uint32_t)gx < (uint32_t)gw) && ((uint32_t)(gy - td.y_min) < live_rows);  // :262
          if (keep_any && mag >= ithr && in) key = gy * gw + gx;           // :251
        }
        // run-length merge of equal neighbouring keys inside the warp
        const int32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
        const bool head = (lane == 0) || (key != prev);
        const uint32_t heads = __ballot_sync(0xffffffffu, head);
        if (head && key >= 0) {
          const uint32_t above = (lane == 31) ? 0u : (heads & (0xFFFFFFFEu << lane));
          const uint32_t next = above ? (uint32_t)(__ffs(above) - 1) : 32u;
          atomicAdd(&cnt[key], next - lane);                               // :265-266
          voted = true;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty0 + 8 * stage);
      if (++stage == stages) {
        stage = 0;
        phase ^= 1u;
      }
      if (td.last) {
        // ---- frame epilogue: Phase 2 (:272-294) -------------------------------------------------
        const bool any = named_bar_or(kBarCons, kCons, voted);
        voted = false;
        if (any || vec_need == 0) {
          const int32_t gh = td.gh;
          const uint32_t wpr = (uint32_t)(gw + 31) >> 5;
          uint32_t* brow = bits + (fseq & 1u) * a.max_bit_words;
          // pass 1: counters → active bit-rows, counters re-zeroed for the next frame
          for (int32_t y = (int32_t)cwarp; y < gh; y += kConsWarps) {
            for (uint32_t w = 0; w < wpr; ++w) {
              const int32_t x = (int32_t)(w * 32 + lane);
              uint32_t c = 0;
              const bool valid = x < gw;
              if (valid) {
                // global counters are voted with L2 atomics: read them past L1 (a plain load could
                // return this SM's stale line from the previous frame)
                c = kGlobalCnt ? __ldcg(&cnt[y * gw + x]) : cnt[y * gw + x];
                cnt[y * gw + x] = 0;
              }
              const uint32_t word = __ballot_sync(0xffffffffu, valid && c >= vec_need);  // :282
              if (lane == 0) brow[(uint32_t)y * wpr + w] = word;
            }
          }
          named_bar_sync(kBarCons, kCons);
          // pass 2: one warp counts centre cells with an active 4-neighbour; the others move on
          if (cwarp == (fseq % kConsWarps)) {
            uint32_t total = 0;
            const uint32_t n_items = live_rows * wpr;
            for (uint32_t i = lane; i < n_items; i += 32) {
              const uint32_t y = (uint32_t)td.y_min + i / wpr;
              const uint32_t w = i % wpr;
              const uint32_t A = brow[y * wpr + w];
              if (A) {
                const uint32_t Lw = w ? brow[y * wpr + w - 1] : 0u;
                const uint32_t Rw = (w + 1 < wpr) ? brow[y * wpr + w + 1] : 0u;
                const uint32_t U = y ? brow[(y - 1) * wpr + w] : 0u;                       // :286 idx-gw
                const uint32_t D = (y + 1 < (uint32_t)gh) ? brow[(y + 1) * wpr + w] : 0u;  // :286 idx+gw
                uint32_t nb = (A << 1) | (Lw >> 31) | (A >> 1) | (Rw << 31) | U | D;  // :284-286
                if (a.adj8) {  // extension (not in the reference): diagonal neighbours too
                  const uint32_t UL = (y && w) ? brow[(y - 1) * wpr + w - 1] : 0u;
                  const uint32_t UR = (y && w + 1 < wpr) ? brow[(y - 1) * wpr + w + 1] : 0u;
                  const uint32_t DL = (y + 1 < (uint32_t)gh && w) ? brow[(y + 1) * wpr + w - 1] : 0u;
                  const uint32_t DR = (y + 1 < (uint32_t)gh && w + 1 < wpr) ? brow[(y + 1) * wpr + w + 1] : 0u;
                  nb |= (U << 1) | (UL >> 31) | (U >> 1) | (UR << 31) | (D << 1) | (DL >> 31) | (D >> 1) | (DR << 31);
                }
                // centre columns are 1 .. gw-2 (:280)
                uint32_t mask = 0xFFFFFFFFu;
                if (w == 0) mask &= ~1u;
                const int32_t hi_bit = gw - 2 - (int32_t)(w * 32);
                if (hi_bit < 31) mask &= (hi_bit < 0) ? 0u : ((2u << hi_bit) - 1u);
                total += (uint32_t)__popc(A & nb & mask);
              }
            }
            total = warp_sum(total);
            if (lane == 0) {
              a.counts[td.frame] = total;
              a.flags[td.frame] = (total >= a.clust_need) ? 1 : 0;   // closed form of :288-289
            }
          }
        } else if (ctid == 0) {
          a.counts[td.frame] = 0;
          a.flags[td.frame] = 0;
        }
        ++fseq;
      }
    }
  }

  // self-resetting work queue: the last CTA to finish re-arms it for the next launch
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const uint32_t done = atomicAdd(a.work + 1, 1u);
    if (done == gridDim.x - 1) {
      a.work[0] = 0;
      a.work[1] = 0;
    }
  }
}

constexpr uint32_t kSmemReserve = 1024;  // per-CTA driver reservation

uint32_t smem_for(uint32_t stages, uint32_t cells_in_smem, uint32_t max_bit_words) {
  return stages * (uint32_t)kTileBytes + stages * (uint32_t)sizeof(TileDesc) + 2 * stages * 8u +
         2 * max_bit_words * 4u + cells_in_smem * 4u + 128u;
}

}  // namespace

bool scan_plan(uint32_t max_cells, uint32_t max_bit_words, uint32_t smem_optin, ScanPlan* plan) {
  const uint32_t sm_total = 228u * 1024u;
  // prefer 2 CTAs/SM with >= 3 stages, else 1 CTA/SM with as deep a ring as fits (<= 6); if the
  // counters do not fit at all they move to global memory
  for (int global_cnt = 0; global_cnt <= 1; ++global_cnt) {
    const uint32_t cells = global_cnt ? 0u : max_cells;
    for (uint32_t ctas = global_cnt ? 1 : 2; ctas >= 1; --ctas) {  // global counters: 1 CTA/SM bounds the scratch
      for (uint32_t st = 6; st >= 2; --st) {
        const uint32_t need = smem_for(st, cells, max_bit_words);
        if (need > smem_optin) continue;
        if (ctas * (need + kSmemReserve) > sm_total) continue;
        if (ctas == 2 && st < 3) continue;
        plan->stages = (ctas == 2 && st > 4) ? 4 : st;
        plan->smem_bytes = smem_for(plan->stages, cells, max_bit_words);
        plan->ctas_per_sm = ctas;
        plan->global_cnt = (uint32_t)global_cnt;
        return true;
      }
    }
  }
  return false;
}

cudaError_t scan_configure(uint32_t smem_optin) {
  cudaError_t e = cudaFuncSetAttribute(ka_scan_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(ka_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin);
}

uint32_t scan_grid(const ScanPlan& plan, int num_sms, uint32_t n_frames) {
  const uint32_t grid = (uint32_t)num_sms * plan.ctas_per_sm;
  return grid > n_frames ? n_frames : grid;
}

cudaError_t scan_launch(const ScanArgs& a, const ScanPlan& plan, int num_sms, cudaStream_t st) {
  if (a.n_frames == 0) return cudaSuccess;
  const uint32_t grid = scan_grid(plan, num_sms, a.n_frames);
  if (plan.global_cnt) ka_scan_kernel<true><<<grid, kThreads, plan.smem_bytes, st>>>(a);
  else ka_scan_kernel<false><<<grid, kThreads, plan.smem_bytes, st>>>(a);
  return cudaGetLastError();
}

}  // namespace mscan
