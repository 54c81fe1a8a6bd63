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
#include <cstdlib>

#include "common.cuh"
#include "kernels.cuh"

namespace mscan {

namespace {

constexpr int kTileRec = 512;                      // native records per ring stage
constexpr int kTileBytes = kTileRec * kRecBytes;   // 20480, multiple of lcm(16,40)=80
// Projected records: half-size stages. K-A<packed> is bound by per-CTA work (issue slots, frame barriers), not by
// bytes in flight, so what pays is more resident CTAs per SM; smaller stages are what lets them fit.
constexpr int kTileBytesPacked = 10240;
// consumer warps per CTA: 8 when two CTAs share an SM, 16 when only one fits (4K grids: the sweep in
// tools/ka_sweep.py shows one CTA of 8 consumer warps cannot keep up with HBM)
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
  uint32_t tile0;  // mvz: index of the frame's first tile in the tile directory
};

__device__ __forceinline__ FrameMeta load_meta(const ScanArgs& a, uint32_t f) {
  FrameMeta m;
  m.o0 = 0;
  m.o1 = 0;
  m.g = DevGeom{0, 0, 0, 0};
  m.tile0 = 0;
  if (f < a.n_frames) {
    if (a.frame_tile0) m.tile0 = __ldg(a.frame_tile0 + f);
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
// kCnt16: counters are 16-bit halves of shared-memory words (half the footprint ⇒ a deeper ring for 4K
// grids). A vote is still one 32-bit atomicAdd on the containing word; a guard stops adding once a
// half reaches 0x7FFF, and since at most kCons lanes × 64 merged votes (two records per lane in the packed layout)
// can pass the guard concurrently a half never carries into its neighbour (16 warps x 64 per trip). Counts >= 0x7FFF are reported as "many": exact for the
// active test because VECTORS_NEEDED <= 255.
// kPacked: the slab holds 8-byte projections of the records (bytes 6..13 of AVMotionVector: src_x, src_y,
// dst_x, dst_y — mscan_mv8) instead of the native 40-byte layout. Same ring, same stage size; a stage then
// carries 2560 records, and since an 8-byte record never straddles a 16-byte boundary the tiles are cut
// on the aligned byte stream rather than on record counts.
// kLayout: kLayoutNative / kLayoutMv8 / kLayoutMvz. mvz (host_project.cpp): the projected records with static
// macroblocks elided — a tile is {mask, base} per block of 32 records, one u32 dst per record, one u32 src per MOVING
// record. A block without a moving record costs one broadcast load and a branch; everything that survives is rebuilt to
// the exact (src, dst) pair and goes through the same int32 test.
template <bool kGlobalCnt, int kConsWarps, bool kCnt16, int kLayout>
__global__ void __launch_bounds__(kConsWarps * 32 + 32, kConsWarps == 8 ? 2 : 1) ka_scan_kernel(const __grid_constant__ ScanArgs a) {
  static_assert(!(kGlobalCnt && kCnt16), "global counters are always 32-bit");
  constexpr bool kPacked = kLayout != (int)kLayoutNative;  // 8-byte-or-less records: half-size stages, sleeping waits
  constexpr bool kMv8 = kLayout == (int)kLayoutMv8;
  constexpr bool kMvz = kLayout == (int)kLayoutMvz;
  constexpr uint32_t kStride = kPacked ? kPackedBytes : kRecBytes;
  constexpr uint32_t kTile = (kPacked && kConsWarps == 8) ? kTileBytesPacked : kTileBytes;  // bytes per ring stage (tile_bytes_for)
  constexpr int kCons = kConsWarps * 32;
  constexpr int kThreads = kCons + 32;
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t stages = a.stages;
  unsigned char* ring = smem;
  TileDesc* desc = reinterpret_cast<TileDesc*>(ring + (size_t)stages * kTile);
  uint64_t* bars = reinterpret_cast<uint64_t*>(desc + stages);  // full[stages], empty[stages]
  uint32_t* bits = reinterpret_cast<uint32_t*>(bars + 2 * stages);
  // vote marks: one byte per 32-cell word of the FLAT grid (cells [32k, 32k+32)) = "received a vote this frame".
  // ceil(max_cells / 32) <= max_bit_words (a row-aligned word never holds more than 32 cells), so the marks fit
  // max_bit_words bytes; the region is padded to a multiple of 16.
  unsigned char* wordv = reinterpret_cast<unsigned char*>(bits + 2 * a.max_bit_words);
  uint32_t* cnt = kGlobalCnt ? a.cnt_scratch + (size_t)blockIdx.x * a.max_cells
                             : reinterpret_cast<uint32_t*>(wordv + ((a.max_bit_words + 15u) & ~15u));

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
    for (uint32_t i = tid; i < (kCnt16 ? (a.max_cells + 1) / 2 : a.max_cells); i += kThreads) cnt[i] = 0;
  for (uint32_t i = tid; i < a.max_bit_words; i += kThreads) wordv[i] = 0;
  for (uint32_t i = tid; i < 2 * a.max_bit_words; i += kThreads) bits[i] = 0;
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
          if (kMvz) {
            const uint32_t n_tiles = (n + kMvzTileRecs - 1) / kMvzTileRecs;
            const uint32_t* dir = a.tile_dir + m_cur.tile0;
            uint32_t e0 = __ldg(dir), e1 = __ldg(dir + 1);
            for (uint32_t t = 0; t < n_tiles; ++t) {
              const uint32_t e2 = (t + 2 <= n_tiles) ? __ldg(dir + t + 2) : 0u;  // next tile's end, loaded a trip ahead
              const uint32_t bytes = (e1 - e0) << 4;
              mbar_wait<kPacked>(bar_empty0 + 8 * stage, phase ^ 1u);
              TileDesc td;
              td.n_rec = min(kMvzTileRecs, n - t * kMvzTileRecs);
              td.byte_off = 0;
              td.frame = f_cur;
              td.last = (t + 1 == n_tiles) ? 1u : 0u;
              td.gw = m_cur.g.gw;
              td.gh = m_cur.g.gh;
              td.y_min = m_cur.g.y_min;
              td.y_max = m_cur.g.y_max;
              desc[stage] = td;
              mbar_arrive_expect_tx(bar_full0 + 8 * stage, bytes);
              bulk_g2s(smem_u32(ring + (size_t)stage * kTile), a.recs + ((size_t)e0 << 4), bytes, bar_full0 + 8 * stage, policy);
              if (++stage == stages) {
                stage = 0;
                phase ^= 1u;
              }
              e0 = e1;
              e1 = e2;
            }
            f_cur = f_next;
            m_cur = m_next;
            f_next = f_next2;
            continue;
          }
          const uint64_t byte0 = m_cur.o0 * (uint64_t)kStride;
          const uint32_t d = (uint32_t)(byte0 & 15u);
          const unsigned char* src = a.recs + (byte0 - d);
          // native: tile t holds records [512t, 512t+512); packed: tile t holds the records that lie in
          // bytes [20480t, 20480t+20480) of the 16-byte aligned stream starting at src
          const uint32_t n_tiles = kMv8 ? (uint32_t)((d + (uint64_t)kPackedBytes * n + kTile - 1) / kTile)
                                           : (n + kTileRec - 1) / kTileRec;
          for (uint32_t t = 0; t < n_tiles; ++t) {
            uint32_t nr, bytes, boff;
            if (kMv8) {
              const uint32_t r_lo = t ? (uint32_t)(((uint64_t)t * kTile - d) / kPackedBytes) : 0u;
              const uint32_t r_hi = (uint32_t)min((uint64_t)n, ((uint64_t)(t + 1) * kTile - d) / kPackedBytes);
              nr = r_hi - r_lo;
              boff = t ? 0u : d;
              bytes = (t + 1 < n_tiles) ? kTile : ((boff + kPackedBytes * nr + 15u) & ~15u);
            } else {
              nr = min((uint32_t)kTileRec, n - t * kTileRec);
              boff = d;
              // full tiles copy kTileBytes; the last one stops at the end of the last record's
              // 16 useful bytes, rounded up to 16 (never past the record's own 40 bytes)
              bytes = (t + 1 < n_tiles) ? (uint32_t)kTileBytes : ((d + kRecBytes * (nr - 1) + 16u + 15u) & ~15u);
            }
            mbar_wait<kPacked>(bar_empty0 + 8 * stage, phase ^ 1u);
            TileDesc td;
            td.n_rec = nr;
            td.byte_off = boff;
            td.frame = f_cur;
            td.last = (t + 1 == n_tiles) ? 1u : 0u;
            td.gw = m_cur.g.gw;
            td.gh = m_cur.g.gh;
            td.y_min = m_cur.g.y_min;
            td.y_max = m_cur.g.y_max;
            desc[stage] = td;
            mbar_arrive_expect_tx(bar_full0 + 8 * stage, bytes);
            bulk_g2s(smem_u32(ring + (size_t)stage * kTile), src + (size_t)t * kTile, bytes,
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
      mbar_wait<kPacked>(bar_empty0 + 8 * stage, phase ^ 1u);
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
      mbar_wait<kPacked>(bar_full0 + 8 * stage, phase);
      const TileDesc td = desc[stage];
      if (td.frame == kEndFrame) break;
      const int32_t gw = td.gw;
      const uint32_t live_rows = (uint32_t)(td.y_max - td.y_min);
      // one vote of `amount` for cell `key` (:265-266), with the carry guard of the 16-bit layout; the 32-cell word of
      // the flat grid that holds the cell is marked so that the epilogue only visits words that received votes
      auto vote = [&](int32_t key, int32_t, uint32_t amount) {
        if (kCnt16) {
          uint32_t* w = &cnt[(uint32_t)key >> 1];
          const uint32_t sh = ((uint32_t)key & 1u) * 16u;
          if (((*reinterpret_cast<volatile uint32_t*>(w) >> sh) & 0xFFFFu) < 0x7FFFu) atomicAdd(w, amount << sh);
        } else {
          atomicAdd(&cnt[key], amount);
        }
        wordv[key >> 5] = 1;  // (the key's 32-cell word of the flat grid: what the epilogue has to look at)
      };
      // exact int32 test of :246-262 on one record; returns the cell key or -1, and the cell's row
      auto key_of = [&](int32_t sx, int32_t sy, int32_t tx, int32_t ty, int32_t* gy_out) -> int32_t {
        const int32_t dx = tx - sx, dy = ty - sy;                                                    // :246-247
        const int32_t mag = (int32_t)((uint32_t)dx * (uint32_t)dx + (uint32_t)dy * (uint32_t)dy);    // :248
        const int32_t gx = tx >> shift, gy = ty >> shift;                                            // :255-256
        const bool in = ((uint32_t)gx < (uint32_t)gw) && ((uint32_t)(gy - td.y_min) < live_rows);    // :262
        *gy_out = gy;
        return (mag >= ithr && in) ? gy * gw + gx : -1;                                              // :251
      };
      // run-length merge of equal neighbouring keys inside the warp (one record per lane), then the votes
      auto merge_vote = [&](int32_t key, int32_t gy) {
        const int32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
        const bool head = (lane == 0) || (key != prev);
        const uint32_t heads = __ballot_sync(0xffffffffu, head);
        if (head && key >= 0) {
          const uint32_t above = (lane == 31) ? 0u : (heads & (0xFFFFFFFEu << lane));
          const uint32_t next = above ? (uint32_t)(__ffs(above) - 1) : 32u;
          vote(key, gy, next - lane);
          voted = true;
        }
      };
      if (kMvz) {
        const unsigned char* sb = ring + (size_t)stage * kTile;
        const uint32_t n_rec = td.n_rec, nb = (n_rec + 31u) >> 5;
        const uint32_t hdr_bytes = (8u * nb + 15u) & ~15u, dst_bytes = (4u * n_rec + 15u) & ~15u;
        const uint2* hdr = reinterpret_cast<const uint2*>(sb);
        const uint32_t* dstA = reinterpret_cast<const uint32_t*>(sb + hdr_bytes);
        const uint32_t* srcA = reinterpret_cast<const uint32_t*>(sb + hdr_bytes + dst_bytes);
        const bool skip_static = ithr > 0;  // a record with src == dst has mag_sq == 0: it cannot vote while T² > 0
        for (uint32_t b = cwarp; keep_any && b < nb; b += kConsWarps) {
          const uint2 h = hdr[b];  // {mask of moving records, index of the block's first src}
          if (skip_static && h.x == 0) continue;
          const uint32_t r = b * 32 + lane;
          const bool moving = (h.x >> lane) & 1u;
          int32_t key = -1, gy = 0;
          if (r < n_rec && (moving || !skip_static)) {
            const uint32_t d = dstA[r];
            const uint32_t sv = moving ? srcA[h.y + (uint32_t)__popc(h.x & ((1u << lane) - 1u))] : d;
            key = key_of((int32_t)(int16_t)(sv & 0xFFFFu), (int32_t)sv >> 16, (int32_t)(int16_t)(d & 0xFFFFu), (int32_t)d >> 16, &gy);
          }
          if (!__any_sync(0xffffffffu, key >= 0)) continue;
          merge_vote(key, gy);
        }
      } else if (kMv8) {
        // Projected records, two per lane: one LDS.128 per 16-byte slot of the stage. A record whose src equals its dst
        // (w.x == w.y: a static macroblock — about 90 % of a CCTV stream) has mag_sq == 0 and cannot vote while the
        // threshold is positive, so one compare per record decides, and a warp trip without a moving record costs the
        // loads, the compares and one vote.any. Survivors go through the exact test above. Slots past the tile's end
        // are loaded too (the address stays inside this CTA's shared memory) and masked by the record-index test.
        const uint4* sp = reinterpret_cast<const uint4*>(ring + (size_t)stage * kTile);
        const uint32_t o = td.byte_off >> 3;  // 0 or 1: the stage's first 8 bytes precede the tile's first record
        const uint32_t n_rec = td.n_rec;
        const uint32_t n_slots = (td.byte_off + (uint32_t)kPackedBytes * n_rec + 15u) >> 4;
        const bool skip_static = ithr > 0;
        // records 2j - o (first half of slot j) and 2j + 1 - o (second half); unsigned compares reject the halves that
        // are not records of this tile
        auto slot = [&](uint32_t j, const uint4& w) {
          const uint32_t ra = 2u * j - o;
          const bool ma = (ra < n_rec) && (!skip_static || w.x != w.y);
          const bool mb = (ra + 1u < n_rec) && (!skip_static || w.z != w.w);
          if (!__any_sync(0xffffffffu, ma || mb)) return;
          int32_t gya = 0, gyb = 0;
          const int32_t ka = ma ? key_of((int32_t)(int16_t)(w.x & 0xFFFFu), (int32_t)w.x >> 16, (int32_t)(int16_t)(w.y & 0xFFFFu), (int32_t)w.y >> 16, &gya) : -1;
          const int32_t kb = mb ? key_of((int32_t)(int16_t)(w.z & 0xFFFFu), (int32_t)w.z >> 16, (int32_t)(int16_t)(w.w & 0xFFFFu), (int32_t)w.w >> 16, &gyb) : -1;
          if (!__any_sync(0xffffffffu, (ka & kb) >= 0)) return;  // (ka & kb) >= 0  ⇔  ka >= 0 || kb >= 0
          // lane-level run-length merge: a lane whose two records vote for one cell (or only one of them votes) is one
          // item; lanes with two different cells vote twice on their own
          const bool mixed = ka >= 0 && kb >= 0 && ka != kb;
          const int32_t K = mixed ? (-2 - (int32_t)lane) : (ka >= 0 ? ka : kb);
          const uint32_t two = __ballot_sync(0xffffffffu, !mixed && ka >= 0 && kb >= 0);
          const int32_t prev = __shfl_up_sync(0xffffffffu, K, 1);
          const bool head = (lane == 0) || (K != prev);
          const uint32_t heads = __ballot_sync(0xffffffffu, head);
          if (mixed) {
            vote(ka, gya, 1u);
            vote(kb, gyb, 1u);
            voted = true;
          } else if (head && K >= 0) {
            const uint32_t above = (lane == 31) ? 0u : (heads & (0xFFFFFFFEu << lane));
            const uint32_t next = above ? (uint32_t)(__ffs(above) - 1) : 32u;
            const uint32_t run = (next == 32u ? 0xFFFFFFFFu : ((1u << next) - 1u)) & (0xFFFFFFFFu << lane);  // lanes [lane, next)
            vote(K, ka >= 0 ? gya : gyb, (next - lane) + (uint32_t)__popc(two & run));
            voted = true;
          }
        };
        if (keep_any) {
          uint32_t j = cwarp * 32 + lane;
          for (; j - lane + kCons < n_slots; j += 2 * kCons) {  // two slots per lane per trip: both loads in flight together
            const uint4 w0 = sp[j], w1 = sp[j + kCons];
            if (skip_static) {
              const uint32_t ra0 = 2u * j - o, ra1 = ra0 + 2u * kCons;
              const bool m = ((ra0 < n_rec) && w0.x != w0.y) || ((ra0 + 1u < n_rec) && w0.z != w0.w) ||
                             ((ra1 < n_rec) && w1.x != w1.y) || ((ra1 + 1u < n_rec) && w1.z != w1.w);
              if (!__any_sync(0xffffffffu, m)) continue;
            }
            slot(j, w0);
            slot(j + kCons, w1);
          }
          if (j - lane < n_slots) slot(j, sp[j]);
        }
      } else {
        const unsigned char* base = ring + (size_t)stage * kTile + td.byte_off;
        for (uint32_t r0 = cwarp * 32; r0 < td.n_rec; r0 += kCons) {
          const uint32_t r = r0 + lane;
          int32_t key = -1, gy = 0;
          if (r < td.n_rec) {
            const unsigned char* p = base + (size_t)r * kRecBytes;
            const uint32_t w1 = *reinterpret_cast<const uint32_t*>(p + 4);  // w | h<<8 | src_x<<16
            const uint2 w23 = *reinterpret_cast<const uint2*>(p + 8);       // src_y | dst_x<<16, dst_y | pad<<16
            const int32_t k = key_of((int32_t)w1 >> 16, (int32_t)(int16_t)(w23.x & 0xFFFFu), (int32_t)w23.x >> 16,
                                     (int32_t)(int16_t)(w23.y & 0xFFFFu), &gy);
            if (keep_any) key = k;
          }
          // nothing to vote in this warp trip (static macroblocks: the common CCTV case) — skip the merge
          if (!__any_sync(0xffffffffu, key >= 0)) continue;
          merge_vote(key, gy);
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
        const uint32_t wpr = (uint32_t)(gw + 31) >> 5;
        if ((any || vec_need == 0) && wpr != 0) {
          const int32_t gh = td.gh;
          const bool all_words = vec_need == 0;  // every cell is active, voted or not
          uint32_t* brow = bits + (fseq & 1u) * a.max_bit_words;
          // pass 1: counters → active bit-rows (:282), counters re-zeroed for the next frame (:229). Votes marked the
          // 32-cell words of the FLAT grid (cells [32k, 32k+32)) they fell into; only those are read. A flat word k
          // spans at most two grid rows; its cells are turned into bits of the row-aligned bit-row words with atomicOr
          // after the bit-rows have been cleared.
          // (the bit-rows of this parity are clean: zeroed at kernel start, and again by the pass-2 warp of the frame
          // that last used them, which joins this frame's barriers only after it has finished)
          const uint32_t n_bw = (uint32_t)gh * wpr;
          const uint32_t n_cells = (uint32_t)gh * (uint32_t)gw;
          const uint32_t n_fw = (n_cells + 31u) >> 5;
          for (uint32_t k0 = cwarp * 32; k0 < n_fw; k0 += kCons) {
            const uint32_t k = k0 + lane;
            const bool mine = k < n_fw && (all_words || wordv[k]);
            if (mine) wordv[k] = 0;
            uint32_t todo = __ballot_sync(0xffffffffu, mine);
            while (todo) {  // the whole warp reads one marked word: 32 cells
              const uint32_t kk = k0 + (uint32_t)(__ffs(todo) - 1);
              todo &= todo - 1;
              const uint32_t cell = kk * 32 + lane;
              uint32_t c = 0;
              const bool valid = cell < n_cells;
              if (valid) {
                if (kCnt16) {
                  uint16_t* h = reinterpret_cast<uint16_t*>(cnt) + cell;
                  c = *h;
                  *h = 0;
                } else {
                  // global counters are voted with L2 atomics: read them past L1 (a plain load could
                  // return this SM's stale line from the previous frame)
                  c = kGlobalCnt ? __ldcg(&cnt[cell]) : cnt[cell];
                  cnt[cell] = 0;
                }
              }
              // (variants measured and dropped, profiles/r03_ka_epilogue_variants.log: one lane converting the word
              // segment-wise — 3 to 7 % slower; marks per row-aligned word with direct ballot stores — 5 to 12 % slower,
              // the extra index arithmetic on every vote costs more than an atomicOr for each of the few active cells)
              if (valid && c >= vec_need) {  // :282
                const uint32_t y = cell / (uint32_t)gw, x = cell - y * (uint32_t)gw;
                atomicOr(&brow[y * wpr + (x >> 5)], 1u << (x & 31u));
              }
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
            __syncwarp();
            for (uint32_t i = lane; i < n_bw; i += 32) brow[i] = 0;  // leave this parity's bit-rows clean for frame n + 2
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

// bytes per ring stage: half-size stages for the projected layouts with 8-warp CTAs (more CTAs per SM); 16-warp CTAs keep
// the full stage — 16 warps x 64 records per trip would leave a 1 280-record stage two thirds empty on its second trip
constexpr uint32_t tile_bytes_for(bool packed, uint32_t cons_warps) {
  return (packed && cons_warps == 8) ? (uint32_t)kTileBytesPacked : (uint32_t)kTileBytes;
}

uint32_t smem_for(uint32_t stages, uint32_t counter_bytes, uint32_t max_bit_words, uint32_t tile_bytes) {
  return stages * tile_bytes + stages * (uint32_t)sizeof(TileDesc) + 2 * stages * 8u +
         2 * max_bit_words * 4u + ((max_bit_words + 15u) & ~15u) + ((counter_bytes + 15u) & ~15u) + 128u;  // 2 bit-row buffers + vote marks
}

// Tuning overrides for experiments (tools/ka_sweep.py): MSCAN_KA_CTAS, MSCAN_KA_STAGES, MSCAN_KA_WARPS, MSCAN_KA_CNT16.
uint32_t env_u32(const char* name) {
  const char* s = std::getenv(name);
  return s ? (uint32_t)std::strtoul(s, nullptr, 10) : 0u;
}

// largest ring depth in [lo, hi] that fits `ctas` CTAs per SM; 0 if none
uint32_t fit_stages(uint32_t ctas, uint32_t counter_bytes, uint32_t bit_words, uint32_t smem_optin, uint32_t lo, uint32_t hi, uint32_t tile_bytes) {
  const uint32_t sm_total = 228u * 1024u;
  for (uint32_t st = hi; st >= lo; --st) {
    const uint32_t need = smem_for(st, counter_bytes, bit_words, tile_bytes);
    if (need <= smem_optin && ctas * (need + kSmemReserve) <= sm_total) return st;
  }
  return 0;
}

}  // namespace

bool scan_plan(uint32_t max_cells, uint32_t max_bit_words, uint32_t smem_optin, bool packed, ScanPlan* plan) {
  const uint32_t b32 = max_cells * 4u, b16 = ((max_cells + 1u) / 2u) * 4u;
  const uint32_t t8 = tile_bytes_for(packed, 8), t16 = tile_bytes_for(packed, 16);
  auto set = [&](uint32_t ctas, uint32_t st, uint32_t cnt16, uint32_t global_cnt, uint32_t warps) {
    plan->stages = st;
    plan->ctas_per_sm = ctas;
    plan->cnt16 = cnt16;
    plan->global_cnt = global_cnt;
    plan->cons_warps = warps;
    plan->smem_bytes = smem_for(st, global_cnt ? 0u : (cnt16 ? b16 : b32), max_bit_words, tile_bytes_for(packed, warps));
    return true;
  };
  const uint32_t want_st = env_u32("MSCAN_KA_STAGES"), want_ctas = env_u32("MSCAN_KA_CTAS");
  if (want_st >= 2 && want_st <= 16 && want_ctas >= 1 && want_ctas <= 6) {
    const uint32_t c16 = env_u32("MSCAN_KA_CNT16") ? 1u : 0u;
    uint32_t w = env_u32("MSCAN_KA_WARPS");
    if (w != 8 && w != 16) w = want_ctas == 1 ? 16 : 8;
    if (fit_stages(want_ctas, c16 ? b16 : b32, max_bit_words, smem_optin, want_st, want_st, tile_bytes_for(packed, w)))
      return set(want_ctas, want_st, c16, 0, w);
  }
  uint32_t st;
  if (packed) {
    // Projected records carry 5x fewer bytes per vote: K-A<packed> is bound by per-CTA work (issue slots, the frame
    // barriers), not by HBM, so the plan buys resident CTAs with ring depth — half-size stages (10 KB) are what lets a
    // fourth CTA fit at 1080p (sweep profiles/r03_ka_sweep_packed.log: 4 CTAs x 8 warps x 3 stages 556 G rec/s,
    // 5 x 8 x 2 517, 3 x 8 x 4 482; 16-warp CTAs are slower)
    if ((st = fit_stages(4, b16, max_bit_words, smem_optin, 3, 3, t8))) return set(4, st, 1, 0, 8);
    if ((st = fit_stages(3, b16, max_bit_words, smem_optin, 2, 4, t8))) return set(3, st, 1, 0, 8);
    // 4K grids: two 16-warp CTAs with two full-size stages each (356 G rec/s on the dense field; four half-size stages 293)
    if ((st = fit_stages(2, b16, max_bit_words, smem_optin, 2, 2, t16))) return set(2, st, 1, 0, 16);
  }
  // 1-4: two CTAs per SM; a 4-stage ring (160 KB in flight per SM) measures ~1.3 % faster than 3 stages
  // (tools/ka_sweep.py), so 16-bit counters are preferred when they are what makes the 4th stage fit
  if ((st = fit_stages(2, b32, max_bit_words, smem_optin, 4, 4, t8))) return set(2, st, 0, 0, 8);
  if ((st = fit_stages(2, b16, max_bit_words, smem_optin, 4, 4, t8))) return set(2, st, 1, 0, 8);
  if ((st = fit_stages(2, b32, max_bit_words, smem_optin, 3, 3, t8))) return set(2, st, 0, 0, 8);
  if ((st = fit_stages(2, b16, max_bit_words, smem_optin, 3, 3, t8))) return set(2, st, 1, 0, 8);
  // 5-7: one CTA per SM with 16 consumer warps and as deep a ring as fits (4K: u16 counters give 7 stages)
  if ((st = fit_stages(1, b32, max_bit_words, smem_optin, 6, 8, t16))) return set(1, st, 0, 0, 16);
  if ((st = fit_stages(1, b16, max_bit_words, smem_optin, 2, 8, t16))) return set(1, st, 1, 0, 16);
  if ((st = fit_stages(1, b32, max_bit_words, smem_optin, 2, 8, t16))) return set(1, st, 0, 0, 16);
  // 8: counters in global memory (8K and larger)
  if ((st = fit_stages(1, 0, max_bit_words, smem_optin, 2, 8, t16))) return set(1, st, 0, 1, 16);
  return false;
}

bool scan_plan_for(const DevGeom* geoms, uint32_t n_geoms, uint32_t smem_optin, bool packed, ScanPlan* plan) {
  uint32_t cells = 0, words = 0;
  for (uint32_t i = 0; i < n_geoms; ++i) {
    cells = max(cells, (uint32_t)geoms[i].gw * (uint32_t)geoms[i].gh);
    words = max(words, (uint32_t)geoms[i].gh * (((uint32_t)geoms[i].gw + 31u) >> 5));
  }
  const bool ok = scan_plan(cells, words, smem_optin, packed, plan);
  plan->cluster = 0;
  plan->cells = cells;
  plan->bit_words = words;
  plan->full_cells = cells;
  plan->full_bit_words = words;
  if (ok && !plan->global_cnt) return true;
  if (env_u32("MSCAN_KA_NO_CLUSTER")) return ok;
  // The grid does not fit one CTA: distribute it over a cluster's shared memory in row bands (ka_scan_cluster.cu).
  // Ring depth decides the streaming rate (≥ 7 stages ≈ 140 KB in flight per SM, as for grids that fit), so the
  // cluster size is the smallest one whose bands leave room for 7 stages, else the one with the deepest ring
  // (8K: 2 CTAs → 4 stages 6.8 TB/s, 4 CTAs → 8 stages 7.2 TB/s; profiles/r02_ka_cluster_sweep.log).
  const uint32_t want_c = env_u32("MSCAN_KA_CLUSTER"), want_stages = env_u32("MSCAN_KA_CLUSTER_STAGES");  // experiments
  uint32_t best_c = 0, best_st = 0, best_cells = 0, best_words = 0;
  for (uint32_t C = 2; C <= 16; C *= 2) {  // 16 is a non-portable cluster size (opt-in, B200 supports it)
    if (want_c && C != want_c) continue;
    uint32_t band_cells = 0, band_words = 0;
    for (uint32_t i = 0; i < n_geoms; ++i) {
      const uint32_t rpr = ((uint32_t)geoms[i].gh + C - 1) / C;
      band_cells = max(band_cells, rpr * (uint32_t)geoms[i].gw);
      band_words = max(band_words, rpr * (((uint32_t)geoms[i].gw + 31u) >> 5));
    }
    uint32_t st = want_stages ? want_stages : 8;
    while (st >= 2 && scan_cluster_smem(st, band_cells, band_words) > smem_optin) --st;
    if (st < 2) continue;
    if (st > best_st) {
      best_c = C;
      best_st = st;
      best_cells = band_cells;
      best_words = band_words;
    }
    if (st >= (packed ? 3u : 7u)) break;  // projected records are issue-bound, not HBM-bound: the smallest cluster wins
  }
  if (best_c) {
    plan->stages = best_st;
    plan->smem_bytes = scan_cluster_smem(best_st, best_cells, best_words);
    plan->ctas_per_sm = 1;
    plan->global_cnt = 0;
    plan->cons_warps = 16;
    plan->cnt16 = 1;
    plan->cluster = best_c;
    plan->cells = best_cells;
    plan->bit_words = best_words;
    return true;
  }
  return ok;
}

namespace {
template <int kPacked>
cudaError_t configure_all(int v) {
  const auto attr = cudaFuncAttributeMaxDynamicSharedMemorySize;
  cudaError_t e = cudaFuncSetAttribute(ka_scan_kernel<false, 8, false, kPacked>, attr, v);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(ka_scan_kernel<false, 16, false, kPacked>, attr, v);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(ka_scan_kernel<false, 8, true, kPacked>, attr, v);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(ka_scan_kernel<false, 16, true, kPacked>, attr, v);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(ka_scan_kernel<true, 8, false, kPacked>, attr, v);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(ka_scan_kernel<true, 16, false, kPacked>, attr, v);
  return e;
}

template <int kPacked>
void launch_one(const ScanArgs& a, const ScanPlan& plan, uint32_t grid, cudaStream_t st) {
  const bool wide = plan.cons_warps == 16;
  const uint32_t threads = plan.cons_warps * 32 + 32;
  if (plan.global_cnt) {
    if (wide) ka_scan_kernel<true, 16, false, kPacked><<<grid, threads, plan.smem_bytes, st>>>(a);
    else ka_scan_kernel<true, 8, false, kPacked><<<grid, threads, plan.smem_bytes, st>>>(a);
  } else if (plan.cnt16) {
    if (wide) ka_scan_kernel<false, 16, true, kPacked><<<grid, threads, plan.smem_bytes, st>>>(a);
    else ka_scan_kernel<false, 8, true, kPacked><<<grid, threads, plan.smem_bytes, st>>>(a);
  } else {
    if (wide) ka_scan_kernel<false, 16, false, kPacked><<<grid, threads, plan.smem_bytes, st>>>(a);
    else ka_scan_kernel<false, 8, false, kPacked><<<grid, threads, plan.smem_bytes, st>>>(a);
  }
}
}  // namespace

cudaError_t scan_configure(uint32_t smem_optin) {
  cudaError_t e = configure_all<(int)kLayoutNative>((int)smem_optin);
  if (e == cudaSuccess) e = configure_all<(int)kLayoutMv8>((int)smem_optin);
  if (e == cudaSuccess) e = configure_all<(int)kLayoutMvz>((int)smem_optin);
  if (e == cudaSuccess) e = scan_cluster_configure(smem_optin);
  return e;
}

uint32_t scan_grid(const ScanPlan& plan, int num_sms, uint32_t n_frames) {
  const uint32_t grid = (uint32_t)num_sms * plan.ctas_per_sm;
  return grid > n_frames ? n_frames : grid;
}

cudaError_t scan_launch(const ScanArgs& a, const ScanPlan& plan, int num_sms, cudaStream_t st) {
  if (a.n_frames == 0) return cudaSuccess;
  if (plan.cluster) return scan_cluster_launch(a, plan, num_sms, st);
  const uint32_t grid = scan_grid(plan, num_sms, a.n_frames);
  if (a.packed == kLayoutMvz) launch_one<(int)kLayoutMvz>(a, plan, grid, st);
  else if (a.packed == kLayoutMv8) launch_one<(int)kLayoutMv8>(a, plan, grid, st);
  else launch_one<(int)kLayoutNative>(a, plan, grid, st);
  return cudaGetLastError();
}

}  // namespace mscan
