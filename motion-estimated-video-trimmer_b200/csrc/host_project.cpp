// host_project.cpp — the host passes over native records: the projection AVMotionVector → mscan_mv8 used by the staging
// pass of mscan_submit and by mscan_pack_records, the static-elided form and the moving-record compaction (compiled by
// g++, not nvcc: it carries AVX-512 code paths chosen at run time).
//
// Bytes 6..13 of a native 40-byte record are src_x, src_y, dst_x, dst_y — the only fields the path reads
// (reference src/motion_scanner.cpp:243-256) — and they are contiguous: the projection is a pure byte selection,
// one unaligned 8-byte load and one store per record in the portable version. No arithmetic of the path runs here.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "kernels.cuh"

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace mscan {

namespace {

constexpr int kRec = 40;

bool env_on(const char* name) {
  const char* e = getenv(name);
  return e && e[0] && e[0] != '0';
}

// Streaming stores by default: the staging buffer is read next by the DMA engine, not by this core.
// MSCAN_PROJECT_STORES=plain keeps the lines in cache instead (a staging ring small enough to stay LLC-resident
// is then served to the DMA engine from the LLC) — an experiment knob, see DESIGN.md §5.
bool plain_stores() {
  static const bool v = [] {
    const char* e = getenv("MSCAN_PROJECT_STORES");
    return e && !strcmp(e, "plain");
  }();
  return v;
}

template <bool kStream>
void project_scalar(const uint8_t* in, uint64_t n, uint64_t* out) {
  for (uint64_t i = 0; i < n; ++i) {
#if defined(__x86_64__)
    // 8 records = 5 cache lines; software prefetch 4 KB ahead measured +13 % from DRAM on the B200 box's
    // host (tools/exp_hostfeed.cu), free when the source is cache-hot. Prefetches never fault.
    if ((i & 7u) == 0) {
      const char* q = reinterpret_cast<const char*>(in + (size_t)kRec * i + 4096);
      _mm_prefetch(q, _MM_HINT_T0);
      _mm_prefetch(q + 64, _MM_HINT_T0);
      _mm_prefetch(q + 128, _MM_HINT_T0);
      _mm_prefetch(q + 192, _MM_HINT_T0);
      _mm_prefetch(q + 256, _MM_HINT_T0);
    }
#endif
    uint64_t v;
    memcpy(&v, in + (size_t)kRec * i + 6, sizeof v);
#if defined(__x86_64__)
    if (kStream) _mm_stream_si64(reinterpret_cast<long long*>(out + i), (long long)v);
    else out[i] = v;
#else
    out[i] = v;
#endif
  }
}

#if defined(__x86_64__)
// 8 records (320 bytes = 5 unaligned 64-byte loads) → 64 output bytes with three byte permutes (AVX-512 VBMI):
// output byte 8i+b comes from input byte 40i+6+b.
struct Tables {
  alignas(64) uint8_t a[64], b[64], c[64];
  uint64_t mask_b, mask_c;
};
const Tables& tables() {
  static const Tables t = [] {
    Tables x{};
    x.mask_b = x.mask_c = 0;
    for (int o = 0; o < 64; ++o) {
      const int src = 40 * (o >> 3) + 6 + (o & 7);  // byte index into the 320-byte block
      x.a[o] = x.b[o] = x.c[o] = 0;
      if (src < 128) x.a[o] = (uint8_t)src;  // Z0|Z1 as the two sources of vpermi2b
      else if (src < 256) {
        x.b[o] = (uint8_t)(src - 128);       // Z2|Z3
        x.mask_b |= 1ull << o;
      } else {
        x.c[o] = (uint8_t)(src - 256);       // Z4
        x.mask_c |= 1ull << o;
      }
    }
    return x;
  }();
  return t;
}

template <bool kStream>
__attribute__((target("avx512f,avx512bw,avx512vl,avx512vbmi"))) void project_vbmi(const uint8_t* in, uint64_t n, uint64_t* out) {
  // head: until the output is 64-byte aligned (streaming stores want full aligned lines)
  uint64_t head = ((64 - (reinterpret_cast<uintptr_t>(out) & 63)) & 63) / 8;
  if (head > n) head = n;
  project_scalar<kStream>(in, head, out);
  in += (size_t)kRec * head;
  out += head;
  n -= head;
  const Tables& t = tables();
  const __m512i ia = _mm512_load_si512(t.a), ib = _mm512_load_si512(t.b), ic = _mm512_load_si512(t.c);
  const __mmask64 mb = t.mask_b, mc = t.mask_c;
  const uint64_t blocks = n / 8;
  for (uint64_t g = 0; g < blocks; ++g) {
    const uint8_t* p = in + 320 * g;
    _mm_prefetch(reinterpret_cast<const char*>(p + 4096), _MM_HINT_T0);
    _mm_prefetch(reinterpret_cast<const char*>(p + 4096 + 64), _MM_HINT_T0);
    _mm_prefetch(reinterpret_cast<const char*>(p + 4096 + 128), _MM_HINT_T0);
    _mm_prefetch(reinterpret_cast<const char*>(p + 4096 + 192), _MM_HINT_T0);
    _mm_prefetch(reinterpret_cast<const char*>(p + 4096 + 256), _MM_HINT_T0);
    const __m512i z0 = _mm512_loadu_si512(p), z1 = _mm512_loadu_si512(p + 64), z2 = _mm512_loadu_si512(p + 128),
                  z3 = _mm512_loadu_si512(p + 192), z4 = _mm512_loadu_si512(p + 256);
    __m512i v = _mm512_permutex2var_epi8(z0, ia, z1);
    v = _mm512_mask_mov_epi8(v, mb, _mm512_permutex2var_epi8(z2, ib, z3));
    v = _mm512_mask_permutexvar_epi8(v, mc, ic, z4);
    if (kStream) _mm512_stream_si512(reinterpret_cast<__m512i*>(out + 8 * g), v);
    else _mm512_store_si512(reinterpret_cast<__m512i*>(out + 8 * g), v);
  }
  project_scalar<kStream>(in + 320 * blocks, n - 8 * blocks, out + 8 * blocks);
}

bool have_vbmi() {
  static const bool v = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
                        __builtin_cpu_supports("avx512vbmi") && !env_on("MSCAN_NO_AVX512");
  return v;
}
#endif

}  // namespace

void project_records(const uint8_t* in, uint64_t n, uint64_t* out) {
  const bool stream = !plain_stores();
#if defined(__x86_64__)
  if (have_vbmi() && n >= 16) {
    if (stream) project_vbmi<true>(in, n, out);
    else project_vbmi<false>(in, n, out);
    if (stream) _mm_sfence();
    return;
  }
#endif
  if (stream) project_scalar<true>(in, n, out);
  else project_scalar<false>(in, n, out);
#if defined(__x86_64__)
  if (stream) _mm_sfence();
#endif
}

}  // namespace mscan

// ---- mvz: the projected records with static macroblocks elided ---------------------------------------------------
// A lossless transport form of the mscan_mv8 sequence: a record whose src equals its dst — a static
// macroblock, ~90 % of a CCTV stream — is sent as its 4 dst bytes plus one mask bit; a moving record as dst + src.
// The kernel rebuilds exactly the 8 bytes per record the path reads, so nothing of the path's arithmetic runs here:
// the only operation on the data is a byte-equality test used to choose the encoding.
//
// A frame is cut into tiles of kMvzTileRecs records (the last one shorter); a tile is
//     hdr  : one {u32 mask, u32 base} per block of 32 records (bit i of mask: record i of the block is moving;
//            base: index of the block's first entry in `src`), padded to 16 bytes
//     dst  : u32 per record (dst_x | dst_y << 16), padded to 16 bytes
//     src  : u32 per MOVING record (src_x | src_y << 16), in record order, padded to 16 bytes
// and tiles follow each other 16-byte aligned. tile_end16[t] receives the end offset of tile t in 16-byte units,
// relative to `out`.
namespace mscan {

namespace {

inline uint32_t round16(uint32_t b) { return (b + 15u) & ~15u; }

// One tile, portable version.
uint32_t mvz_tile_scalar(const uint8_t* in, uint32_t n, uint8_t* out) {
  const uint32_t nb = (n + 31u) >> 5;
  const uint32_t hdr_bytes = round16(8u * nb), dst_bytes = round16(4u * n);
  uint32_t* hdr = reinterpret_cast<uint32_t*>(out);
  uint32_t* dst = reinterpret_cast<uint32_t*>(out + hdr_bytes);
  uint32_t* src = reinterpret_cast<uint32_t*>(out + hdr_bytes + dst_bytes);
  uint32_t m = 0;
  for (uint32_t b = 0; b < nb; ++b) {
    const uint32_t r0 = b * 32u, r1 = r0 + 32u < n ? r0 + 32u : n;
    uint32_t mask = 0;
    const uint32_t base = m;
    for (uint32_t r = r0; r < r1; ++r) {
      uint64_t v;
      memcpy(&v, in + (size_t)kRec * r + 6, sizeof v);
      const uint32_t s = (uint32_t)v, d = (uint32_t)(v >> 32);
      dst[r] = d;
      src[m] = s;
      const uint32_t mov = s != d;
      m += mov;
      mask |= mov << (r - r0);
    }
    hdr[2 * b] = mask;
    hdr[2 * b + 1] = base;
  }
  for (uint32_t i = 2 * nb; i < hdr_bytes / 4; ++i) hdr[i] = 0;
  for (uint32_t i = n; i < dst_bytes / 4; ++i) dst[i] = 0;
  const uint32_t src_bytes = round16(4u * m);
  for (uint32_t i = m; i < src_bytes / 4; ++i) src[i] = 0;
  return hdr_bytes + dst_bytes + src_bytes;
}

#if defined(__x86_64__)
// One tile with AVX-512 VBMI: 16 records per step — two of the projection's 8-record byte gathers, a dword
// de-interleave into 16 src / 16 dst, one compare, one compress.
__attribute__((target("avx512f,avx512bw,avx512vl,avx512vbmi"))) uint32_t mvz_tile_vbmi(const uint8_t* in, uint32_t n, uint8_t* out) {
  const uint32_t nb = (n + 31u) >> 5;
  const uint32_t hdr_bytes = round16(8u * nb), dst_bytes = round16(4u * n);
  uint32_t* hdr = reinterpret_cast<uint32_t*>(out);
  uint32_t* dst = reinterpret_cast<uint32_t*>(out + hdr_bytes);
  uint32_t* src = reinterpret_cast<uint32_t*>(out + hdr_bytes + dst_bytes);
  const Tables& t = tables();
  const __m512i ia = _mm512_load_si512(t.a), ib = _mm512_load_si512(t.b), ic = _mm512_load_si512(t.c);
  const __mmask64 mb = t.mask_b, mc = t.mask_c;
  const __m512i even = _mm512_set_epi32(30, 28, 26, 24, 22, 20, 18, 16, 14, 12, 10, 8, 6, 4, 2, 0);
  const __m512i odd = _mm512_set_epi32(31, 29, 27, 25, 23, 21, 19, 17, 15, 13, 11, 9, 7, 5, 3, 1);
#define MVZ_GATHER8(out, p)                                                                                          \
  do {                                                                                                               \
    const __m512i z0_ = _mm512_loadu_si512(p), z1_ = _mm512_loadu_si512((p) + 64), z2_ = _mm512_loadu_si512((p) + 128), \
                  z3_ = _mm512_loadu_si512((p) + 192), z4_ = _mm512_loadu_si512((p) + 256);                           \
    __m512i v_ = _mm512_permutex2var_epi8(z0_, ia, z1_);                                                             \
    v_ = _mm512_mask_mov_epi8(v_, mb, _mm512_permutex2var_epi8(z2_, ib, z3_));                                       \
    out = _mm512_mask_permutexvar_epi8(v_, mc, ic, z4_);                                                             \
  } while (0)
  uint32_t m = 0;
  const uint32_t full16 = n / 16u;  // steps of 16 whole records
  uint32_t mask = 0, base = 0;
  for (uint32_t g = 0; g < full16; ++g) {
    const uint8_t* p = in + (size_t)640 * g;
    _mm_prefetch(reinterpret_cast<const char*>(p + 4096), _MM_HINT_T0);
    _mm_prefetch(reinterpret_cast<const char*>(p + 4096 + 128), _MM_HINT_T0);
    _mm_prefetch(reinterpret_cast<const char*>(p + 4096 + 256), _MM_HINT_T0);
    _mm_prefetch(reinterpret_cast<const char*>(p + 4096 + 384), _MM_HINT_T0);
    _mm_prefetch(reinterpret_cast<const char*>(p + 4096 + 512), _MM_HINT_T0);
    __m512i q0, q1;  // 2 x 8 records: src | dst << 32
    MVZ_GATHER8(q0, p);
    MVZ_GATHER8(q1, p + 320);
    const __m512i s = _mm512_permutex2var_epi32(q0, even, q1);  // 16 src dwords
    const __m512i d = _mm512_permutex2var_epi32(q0, odd, q1);   // 16 dst dwords
    const __mmask16 k = _mm512_cmpneq_epi32_mask(s, d);
    _mm512_storeu_si512(dst + 16 * g, d);
    _mm512_storeu_si512(src + m, _mm512_maskz_compress_epi32(k, s));  // (writes 64 bytes; the next store overlaps the slack)
    if ((g & 1u) == 0) {
      base = m;
      mask = (uint32_t)k;
    } else {
      hdr[g - 1] = mask | ((uint32_t)k << 16);  // block b = g / 2: words 2b, 2b + 1
      hdr[g] = base;
    }
    m += (uint32_t)__builtin_popcount((unsigned)k);
  }
  // the remaining (< 16) records, and the header of a block whose second half they are
  uint32_t r = full16 * 16u;
  if (r < n || (full16 & 1u)) {
    if ((full16 & 1u) == 0) {
      base = m;
      mask = 0;
    }
    const uint32_t b = r >> 5;
    for (; r < n; ++r) {
      uint64_t v;
      memcpy(&v, in + (size_t)kRec * r + 6, sizeof v);
      const uint32_t sv = (uint32_t)v, dv = (uint32_t)(v >> 32);
      dst[r] = dv;
      src[m] = sv;
      const uint32_t mov = sv != dv;
      m += mov;
      mask |= mov << (r & 31u);
    }
    hdr[2 * b] = mask;
    hdr[2 * b + 1] = base;
  }
  for (uint32_t i = 2 * nb; i < hdr_bytes / 4; ++i) hdr[i] = 0;
  for (uint32_t i = n; i < dst_bytes / 4; ++i) dst[i] = 0;
  const uint32_t src_bytes = round16(4u * m);
  for (uint32_t i = m; i < src_bytes / 4; ++i) src[i] = 0;
  return hdr_bytes + dst_bytes + src_bytes;
}
#undef MVZ_GATHER8
#endif

}  // namespace

uint64_t mvz_bound(uint64_t n_recs, uint64_t n_frames) {
  const uint64_t tiles = n_recs / kMvzTileRecs + n_frames;  // >= Σ ceil(n_i / kMvzTileRecs)
  // per tile: header, the padding of its three sections, the slack of the 64-byte compress stores
  return tiles * (uint64_t)(round16(8u * (kMvzTileRecs / 32u)) + 48u + 64u) + 8u * n_recs + 64u;
}

uint64_t mvz_encode_frame(const uint8_t* in, uint32_t n, uint8_t* out, uint32_t* tile_end16) {
  uint64_t at = 0;
  uint32_t t = 0;
  for (uint32_t r = 0; r < n; r += kMvzTileRecs, ++t) {
    const uint32_t nt = n - r < kMvzTileRecs ? n - r : kMvzTileRecs;
    uint32_t bytes;
#if defined(__x86_64__)
    if (have_vbmi() && nt >= 32) bytes = mvz_tile_vbmi(in + (size_t)kRec * r, nt, out + at);
    else
#endif
      bytes = mvz_tile_scalar(in + (size_t)kRec * r, nt, out + at);
    at += bytes;
    tile_end16[t] = (uint32_t)(at >> 4);
  }
  return at;
}

// ---- moving-record compaction ----------------------------------------------------------------------------------
// What MSCAN_STAGING_COMPACT puts on the wire: the mscan_mv8 projections of the records whose src differs from their
// dst, in record order — nothing at all for a static macroblock. Valid while the threshold is positive: a record with
// src == dst has mag_sq == 0 (src/motion_scanner.cpp:246-248) and `0 < T²` sends it to `continue` at :251 before it
// can vote, so the frame's flag and cluster count do not depend on it. The only operation on the data is the byte
// equality of the two halves of the projection; every record that could pass :251 reaches the kernel unchanged.
// out needs room for n records; returns the number written.
namespace {

uint64_t compact_scalar(const uint8_t* in, uint64_t n, uint64_t* out) {
  uint64_t m = 0;
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t v;
    memcpy(&v, in + (size_t)kRec * i + 6, sizeof v);
    out[m] = v;
    m += (uint32_t)v != (uint32_t)(v >> 32);
  }
  return m;
}

#if defined(__x86_64__)
// 8 records per step: the projection's byte gather, then "halves differ" as one qword compare and one compress.
// The cost of the pass when the frame is cache-hot is its loads, and a 64-byte load that straddles two lines costs
// two: since 40 k ≡ -a (mod 64) has a solution k < 8 for every 8-byte aligned a, the first k records go through the
// portable loop and every block after them starts on a line (tools/exp_gather.cpp: 0.55 against 0.8 ns/record).
__attribute__((target("avx512f,avx512bw,avx512vl,avx512vbmi"))) uint64_t compact_vbmi(const uint8_t* in, uint64_t n, uint64_t* out) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(in);
  if (a & 7u) return compact_scalar(in, n, out);  // (AVMotionVector holds a uint64: never the case for real side data)
  uint64_t head = ((((64 - (a & 63)) & 63) >> 3) * 5) & 7;  // k with 40 k ≡ -a (mod 64): 5 is its own inverse mod 8
  if (head > n) head = n;
  uint64_t m = compact_scalar(in, head, out);
  in += (size_t)kRec * head;
  n -= head;
  const Tables& t = tables();
  const __m512i ia = _mm512_load_si512(t.a), ib = _mm512_load_si512(t.b), ic = _mm512_load_si512(t.c);
  const __mmask64 mb = t.mask_b, mc = t.mask_c;
  const uint64_t blocks = n / 8;
  // One step: 8 records → mask of the moving ones. Where moving records are rare and clustered (CCTV: quiet areas, a few
  // objects) the compress + store are skipped with a well-predicted branch; where they are scattered (10 % uniformly in
  // the SURVEY §8(d) stream: 57 % of the steps hold one, at random) that branch mispredicts every other step and doubles
  // the cost of the pass (profiles/r03_bench_stream1e9_spec.json → …_spec_after.json: 56 → 28 ms per 60 M records), so stretches of 64 steps that held more
  // than 8 moving records are followed by a stretch without the branch.
#define MSCAN_COMPACT_STEP(BRANCH)                                                                                     \
  {                                                                                                                    \
    const uint8_t* p = in + 320 * g;                                                                                   \
    _mm_prefetch(reinterpret_cast<const char*>(p + 4096), _MM_HINT_T0);                                                \
    _mm_prefetch(reinterpret_cast<const char*>(p + 4096 + 64), _MM_HINT_T0);                                           \
    _mm_prefetch(reinterpret_cast<const char*>(p + 4096 + 128), _MM_HINT_T0);                                          \
    _mm_prefetch(reinterpret_cast<const char*>(p + 4096 + 192), _MM_HINT_T0);                                          \
    _mm_prefetch(reinterpret_cast<const char*>(p + 4096 + 256), _MM_HINT_T0);                                          \
    const __m512i z0 = _mm512_load_si512(p), z1 = _mm512_load_si512(p + 64), z2 = _mm512_load_si512(p + 128),          \
                  z3 = _mm512_load_si512(p + 192), z4 = _mm512_load_si512(p + 256);                                    \
    __m512i v = _mm512_permutex2var_epi8(z0, ia, z1);                                                                  \
    v = _mm512_mask_mov_epi8(v, mb, _mm512_permutex2var_epi8(z2, ib, z3));                                             \
    v = _mm512_mask_permutexvar_epi8(v, mc, ic, z4); /* 8 records: src | dst << 32 */                                  \
    const __mmask8 k = _mm512_cmpneq_epi64_mask(v, _mm512_rol_epi64(v, 32)); /* halves differ ⇔ moving */              \
    if (!(BRANCH) || k) {                                                                                              \
      /* a whole register: at most 8 g records precede it, so it ends inside out[n) */                                 \
      _mm512_storeu_si512(out + m, _mm512_maskz_compress_epi64(k, v));                                                 \
      m += (uint64_t)__builtin_popcount((unsigned)k);                                                                  \
    }                                                                                                                  \
  }
  bool scattered = false;
  for (uint64_t g = 0; g < blocks;) {
    const uint64_t g_end = g + 64 < blocks ? g + 64 : blocks, m0 = m;
    if (scattered) for (; g < g_end; ++g) MSCAN_COMPACT_STEP(false)
    else for (; g < g_end; ++g) MSCAN_COMPACT_STEP(true)
    scattered = m - m0 > 8;
  }
#undef MSCAN_COMPACT_STEP
  return m + compact_scalar(in + 320 * blocks, n - 8 * blocks, out + m);
}
#endif

}  // namespace

uint64_t compact_moving(const uint8_t* in, uint64_t n, uint64_t* out) {
#if defined(__x86_64__)
  if (have_vbmi() && n >= 16) return compact_vbmi(in, n, out);
#endif
  return compact_scalar(in, n, out);
}

#if defined(__x86_64__)
namespace {
__attribute__((target("avx512f"))) uint64_t stream_copy_512(const uint8_t* from, uint8_t* to, uint64_t bytes) {
  uint64_t i = 0;
  for (; i + 64 <= bytes; i += 64) _mm512_stream_si512(reinterpret_cast<__m512i*>(to + i), _mm512_loadu_si512(from + i));
  return i;
}
}  // namespace
#endif

// staging copy of an encoded piece (cache-hot scratch → pinned ring): streaming stores, the buffer is read next by the
// DMA engine, not by this core
void stream_copy(const uint8_t* from, uint8_t* to, uint64_t bytes) {
#if defined(__x86_64__)
  if (!plain_stores() && (reinterpret_cast<uintptr_t>(to) & 7u) == 0) {
    uint64_t i = 0;
    if ((reinterpret_cast<uintptr_t>(to) & 15u) && bytes >= 8) {  // mscan_mv8 offsets are multiples of 8
      long long v;
      memcpy(&v, from, sizeof v);
      _mm_stream_si64(reinterpret_cast<long long*>(to), v);
      i = 8;
    }
    for (; i + 16 <= bytes && ((reinterpret_cast<uintptr_t>(to + i)) & 63u); i += 16)  // up to the first full line
      _mm_stream_si128(reinterpret_cast<__m128i*>(to + i), _mm_loadu_si128(reinterpret_cast<const __m128i*>(from + i)));
    if (have_vbmi()) i += stream_copy_512(from + i, to + i, bytes - i);
    for (; i + 16 <= bytes; i += 16)
      _mm_stream_si128(reinterpret_cast<__m128i*>(to + i), _mm_loadu_si128(reinterpret_cast<const __m128i*>(from + i)));
    if (i < bytes) memcpy(to + i, from + i, bytes - i);
    _mm_sfence();
    return;
  }
#endif
  memcpy(to, from, bytes);
}

}  // namespace mscan
