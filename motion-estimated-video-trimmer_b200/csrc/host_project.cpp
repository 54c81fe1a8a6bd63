// host_project.cpp — the host projection AVMotionVector → mscan_mv8 used by the staging pass of mscan_submit and
// by mscan_pack_records (compiled by g++, not nvcc: it carries AVX-512 code paths chosen at run time).
//
// Bytes 6..13 of a native 40-byte record are src_x, src_y, dst_x, dst_y — the only fields the path reads
// (reference src/motion_scanner.cpp:243-256) — and they are contiguous: the projection is a pure byte selection,
// one unaligned 8-byte load and one store per record in the portable version. No arithmetic of the path runs here.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

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
