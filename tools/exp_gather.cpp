// Experiment: two ways to pull bytes 6..13 out of eight 40-byte records with AVX-512, cache-hot, 1..T threads.
//   A  three byte permutes (VBMI vpermi2b / vpermb) over five 64-byte loads at the record base
//   B  five merging masked qword loads at base + 6 and one qword permute (AVX-512F)
// Both feed the compaction (compare halves, compress) and the plain projection (store). Build: g++ -O2 -pthread.
#include <immintrin.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <thread>
#include <vector>
#include <atomic>
#include <sched.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct Tables { alignas(64) uint8_t a[64], b[64], c[64]; uint64_t mb, mc; };
static Tables T;
static void init_tables() {
  T.mb = T.mc = 0;
  for (int o = 0; o < 64; ++o) {
    const int src = 40 * (o >> 3) + 6 + (o & 7);
    T.a[o] = T.b[o] = T.c[o] = 0;
    if (src < 128) T.a[o] = (uint8_t)src;
    else if (src < 256) { T.b[o] = (uint8_t)(src - 128); T.mb |= 1ull << o; }
    else { T.c[o] = (uint8_t)(src - 256); T.mc |= 1ull << o; }
  }
}
#define TGT_A __attribute__((target("avx512f,avx512bw,avx512vl,avx512vbmi")))
#define TGT_B __attribute__((target("avx512f")))
TGT_A static inline __m512i gatherA(const uint8_t* p, __m512i ia, __m512i ib, __m512i ic) {
  const __m512i z0 = _mm512_loadu_si512(p), z1 = _mm512_loadu_si512(p + 64), z2 = _mm512_loadu_si512(p + 128), z3 = _mm512_loadu_si512(p + 192), z4 = _mm512_loadu_si512(p + 256);
  __m512i v = _mm512_permutex2var_epi8(z0, ia, z1);
  v = _mm512_mask_mov_epi8(v, T.mb, _mm512_permutex2var_epi8(z2, ib, z3));
  return _mm512_mask_permutexvar_epi8(v, T.mc, ic, z4);
}
TGT_B static inline __m512i gatherB(const uint8_t* p, __m512i order) {
  const uint8_t* q = p + 6;
  __m512i w = _mm512_loadu_si512(q);
  w = _mm512_mask_loadu_epi64(w, 0x84, q + 64);
  w = _mm512_mask_loadu_epi64(w, 0x10, q + 128);
  w = _mm512_mask_loadu_epi64(w, 0x42, q + 192);
  w = _mm512_mask_loadu_epi64(w, 0x08, q + 256);
  return _mm512_permutexvar_epi64(order, w);
}
TGT_A static uint64_t compactA(const uint8_t* in, uint64_t n, uint64_t* out) {
  const __m512i ia = _mm512_load_si512(T.a), ib = _mm512_load_si512(T.b), ic = _mm512_load_si512(T.c);
  uint64_t m = 0;
  for (uint64_t g = 0; g < n / 8; ++g) {
    const __m512i v = gatherA(in + 320 * g, ia, ib, ic);
    const __mmask8 k = _mm512_cmpneq_epi64_mask(v, _mm512_rol_epi64(v, 32));
    if (k) { _mm512_storeu_si512(out + m, _mm512_maskz_compress_epi64(k, v)); m += __builtin_popcount((unsigned)k); }
  }
  return m;
}
TGT_B static uint64_t compactB(const uint8_t* in, uint64_t n, uint64_t* out) {
  const __m512i order = _mm512_set_epi64(3, 6, 1, 4, 7, 2, 5, 0);
  uint64_t m = 0;
  for (uint64_t g = 0; g < n / 8; ++g) {
    const __m512i v = gatherB(in + 320 * g, order);
    const __mmask8 k = _mm512_cmpneq_epi64_mask(v, _mm512_rol_epi64(v, 32));
    if (k) { _mm512_storeu_si512(out + m, _mm512_maskz_compress_epi64(k, v)); m += __builtin_popcount((unsigned)k); }
  }
  return m;
}
// A with the input peeled to a 64-byte boundary first (40 k = -a mod 64 has a solution k < 8 for every 8-byte aligned a)
static uint64_t compact_scalar(const uint8_t* in, uint64_t n, uint64_t* out) {
  uint64_t m = 0;
  for (uint64_t i = 0; i < n; ++i) { uint64_t v; memcpy(&v, in + 40 * i + 6, 8); out[m] = v; m += (uint32_t)v != (uint32_t)(v >> 32); }
  return m;
}
TGT_A static uint64_t compactC(const uint8_t* in, uint64_t n, uint64_t* out) {
  const __m512i ia = _mm512_load_si512(T.a), ib = _mm512_load_si512(T.b), ic = _mm512_load_si512(T.c);
  uint64_t head = (((64 - (reinterpret_cast<uintptr_t>(in) & 63)) & 63) / 8 * 5) & 7;
  if ((reinterpret_cast<uintptr_t>(in) & 7) || head > n) head = n < 8 ? n : 0;
  uint64_t m = compact_scalar(in, head, out);
  in += 40 * head; n -= head;
  const bool al = (reinterpret_cast<uintptr_t>(in) & 63) == 0;
  uint64_t g = 0;
  if (al) for (; g < n / 8; ++g) {
    const uint8_t* p = in + 320 * g;
    const __m512i z0 = _mm512_load_si512(p), z1 = _mm512_load_si512(p + 64), z2 = _mm512_load_si512(p + 128), z3 = _mm512_load_si512(p + 192), z4 = _mm512_load_si512(p + 256);
    __m512i v = _mm512_permutex2var_epi8(z0, ia, z1);
    v = _mm512_mask_mov_epi8(v, T.mb, _mm512_permutex2var_epi8(z2, ib, z3));
    v = _mm512_mask_permutexvar_epi8(v, T.mc, ic, z4);
    const __mmask8 k = _mm512_cmpneq_epi64_mask(v, _mm512_rol_epi64(v, 32));
    if (k) { _mm512_storeu_si512(out + m, _mm512_maskz_compress_epi64(k, v)); m += __builtin_popcount((unsigned)k); }
  }
  return m + compact_scalar(in + 320 * g, n - 8 * g, out + m);
}
TGT_A static uint64_t compactD(const uint8_t* in, uint64_t n, uint64_t* out) {
  const __m512i ia = _mm512_load_si512(T.a), ib = _mm512_load_si512(T.b), ic = _mm512_load_si512(T.c);
  uint64_t head = (((64 - (reinterpret_cast<uintptr_t>(in) & 63)) & 63) / 8 * 5) & 7;
  if ((reinterpret_cast<uintptr_t>(in) & 7) || head > n) head = n < 8 ? n : 0;
  uint64_t m = compact_scalar(in, head, out);
  in += 40 * head; n -= head;
  const bool al = (reinterpret_cast<uintptr_t>(in) & 63) == 0;
  uint64_t g = 0;
  if (al) for (; g < n / 8; ++g) {
    const uint8_t* p = in + 320 * g;
    const __m512i z0 = _mm512_load_si512(p), z1 = _mm512_load_si512(p + 64), z2 = _mm512_load_si512(p + 128), z3 = _mm512_load_si512(p + 192), z4 = _mm512_load_si512(p + 256);
    __m512i v = _mm512_permutex2var_epi8(z0, ia, z1);
    v = _mm512_mask_mov_epi8(v, T.mb, _mm512_permutex2var_epi8(z2, ib, z3));
    v = _mm512_mask_permutexvar_epi8(v, T.mc, ic, z4);
    const __mmask8 k = _mm512_cmpneq_epi64_mask(v, _mm512_rol_epi64(v, 32));
    _mm512_storeu_si512(out + m, _mm512_maskz_compress_epi64(k, v)); m += __builtin_popcount((unsigned)k);
  }
  return m + compact_scalar(in + 320 * g, n - 8 * g, out + m);
}
TGT_A static uint64_t projectA(const uint8_t* in, uint64_t n, uint64_t* out) {
  const __m512i ia = _mm512_load_si512(T.a), ib = _mm512_load_si512(T.b), ic = _mm512_load_si512(T.c);
  for (uint64_t g = 0; g < n / 8; ++g) _mm512_storeu_si512(out + 8 * g, gatherA(in + 320 * g, ia, ib, ic));
  return n;
}
TGT_B static uint64_t projectB(const uint8_t* in, uint64_t n, uint64_t* out) {
  const __m512i order = _mm512_set_epi64(3, 6, 1, 4, 7, 2, 5, 0);
  for (uint64_t g = 0; g < n / 8; ++g) _mm512_storeu_si512(out + 8 * g, gatherB(in + 320 * g, order));
  return n;
}
typedef uint64_t (*Fn)(const uint8_t*, uint64_t, uint64_t*);
int main(int argc, char** argv) {
  init_tables();
  const size_t n = argc > 1 ? atol(argv[1]) : 10000;
  const int maxT = argc > 2 ? atoi(argv[2]) : (int)std::thread::hardware_concurrency();
  const int uniform_pct = argc > 3 ? atoi(argv[3]) : 0;  // 0: one moving record in 40, regularly; else: that percentage, scattered
  struct V { const char* name; Fn f; } vs[] = {{"compact A (vbmi permutes)", compactA}, {"compact B (masked loads) ", compactB}, {"compact C (A, input peeled)", compactC}, {"compact D (C, branchless)   ", compactD}, {"project A (vbmi permutes)", projectA}, {"project B (masked loads) ", projectB}};
  for (int threads : {1, maxT / 2, maxT}) {
    if (threads < 1) continue;
    for (int shift : {0, 24}) {
      for (auto& v : vs) {
        double best = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
          std::atomic<int> ready{0};
          std::atomic<bool> go{false};
          std::vector<double> dt(threads);
          std::vector<std::thread> th;
          for (int t = 0; t < threads; ++t)
            th.emplace_back([&, t] {
              cpu_set_t m; CPU_ZERO(&m); CPU_SET(t, &m); sched_setaffinity(0, sizeof m, &m);
              uint8_t* buf = (uint8_t*)aligned_alloc(64, 40 * n + 256);
              for (size_t i = 0; i < 40 * n + 256; ++i) buf[i] = (uint8_t)(i * 7 + t);
              uint8_t* in = buf + shift;
              uint64_t lcg = 88172645463325252ull + t;
              for (size_t r = 0; r < n; ++r) {
                lcg ^= lcg << 13; lcg ^= lcg >> 7; lcg ^= lcg << 17;
                const bool moving = uniform_pct ? (int)(lcg % 100) < uniform_pct : (r % 40) == 0;
                if (!moving) memcpy(in + 40 * r + 6, in + 40 * r + 10, 4);
              }
              std::vector<uint64_t> out(n + 16);
              volatile uint64_t sink = 0;
              for (int i = 0; i < 50; ++i) sink += v.f(in, n, out.data());
              ready.fetch_add(1);
              while (!go.load()) {}
              const int reps = (int)(200000000 / (40 * n)) + 1;
              const double t0 = now();
              for (int i = 0; i < reps; ++i) sink += v.f(in, n, out.data());
              dt[t] = (now() - t0) / reps;
              free(buf);
            });
          while (ready.load() < threads) {}
          go.store(true);
          for (auto& x : th) x.join();
          double worst = 0;
          for (double d : dt) worst = d > worst ? d : worst;
          best = worst < best ? worst : best;
        }
        printf("threads %2d shift %2d  %s  %.3f ns/rec/thread  (%.1f G rec/s aggregate)\n", threads, shift, v.name, best / n * 1e9, threads * n / best / 1e9);
      }
    }
  }
}
