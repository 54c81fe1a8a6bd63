// exp_hostfeed.cu — one-off measurements behind the host-fed design:
//   (a) pinned H2D cudaMemcpyAsync, 1 GiB                         → PCIe roofline
//   (b) cudaMemcpy2DAsync gathering bytes 6..13 of each 40-B rec  → can the copy engine do the projection?
//   (c) host projection 40 B → 8 B (bytes 6..13), T threads, DRAM-resident source, plain / streaming stores
//   (d) the same from a cache-resident source (what a decode thread sees right after avcodec_receive_frame)
// Build: nvcc -O3 -std=c++17 -arch=sm_100a -Xcompiler -O3,-march=native,-pthread tools/exp_hostfeed.cu -o gpurun_out/exp_hostfeed
#include <cuda_runtime.h>
#include <immintrin.h>
#include <sys/mman.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      std::exit(1);                                                                \
    }                                                                              \
  } while (0)

using clk = std::chrono::steady_clock;
static double secs(clk::time_point a) { return std::chrono::duration<double>(clk::now() - a).count(); }

static void pack_plain(const uint8_t* in, uint64_t n, uint64_t* out) {
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t v;
    std::memcpy(&v, in + 40 * i + 6, 8);
    out[i] = v;
  }
}
static void pack_stream(const uint8_t* in, uint64_t n, uint64_t* out) {
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t v;
    std::memcpy(&v, in + 40 * i + 6, 8);
    _mm_stream_si64(reinterpret_cast<long long*>(out + i), (long long)v);
  }
  _mm_sfence();
}

template <int kDist, int kHint>
static void pack_stream_pf(const uint8_t* in, uint64_t n, uint64_t* out) {
  for (uint64_t i = 0; i < n; ++i) {
    if ((i & 7) == 0) {  // 8 records = 320 B = 5 lines
      const char* q = reinterpret_cast<const char*>(in + 40 * i + kDist);
#pragma unroll
      for (int l = 0; l < 5; ++l) _mm_prefetch(q + 64 * l, (_mm_hint)kHint);
    }
    uint64_t v;
    std::memcpy(&v, in + 40 * i + 6, 8);
    _mm_stream_si64(reinterpret_cast<long long*>(out + i), (long long)v);
  }
  _mm_sfence();
}
// two interleaved streams per thread (more independent miss streams for the L2 prefetcher)
static void pack_stream_2way(const uint8_t* in, uint64_t n, uint64_t* out) {
  const uint64_t h = n / 2;
  for (uint64_t i = 0; i < h; ++i) {
    uint64_t a, b;
    std::memcpy(&a, in + 40 * i + 6, 8);
    std::memcpy(&b, in + 40 * (h + i) + 6, 8);
    _mm_stream_si64(reinterpret_cast<long long*>(out + i), (long long)a);
    _mm_stream_si64(reinterpret_cast<long long*>(out + h + i), (long long)b);
  }
  for (uint64_t i = 2 * h; i < n; ++i) {
    uint64_t v;
    std::memcpy(&v, in + 40 * i + 6, 8);
    out[i] = v;
  }
  _mm_sfence();
}
static void pack_stream_4way(const uint8_t* in, uint64_t n, uint64_t* out) {
  const uint64_t q = n / 4;
  for (uint64_t i = 0; i < q; ++i) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint64_t a;
      std::memcpy(&a, in + 40 * (k * q + i) + 6, 8);
      _mm_stream_si64(reinterpret_cast<long long*>(out + k * q + i), (long long)a);
    }
  }
  for (uint64_t i = 4 * q; i < n; ++i) {
    uint64_t v;
    std::memcpy(&v, in + 40 * i + 6, 8);
    out[i] = v;
  }
  _mm_sfence();
}

template <typename F>
static double run_threads(int T, uint64_t n, F fn) {
  std::vector<std::thread> th;
  auto t0 = clk::now();
  for (int t = 0; t < T; ++t) {
    const uint64_t a = n * t / T, b = n * (t + 1) / T;
    th.emplace_back([=] { fn(a, b); });
  }
  for (auto& x : th) x.join();
  return secs(t0);
}

int main() {
  const uint64_t n = 64ull << 20;  // 64 Mi records = 2.5 GiB native, 512 MiB packed
  uint8_t* h_nat;
  uint64_t* h_pk;
  CK(cudaHostAlloc((void**)&h_nat, n * 40, cudaHostAllocDefault));
  CK(cudaHostAlloc((void**)&h_pk, n * 8, cudaHostAllocDefault));
  {
    const int T = (int)std::thread::hardware_concurrency();
    run_threads(T, n, [&](uint64_t a, uint64_t b) {
      for (uint64_t i = a * 40; i < b * 40; ++i) h_nat[i] = (uint8_t)(i * 2654435761u >> 13);
    });
  }
  uint8_t* d;
  CK(cudaMalloc((void**)&d, n * 40));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float ms;

  // (a)
  double best = 0;
  for (int i = 0; i < 5; ++i) {
    CK(cudaEventRecord(e0, st));
    CK(cudaMemcpyAsync(d, h_nat, 1ull << 30, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    best = std::max(best, (double)(1ull << 30) / ms / 1e6);
  }
  std::printf("(a) pinned H2D 1 GiB: %.1f GB/s = %.2f G native rec/s, %.2f G packed rec/s\n", best, best / 40, best / 8);

  // (b)
  for (uint64_t rows : {1ull << 20, 8ull << 20}) {
    double bb = 0;
    for (int i = 0; i < 3; ++i) {
      CK(cudaEventRecord(e0, st));
      CK(cudaMemcpy2DAsync(d, 8, h_nat + 6, 40, 8, rows, cudaMemcpyHostToDevice, st));
      CK(cudaEventRecord(e1, st));
      CK(cudaStreamSynchronize(st));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      bb = std::max(bb, (double)rows / ms / 1e6);
    }
    std::printf("(b) cudaMemcpy2DAsync width 8 pitch 40, %llu rows: %.3f G rec/s\n", (unsigned long long)rows, bb);
  }
  {
    double bb = 0;
    const uint64_t rows = 8ull << 20;
    for (int i = 0; i < 3; ++i) {
      CK(cudaEventRecord(e0, st));
      CK(cudaMemcpy2DAsync(d, 16, h_nat, 40, 16, rows, cudaMemcpyHostToDevice, st));
      CK(cudaEventRecord(e1, st));
      CK(cudaStreamSynchronize(st));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      bb = std::max(bb, (double)rows / ms / 1e6);
    }
    std::printf("(b') cudaMemcpy2DAsync width 16 pitch 40 (aligned), %llu rows: %.3f G rec/s\n", (unsigned long long)rows, bb);
  }

  // (c) DRAM-resident source
  const int hw = (int)std::thread::hardware_concurrency();
  for (int T : {1, 16}) {
    if (T > hw) break;
    double tp = 1e9, ts = 1e9;
    for (int rep = 0; rep < 3; ++rep) {
      tp = std::min(tp, run_threads(T, n, [&](uint64_t a, uint64_t b) { pack_plain(h_nat + 40 * a, b - a, h_pk + a); }));
      ts = std::min(ts, run_threads(T, n, [&](uint64_t a, uint64_t b) { pack_stream(h_nat + 40 * a, b - a, h_pk + a); }));
    }
    std::printf("(c) host projection from DRAM, %2d threads: plain %.2f G rec/s (%.0f GB/s read), streaming %.2f G rec/s (%.0f GB/s read)\n",
                T, n / tp / 1e9, n * 40 / tp / 1e9, n / ts / 1e9, n * 40 / ts / 1e9);
  }
  // (c2) variants at full thread count and at 8 threads
  for (int T : {8, hw}) {
    auto best_of = [&](auto fn) {
      double t = 1e9;
      for (int rep = 0; rep < 3; ++rep) t = std::min(t, run_threads(T, n, [&](uint64_t a, uint64_t b) { fn(h_nat + 40 * a, b - a, h_pk + a); }));
      return n / t / 1e9;
    };
    std::printf("(c2) %2d threads: base %.2f | pf512/T0 %.2f | pf1024/T0 %.2f | pf2048/T0 %.2f | pf4096/T0 %.2f | pf1024/NTA %.2f | pf2048/T1 %.2f | 2way %.2f | 4way %.2f  G rec/s\n",
                T, best_of(pack_stream), best_of(pack_stream_pf<512, _MM_HINT_T0>), best_of(pack_stream_pf<1024, _MM_HINT_T0>),
                best_of(pack_stream_pf<2048, _MM_HINT_T0>), best_of(pack_stream_pf<4096, _MM_HINT_T0>),
                best_of(pack_stream_pf<1024, _MM_HINT_NTA>), best_of(pack_stream_pf<2048, _MM_HINT_T1>),
                best_of(pack_stream_2way), best_of(pack_stream_4way));
  }
  // (c3) longer prefetch distances, and a transparent-huge-page backed pinned source
  {
    const int T = hw;
    auto best_of = [&](const uint8_t* src, auto fn) {
      double t = 1e9;
      for (int rep = 0; rep < 3; ++rep) t = std::min(t, run_threads(T, n, [&](uint64_t a, uint64_t b) { fn(src + 40 * a, b - a, h_pk + a); }));
      return n / t / 1e9;
    };
    std::printf("(c3) %d threads, cudaHostAlloc source: pf4096 %.2f | pf8192 %.2f | pf16384 %.2f | pf32768 %.2f G rec/s\n", T,
                best_of(h_nat, pack_stream_pf<4096, _MM_HINT_T0>), best_of(h_nat, pack_stream_pf<8192, _MM_HINT_T0>),
                best_of(h_nat, pack_stream_pf<16384, _MM_HINT_T0>), best_of(h_nat, pack_stream_pf<32768, _MM_HINT_T0>));
    const size_t bytes = n * 40;
    void* m = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (m != MAP_FAILED) {
      const int adv = madvise(m, bytes, MADV_HUGEPAGE);
      uint8_t* hp = static_cast<uint8_t*>(m);
      run_threads(T, n, [&](uint64_t a, uint64_t b) { std::memcpy(hp + 40 * a, h_nat + 40 * a, 40 * (b - a)); });
      const cudaError_t reg = cudaHostRegister(hp, bytes, cudaHostRegisterPortable);
      std::printf("(c3) THP source: madvise=%d cudaHostRegister=%s\n", adv, cudaGetErrorString(reg));
      std::printf("(c3) %d threads, THP source: base %.2f | pf4096 %.2f | pf16384 %.2f G rec/s\n", T, best_of(hp, pack_stream),
                  best_of(hp, pack_stream_pf<4096, _MM_HINT_T0>), best_of(hp, pack_stream_pf<16384, _MM_HINT_T0>));
      if (reg == cudaSuccess) {
        double bb = 0;
        for (int i = 0; i < 4; ++i) {
          CK(cudaEventRecord(e0, st));
          CK(cudaMemcpyAsync(d, hp, 1ull << 30, cudaMemcpyHostToDevice, st));
          CK(cudaEventRecord(e1, st));
          CK(cudaStreamSynchronize(st));
          CK(cudaEventElapsedTime(&ms, e0, e1));
          bb = std::max(bb, (double)(1ull << 30) / ms / 1e6);
        }
        std::printf("(c3) H2D from the THP-backed registered buffer: %.1f GB/s\n", bb);
        cudaHostUnregister(hp);
      }
      FILE* f = std::fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r");
      if (f) {
        char line[128] = {0};
        if (std::fgets(line, sizeof line, f)) std::printf("(c3) THP setting: %s", line);
        std::fclose(f);
      }
      munmap(m, bytes);
    }
  }
  // (c') projection overlapped with the DMA of the previous chunk (what the staging ring does)
  {
    const int T = hw;
    const uint64_t chunk = 8ull << 20;  // records per chunk: 64 MiB packed
    const uint64_t n_chunks = n / chunk;
    auto t0 = clk::now();
    for (uint64_t c = 0; c < n_chunks; ++c) {
      run_threads(T, chunk, [&](uint64_t a, uint64_t b) { pack_stream(h_nat + 40 * (c * chunk + a), b - a, h_pk + c * chunk + a); });
      CK(cudaMemcpyAsync(d + 8 * c * chunk, h_pk + c * chunk, 8 * chunk, cudaMemcpyHostToDevice, st));
    }
    CK(cudaStreamSynchronize(st));
    const double t = secs(t0);
    std::printf("(c') projection (%d threads) + async H2D of packed chunks, overlapped: %.2f G rec/s\n", T, n / t / 1e9);
  }

  // (d) cache-resident source: each thread re-projects its own 16 320-record frame (652.8 KB) many times
  for (int T : {1, 16}) {
    if (T > hw) break;
    const uint64_t fr = 16320, reps = 2000;
    double t = run_threads(T, (uint64_t)T, [&](uint64_t a, uint64_t) {
      const uint8_t* src = h_nat + a * fr * 40 * 64;
      uint64_t* dst = h_pk + a * fr * 64;
      for (uint64_t r = 0; r < reps; ++r) pack_stream(src, fr, dst + (r & 31) * fr);
    });
    std::printf("(d) host projection of a cache-hot frame, %2d threads: %.2f G rec/s total\n", T, T * fr * reps / t / 1e9);
  }
  return 0;
}
