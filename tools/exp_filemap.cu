// exp_filemap.cu — how fast can records that live in a (tmpfs/page-cache) file reach the projection pass or the DMA engine?
//   mmap MAP_PRIVATE [+MAP_POPULATE] → 16-thread projection;  → cudaHostRegister(ReadOnly) → H2D;  read() into pinned memory.
#include <cuda_runtime.h>
#include <emmintrin.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

using clk = std::chrono::steady_clock;
static double since(clk::time_point a) { return std::chrono::duration<double>(clk::now() - a).count(); }

static void project(const uint8_t* in, uint64_t n, uint64_t* out) {
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t v;
    std::memcpy(&v, in + 40 * i + 6, 8);
    _mm_stream_si64((long long*)(out + i), (long long)v);
  }
  _mm_sfence();
}
template <typename F>
static double par(int T, uint64_t n, F fn) {
  std::vector<std::thread> th;
  auto t0 = clk::now();
  for (int t = 0; t < T; ++t) th.emplace_back([=] { fn(n * t / T, n * (t + 1) / T); });
  for (auto& x : th) x.join();
  return since(t0);
}

int main(int argc, char** argv) {
  const char* path = argc > 1 ? argv[1] : "/dev/shm/exp_filemap.bin";
  const uint64_t n = 18ull << 20;  // ~one 60 s clip: 18 Mi records = 755 MB
  const size_t bytes = n * 40;
  const int T = (int)std::thread::hardware_concurrency();
  {
    std::vector<uint8_t> buf(64 << 20);
    for (size_t i = 0; i < buf.size(); ++i) buf[i] = (uint8_t)(i * 2654435761u >> 11);
    int fd = open(path, O_CREAT | O_TRUNC | O_WRONLY, 0644);
    for (size_t off = 0; off < bytes; off += buf.size()) if (write(fd, buf.data(), std::min(buf.size(), bytes - off)) < 0) return 1;
    close(fd);
  }
  uint64_t* out;
  cudaHostAlloc((void**)&out, n * 8, cudaHostAllocDefault);
  uint8_t* dev;
  cudaMalloc((void**)&dev, bytes);
  cudaStream_t st;
  cudaStreamCreate(&st);
  for (int populate = 1; populate >= 0; --populate) {
    int fd = open(path, O_RDONLY);
    auto t0 = clk::now();
    uint8_t* m = (uint8_t*)mmap(nullptr, bytes, PROT_READ, MAP_PRIVATE | (populate ? MAP_POPULATE : 0), fd, 0);
    const double t_map = since(t0);
    const double t1 = par(T, n, [&](uint64_t a, uint64_t b) { project(m + 40 * a, b - a, out + a); });
    const double t2 = par(T, n, [&](uint64_t a, uint64_t b) { project(m + 40 * a, b - a, out + a); });
    std::printf("mmap populate=%d: map %.3f s | %d-thread projection 1st pass %.3f s (%.2f G rec/s) 2nd pass %.3f s (%.2f G rec/s)\n", populate,
                t_map, T, t1, n / t1 / 1e9, t2, n / t2 / 1e9);
    t0 = clk::now();
    cudaError_t e = cudaHostRegister(m, bytes, cudaHostRegisterPortable | cudaHostRegisterReadOnly);
    const double t_reg = since(t0);
    if (e == cudaSuccess) {
      for (int rep = 0; rep < 2; ++rep) {
        t0 = clk::now();
        cudaMemcpyAsync(dev, m, bytes, cudaMemcpyHostToDevice, st);
        cudaStreamSynchronize(st);
        const double t_dma = since(t0);
        std::printf("   cudaHostRegister(ReadOnly) %.3f s; H2D of the mapping pass %d: %.3f s = %.1f GB/s\n", t_reg, rep, t_dma, bytes / t_dma / 1e9);
      }
      t0 = clk::now();
      cudaHostUnregister(m);
      std::printf("   cudaHostUnregister %.3f s\n", since(t0));
    } else {
      std::printf("   cudaHostRegister failed: %s (%.3f s)\n", cudaGetErrorString(e), t_reg);
      cudaGetLastError();
    }
    t0 = clk::now();
    munmap(m, bytes);
    close(fd);
    std::printf("   munmap %.3f s\n", since(t0));
  }
  {  // read() into pinned memory, T threads with pread
    uint8_t* pin;
    cudaHostAlloc((void**)&pin, bytes, cudaHostAllocDefault);
    int fd = open(path, O_RDONLY);
    for (int rep = 0; rep < 2; ++rep) {
      const double t = par(T, bytes >> 20, [&](uint64_t a, uint64_t b) {
        for (uint64_t mb = a; mb < b; ++mb) if (pread(fd, pin + (mb << 20), 1 << 20, (off_t)(mb << 20)) < 0) return;
      });
      std::printf("pread into pinned memory, %d threads, pass %d: %.3f s = %.1f GB/s\n", T, rep, t, bytes / t / 1e9);
    }
    const double t1 = par(T, n, [&](uint64_t a, uint64_t b) { project(pin + 40 * a, b - a, out + a); });
    std::printf("projection from that pinned copy: %.3f s (%.2f G rec/s)\n", t1, n / t1 / 1e9);
    close(fd);
  }
  unlink(path);
  return 0;
}
