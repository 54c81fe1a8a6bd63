// gpu_pool.cpp — libmotionscan contexts, one per GPU.
#include "motion_trim/gpu_pool.hpp"

#include <algorithm>
#include <cstdlib>
#include <string>
#include <thread>

#include "motion_trim/config.hpp"

namespace motion_trim {

GpuPool::~GpuPool() {
  for (mscan_ctx* c : ctx_) mscan_destroy(c);
}

bool GpuPool::open(int max_gpus) {
  if (!Config::validate()) {
    error_ = "unparsable motion_trim environment knob";
    return false;
  }
  params_ = Config::scan_params();
  int n = 0;
  int rc = mscan_device_count(&n);
  if (rc != MSCAN_OK || n <= 0) {
    error_ = "no CUDA device (libmotionscan has no CPU fallback)";
    return false;
  }
  if (max_gpus > 0 && max_gpus < n) n = max_gpus;
  // one creator thread per GPU: the first context on a device costs ≈ 0.5 s (primary context + module load),
  // which would otherwise add up serially on an 8-GPU box
  std::vector<mscan_ctx*> made((size_t)n, nullptr);
  std::vector<int> rcs((size_t)n, MSCAN_OK);
  std::vector<std::thread> creators;
  const uint64_t slab_bytes = (uint64_t)std::max(0, Config::slab_mb()) << 20;
  for (int g = 0; g < n; ++g)
    creators.emplace_back([&, g] {
      rcs[(size_t)g] = mscan_create(g, &params_, 0, slab_bytes, &made[(size_t)g]);
      // the pinned staging ring now, not under the feet of the stream threads (see mscan_reserve_staging)
      if (rcs[(size_t)g] == MSCAN_OK && !std::getenv("MOTION_TRIM_LAZY_STAGING")) mscan_reserve_staging(made[(size_t)g]);
    });
  for (auto& t : creators) t.join();
  for (int g = 0; g < n; ++g) {
    if (rcs[(size_t)g] != MSCAN_OK) {
      error_ = std::string("mscan_create failed on GPU ") + std::to_string(g) + ": " + mscan_status_string(rcs[(size_t)g]);
      for (mscan_ctx* c : made) mscan_destroy(c);
      return false;
    }
  }
  ctx_ = made;
  // MOTION_TRIM_STAGING (experiments; not a reference knob): how mscan_submit moves the workers' records, see
  // mscan_set_staging_mode. Unset = the library's default.
  if (const char* st = std::getenv("MOTION_TRIM_STAGING")) {
    const std::string v(st);
    const int mode = v == "pack" ? MSCAN_STAGING_PACK : v == "native" ? MSCAN_STAGING_NATIVE : v == "elide" ? MSCAN_STAGING_ELIDE
                     : v == "compact" ? MSCAN_STAGING_COMPACT : MSCAN_STAGING_AUTO;
    for (mscan_ctx* c : ctx_) mscan_set_staging_mode(c, mode);
  }
  // (the library's projection pool is one per process, shared by these contexts: nothing to size per GPU)
  return true;
}

}  // namespace motion_trim
