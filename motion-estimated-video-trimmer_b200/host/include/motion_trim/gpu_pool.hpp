// gpu_pool.hpp — one libmotionscan context per GPU of the box (new; the reference is CPU-only).
#pragma once

#include <atomic>
#include <string>
#include <vector>

#include "motionscan.h"

namespace motion_trim {

class GpuPool {
 public:
  GpuPool() = default;
  ~GpuPool();
  GpuPool(const GpuPool&) = delete;
  GpuPool& operator=(const GpuPool&) = delete;

  // Creates contexts on the first `max_gpus` devices (0 = all) with knobs from the environment.
  // Returns false (and sets error()) when no GPU can be used: there is no CPU fallback.
  bool open(int max_gpus);
  int size() const { return (int)ctx_.size(); }
  mscan_ctx* ctx(int gpu) const { return ctx_[(size_t)gpu]; }
  const mscan_params& params() const { return params_; }
  const std::string& error() const { return error_; }
  // process-unique video ids
  uint32_t next_video_id() { return next_id_.fetch_add(1) + 1; }

 private:
  std::vector<mscan_ctx*> ctx_;
  mscan_params params_{};
  std::string error_;
  std::atomic<uint32_t> next_id_{0};
};

}  // namespace motion_trim
