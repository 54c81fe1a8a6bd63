// config.hpp — environment knobs of the host mirror. Same accessor names, env-var names and code
// defaults as the reference's motion_trim::Config (include/motion_trim/config.hpp:56-177), so code
// written against the reference reads the same here. Parsing is strtod/strtol based: an unparsable
// value falls back to the default and is reported once (the reference's std::stod would terminate).
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>

namespace motion_trim {
namespace Config {

namespace detail {
inline double env_f64(const char* name, double dflt) {
  const char* s = std::getenv(name);
  if (!s) return dflt;
  char* end = nullptr;
  const double v = std::strtod(s, &end);
  if (end == s) {
    std::fprintf(stderr, "[WARN] %s='%s' is not a number; using %g\n", name, s, dflt);
    return dflt;
  }
  return v;
}
inline int env_i32(const char* name, int dflt) {
  const char* s = std::getenv(name);
  if (!s) return dflt;
  char* end = nullptr;
  const long v = std::strtol(s, &end, 10);
  if (end == s) {
    std::fprintf(stderr, "[WARN] %s='%s' is not an integer; using %d\n", name, s, dflt);
    return dflt;
  }
  return (int)v;
}
}  // namespace detail

#define MT_KNOB(type, fn, expr) \
  inline type fn() {            \
    static const type v = expr; \
    return v;                   \
  }

// motion scan (consumed inside libmotionscan through mscan_params_from_env; mirrored for logging)
MT_KNOB(double, mv_threshold_sq, detail::env_f64("MV_THRESHOLD_SQ", 16.0))
MT_KNOB(int, block_size, detail::env_i32("BLOCK_SIZE", 16))
MT_KNOB(int, block_shift, detail::env_i32("BLOCK_SHIFT", 4))
MT_KNOB(uint8_t, vectors_needed, (uint8_t)detail::env_i32("VECTORS_NEEDED", 2))
MT_KNOB(int, clusters_needed, detail::env_i32("CLUSTERS_NEEDED", 2))
MT_KNOB(float, vertical_mask, (float)detail::env_f64("VERTICAL_MASK", 0.05))
MT_KNOB(double, max_gap_sec, detail::env_f64("MAX_GAP_SEC", 5.0))
MT_KNOB(double, padding_sec, detail::env_f64("PADDING_SEC", 0.5))
MT_KNOB(double, min_savings_pct, detail::env_f64("MIN_SAVINGS_PCT", 5.0))
// host-side scheduling
MT_KNOB(double, chunk_duration_sec, detail::env_f64("CHUNK_DURATION_SEC", 30.0))
MT_KNOB(double, target_fps, detail::env_f64("TARGET_FPS", 0.0))
MT_KNOB(int, parallel_streams, detail::env_i32("PARALLEL_STREAMS", 0))      // here: streams PER GPU (0 = 2)
MT_KNOB(int, threads_per_stream, detail::env_i32("THREADS_PER_STREAM", 0))  // chunk workers per stream (0 = auto)
MT_KNOB(bool, watch_mode, detail::env_i32("WATCH_MODE", 0) != 0)
// new: how many GPUs of the box to use (0 = all)
MT_KNOB(int, gpus, detail::env_i32("MOTION_TRIM_GPUS", 0))
// new: size of each of the library's three record slabs (+ pinned staging), MiB (0 = library default)
MT_KNOB(int, slab_mb, detail::env_i32("MOTION_TRIM_SLAB_MB", 0))
// new: single-file mode only — split the one video over this many GPUs (0/1 = one GPU); SURVEY §8(f) N4
MT_KNOB(int, split_gpus, detail::env_i32("MOTION_TRIM_SPLIT_GPUS", 0))
// new: watch mode ends after this many idle seconds (0 = never, like the reference)
MT_KNOB(double, watch_idle_exit_sec, detail::env_f64("MOTION_TRIM_WATCH_IDLE_EXIT_SEC", 0.0))

#undef MT_KNOB

}  // namespace Config
}  // namespace motion_trim
