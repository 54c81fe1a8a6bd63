// config.hpp — environment knobs of the host mirror. Same accessor names, env-var names and code
// defaults as the reference's motion_trim::Config (include/motion_trim/config.hpp:56-177), so code
// written against the reference reads the same here.
// The scan knobs (MV_THRESHOLD_SQ … MIN_SAVINGS_PCT) have ONE parser: the library's mscan_params_from_env
// (strtod / strtol / strtof for the float32 VERTICAL_MASK, exactly what the kernels will use); the accessors
// below read that struct. An unparsable value is fatal, as in the reference, whose std::stod / std::stoi
// (config.hpp:28-53) throw and terminate: Config::validate() reports it and main() exits.
#pragma once

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "motionscan.h"

namespace motion_trim {
namespace Config {

namespace detail {
inline std::atomic<int>& bad_knobs() {
  static std::atomic<int> n{0};
  return n;
}
inline double env_f64(const char* name, double dflt) {
  const char* s = std::getenv(name);
  if (!s) return dflt;
  char* end = nullptr;
  const double v = std::strtod(s, &end);
  if (end == s) {
    std::fprintf(stderr, "[ERROR] %s='%s' is not a number\n", name, s);
    bad_knobs()++;
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
    std::fprintf(stderr, "[ERROR] %s='%s' is not an integer\n", name, s);
    bad_knobs()++;
    return dflt;
  }
  return (int)v;
}
struct ScanKnobs {
  mscan_params p;
  bool ok;
};
inline const ScanKnobs& scan_knobs() {
  static const ScanKnobs k = [] {
    ScanKnobs x{};
    x.ok = mscan_params_from_env(&x.p) == MSCAN_OK;
    if (!x.ok) {
      std::fprintf(stderr, "[ERROR] unparsable motion-scan knob in the environment (MV_THRESHOLD_SQ, BLOCK_SIZE, BLOCK_SHIFT, "
                           "VECTORS_NEEDED, CLUSTERS_NEEDED, VERTICAL_MASK, CLUSTER_ADJACENCY, MAX_GAP_SEC, PADDING_SEC, MIN_SAVINGS_PCT)\n");
      mscan_params_default(&x.p);
    }
    return x;
  }();
  return k;
}
}  // namespace detail

// the scan knobs exactly as the library parsed them (what mscan_create receives)
inline const mscan_params& scan_params() { return detail::scan_knobs().p; }

#define MT_KNOB(type, fn, expr) \
  inline type fn() {            \
    static const type v = expr; \
    return v;                   \
  }

// motion scan: read through the library's parser (see the header comment)
inline double mv_threshold_sq() { return scan_params().mv_threshold_sq; }
inline int block_size() { return scan_params().block_size; }
inline int block_shift() { return scan_params().block_shift; }
inline uint8_t vectors_needed() { return (uint8_t)scan_params().vectors_needed; }  // config.hpp:75 static_cast<uint8_t>
inline int clusters_needed() { return scan_params().clusters_needed; }
inline float vertical_mask() { return scan_params().vertical_mask; }
inline double max_gap_sec() { return scan_params().max_gap_sec; }
inline double padding_sec() { return scan_params().padding_sec; }
inline double min_savings_pct() { return scan_params().min_savings_pct; }
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

// Parses every knob now; false (after naming the offenders on stderr) if any value is not a number. main() refuses
// to start then, like the reference, whose parse throws out of main.
inline bool validate() {
  const bool scan_ok = detail::scan_knobs().ok;
  chunk_duration_sec();
  target_fps();
  parallel_streams();
  threads_per_stream();
  watch_mode();
  gpus();
  slab_mb();
  split_gpus();
  watch_idle_exit_sec();
  return scan_ok && detail::bad_knobs().load() == 0;
}

}  // namespace Config
}  // namespace motion_trim
