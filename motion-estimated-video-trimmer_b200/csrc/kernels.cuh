// kernels.cuh — host-callable launchers of the motion-scan kernels (internal to libmotionscan.so).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/motionscan.h"
#include "../../include/mvgen_core.h"

namespace mscan {

constexpr int kRecBytes = 40;    // sizeof(AVMotionVector), src/motion_scanner.cpp:226
constexpr int kPackedBytes = 8;  // sizeof(mscan_mv8): bytes 6..13 of the native record

// Geometry as the kernel consumes it: rows [y_min, y_max) are live (clamped to the grid).
struct DevGeom {
  int32_t gw, gh, y_min, y_max;
};

// ---- K-A: vote scatter + cluster count + activity flag (check_frame, motion_scanner.cpp:217-295)
struct ScanArgs {
  const uint8_t* recs;         // 16-byte aligned base of 40-byte records (8-byte when `packed`)
  const uint64_t* rec_off;     // [n_frames+1] record indices
  const uint32_t* frame_geom;  // [n_frames] index into geoms, or nullptr → 0
  const DevGeom* geoms;        // device table
  uint8_t* flags;              // [n_frames]
  uint32_t* counts;            // [n_frames] full cluster counts
  uint32_t* work;              // work[0] = next frame, work[1] = finished CTAs (self-resetting)
  uint32_t n_frames;
  int32_t ithr;                // keep iff mag_sq >= ithr   (== !((double)mag_sq < T²))
  int32_t keep_none;           // T² above INT32_MAX / +inf: nothing ever votes
  int32_t shift;
  uint32_t vec_need;           // uint8 VECTORS_NEEDED
  uint32_t clust_need;         // max(1, CLUSTERS_NEEDED)
  uint32_t stages;             // ring depth
  uint32_t max_cells;          // vote counters per CTA (shared memory, or a slice of cnt_scratch)
  uint32_t max_bit_words;      // words per bit-row buffer
  uint32_t adj8;               // extension: 8-neighbour clusters (reference is 4-neighbour only)
  uint32_t packed;             // record layout: 0 native AVMotionVector, 1 mscan_mv8 projections, 2 mvz (host_project.cpp)
  uint32_t* cnt_scratch;       // zeroed global scratch, grid × max_cells, only when plan.global_cnt
  // mvz only: recs = base of the segment's tiles; tile_dir[k] = start of tile k in 16-byte units from recs (tile k ends
  // at tile_dir[k + 1]); frame_tile0[f] = index of frame f's first tile. rec_off still gives the record counts.
  const uint32_t* tile_dir;
  const uint32_t* frame_tile0;
};
constexpr uint32_t kLayoutNative = 0, kLayoutMv8 = 1, kLayoutMvz = 2;

struct ScanPlan {
  uint32_t stages;
  uint32_t smem_bytes;
  uint32_t ctas_per_sm;
  uint32_t global_cnt;  // counters do not fit shared memory: use the global scratch
  uint32_t cons_warps;  // 8 (two CTAs per SM) or 16 (one CTA per SM)
  uint32_t cnt16;       // 16-bit shared-memory counters (half the footprint, guarded against carry)
  uint32_t cluster;     // 0, or the cluster size (2/4/8) of ka_scan_cluster.cu: the grid lives in DSMEM row bands
  uint32_t cells;       // → ScanArgs.max_cells (whole grid, or one band under a cluster plan)
  uint32_t bit_words;   // → ScanArgs.max_bit_words
  uint32_t full_cells, full_bit_words;  // the whole largest grid (what a non-cluster plan would need)
};

// Chooses ring depth / occupancy for the largest geometry and the record layout; false if it cannot fit.
bool scan_plan(uint32_t max_cells, uint32_t max_bit_words, uint32_t smem_optin, bool packed, ScanPlan* plan);
// The plan for a set of geometries: single-CTA if the largest grid fits shared memory, else a cluster plan
// (grid distributed over 2/4/8 CTAs), else the global-counter fallback. Fills plan->cells / bit_words.
bool scan_plan_for(const DevGeom* geoms, uint32_t n_geoms, uint32_t smem_optin, bool packed, ScanPlan* plan);
cudaError_t scan_configure(uint32_t smem_optin);
uint32_t scan_cluster_smem(uint32_t stages, uint32_t band_cells, uint32_t band_bit_words);
cudaError_t scan_cluster_configure(uint32_t smem_optin);
cudaError_t scan_cluster_launch(const ScanArgs& a, const ScanPlan& plan, int num_sms, cudaStream_t st);
uint32_t scan_grid(const ScanPlan& plan, int num_sms, uint32_t n_frames);
cudaError_t scan_launch(const ScanArgs& a, const ScanPlan& plan, int num_sms, cudaStream_t st);

// ---- K-C: compaction + sort/unique + gap merge + savings + decision (pipeline.cpp:297-404)
struct SegExtent {
  uint64_t start;  // first frame index into pts/flags
  uint64_t n;
};
struct SegJob {
  uint32_t ext_begin, ext_end;  // extents of this video, in order
  uint32_t reserved0, reserved1;
  uint64_t ts_base;   // offset into ts scratch A/B (capacity = pow2 >= frames)
  uint64_t ts_cap;    // power of two
  uint64_t seg_base;  // offset into segment output
  double duration;
};
struct SegArgs {
  const SegJob* jobs;
  const SegExtent* extents;
  const double* pts;
  const uint8_t* flags;
  double* ts_a;
  double* ts_b;
  mscan_segment* segs;
  mscan_video_result* results;
  double max_gap, padding, min_savings_pct;
};
cudaError_t segments_launch(const SegArgs& a, uint32_t n_videos, cudaStream_t st);

// ---- host side: projection of native records to mscan_mv8 (host_project.cpp; plain C++ with AVX-512 paths)
void project_records(const uint8_t* in, uint64_t n, uint64_t* out);
// mvz: projected records with static macroblocks elided (format: host_project.cpp). kMvzTileRecs records per tile.
constexpr uint32_t kMvzTileRecs = 1024;
uint64_t mvz_bound(uint64_t n_recs, uint64_t n_frames);  // bytes the encoding of n_recs records in n_frames frames can take, incl. slack
// encodes ONE frame of n native records at `out` (16-byte aligned); tile_end16[t] = end of tile t in 16-byte units
// from `out`; returns the bytes written (a multiple of 16)
uint64_t mvz_encode_frame(const uint8_t* in, uint32_t n, uint8_t* out, uint32_t* tile_end16);
void stream_copy(const uint8_t* from, uint8_t* to, uint64_t bytes);
// mscan_mv8 projections of the records with src != dst, in order (host_project.cpp); out: room for n records + 64 bytes
uint64_t compact_moving(const uint8_t* in, uint64_t n, uint64_t* out);

// ---- aux: exclusive scan of per-frame record counts, synthetic stream generation
cudaError_t offsets_launch(const uint32_t* counts, uint32_t n, uint64_t* off, uint64_t* block_scratch,
                           cudaStream_t st);
uint32_t offsets_scratch_elems(uint32_t n);
cudaError_t project_launch(const mscan_mv* recs, uint64_t n, mscan_mv8* out, cudaStream_t st);
cudaError_t synth_counts_launch(const mvgen_spec& spec, uint64_t frame0, uint32_t n_frames, uint32_t* counts,
                                cudaStream_t st);
cudaError_t synth_fill_launch(const mvgen_spec& spec, uint64_t frame0, uint32_t n_frames, const uint64_t* rec_off,
                              mscan_mv* recs, double* pts, cudaStream_t st);

}  // namespace mscan
