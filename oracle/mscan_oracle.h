/*
 * mscan_oracle.h — CPU oracle for the motion-scan hot path. TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call this. The product (libmotionscan.so) never does; it has no CPU fallback.
 *
 * Pinning status: the reference ships no tests, fixtures or golden vectors for this path
 * (SURVEY.md §8(c)). The restatement is pinned by (1) the derived known-answer vectors K1-K20 /
 * S1-S8 of SURVEY.md §4 and (2) outputs of the reference's OWN sources compiled here against a
 * fake-libav shim (oracle/_ref, see oracle/ref_harness.cpp and tests/golden/). See DESIGN.md §3.
 */
#ifndef MSCAN_ORACLE_H
#define MSCAN_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cfg struct of include/motion_trim/motion_scanner.hpp:86-93 + grid_w/grid_h (:75-77) */
typedef struct orc_cfg {
  double mv_threshold_sq;
  int32_t block_shift;
  int32_t clusters_needed;
  int32_t vertical_margin;
  int32_t grid_w;
  int32_t grid_h;
  uint8_t vectors_needed;
} orc_cfg;

typedef struct orc_segment {
  double start, end;
} orc_segment;

typedef struct orc_result {
  int32_t decision; /* 0 no motion, 1 cut, 2 full copy */
  uint32_t n_motion_frames;
  uint32_t n_segments;
  uint32_t reserved;
  double out_dur, time_removed, saved_pct;
} orc_result;

/* src/motion_scanner.cpp:189-196 */
void orc_geometry(int width, int height, int block_size, int block_shift, float vertical_mask,
                  int32_t* grid_w, int32_t* grid_h, int32_t* vertical_margin);

/* src/motion_scanner.cpp:217-295, reference semantics (early exit). recs == NULL ⇒ no side data.
 * grid: scratch of grid_w*grid_h bytes. */
int orc_check_frame(const orc_cfg* c, const void* recs, int64_t size_bytes, uint8_t* grid);

/* Same vote phase, cluster pass without the early exit (:288-289 removed): the full count. */
uint32_t orc_full_count(const orc_cfg* c, const void* recs, int64_t size_bytes, uint8_t* grid);

/* Batch drivers: frame i owns records [rec_off[i], rec_off[i+1]); zero records ⇒ no side data.
 * early_exit != 0: flags only via orc_check_frame (counts untouched);
 * early_exit == 0: counts via orc_full_count and flags = count >= max(1, clusters_needed). */
void orc_scan_frames(const orc_cfg* c, const void* recs, const uint64_t* rec_off, uint32_t n_frames,
                     uint8_t* flags, uint32_t* counts, int early_exit);
/* pthread version, contiguous frame ranges, one private grid per thread (pipeline.cpp:197). */
void orc_scan_frames_mt(const orc_cfg* c, const void* recs, const uint64_t* rec_off, uint32_t n_frames,
                        uint8_t* flags, uint32_t* counts, int early_exit, int n_threads);

/* src/motion_scanner.cpp:303-371: the frames ONE scan_range(start,end) call hands to check_frame.
 * pts_ticks/is_key describe the stream in decode order; the seek lands on the last key frame whose
 * pts <= (int64)(start/time_base) (AVSEEK_FLAG_BACKWARD; frame 0 when there is none), the skip counter
 * starts there (:314,:357 precede the range test :364). Writes frame indices, returns how many. */
uint32_t orc_select_range(const int64_t* pts_ticks, const uint8_t* is_key, uint32_t n_frames, double time_base,
                          double video_fps, double target_fps, double start, double end, uint32_t* out_idx);
/* src/pipeline.cpp:163-167 chunking + one orc_select_range per chunk; indices in chunk order. */
uint32_t orc_select_pipeline(const int64_t* pts_ticks, const uint8_t* is_key, uint32_t n_frames, double time_base,
                             double video_fps, double target_fps, double duration, double chunk_sec, uint32_t* out_idx);

/* src/pipeline.cpp:302-304: sort + unique in place; returns the new length. */
uint32_t orc_merge_timestamps(double* ts, uint32_t n);
/* src/pipeline.cpp:325-344: returns the number of segments written (ts sorted-unique, n >= 1). */
uint32_t orc_build_segments(const double* ts, uint32_t n, double max_gap, double padding, orc_segment* out);
/* src/pipeline.cpp:349-356 (clamps in place). */
void orc_savings(orc_segment* segs, uint32_t n, double duration, double* out_dur, double* time_removed,
                 double* saved_pct);
/* Whole tail: flags+pts → merged ts → segments → savings → decision (pipeline.cpp:297-404).
 * segs capacity >= n_frames. Writes the clamped motion segments whatever the decision. */
void orc_video_tail(const double* pts, const uint8_t* flags, uint32_t n_frames, double duration,
                    double max_gap, double padding, double min_savings_pct, orc_segment* segs,
                    orc_result* res);

#ifdef __cplusplus
}
#endif
#endif
