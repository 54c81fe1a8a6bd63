/*
 * mscan_oracle.c — CPU restatement of the reference's motion-scan hot path.
 * TEST INFRASTRUCTURE ONLY (see mscan_oracle.h): never linked into the product.
 *
 * Follows, operation for operation (citations relative to /root/reference):
 *   orc_geometry          src/motion_scanner.cpp:189-196
 *   vote()                src/motion_scanner.cpp:229-268   (Phase 0 + Phase 1)
 *   orc_check_frame       src/motion_scanner.cpp:217-295   (Phase 2 with early exit :288-289)
 *   orc_full_count        same Phase 2 with the early exit removed
 *   orc_select_range      src/motion_scanner.cpp:303-371   (frame skip, seek, pts, range filter)
 *   orc_select_pipeline   src/pipeline.cpp:163-167         (chunk queue) + orc_select_range per chunk
 *   orc_merge_timestamps  src/pipeline.cpp:302-304
 *   orc_build_segments    src/pipeline.cpp:325-344
 *   orc_savings           src/pipeline.cpp:349-356
 *   orc_video_tail        src/pipeline.cpp:297-404 (+ src/motion_scanner.cpp:382-383)
 *
 * Parity domain notes:
 *   - int32 overflow of dx*dx+dy*dy is UB in the reference; here it wraps (what x86 gcc emits).
 *   - with vertical_margin == 0 the reference reads rows -1 / grid_h (out of bounds, UB); here
 *     out-of-grid neighbours are inactive.
 */
#include "mscan_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

enum { MV_STRIDE = 40, OFF_SRC_X = 6, OFF_SRC_Y = 8, OFF_DST_X = 10, OFF_DST_Y = 12 };

/* Layout the offsets above come from (FFmpeg 8.0 libavutil/motion_vector.h). */
struct orc_mv_layout {
  int32_t source;
  uint8_t w, h;
  int16_t src_x, src_y, dst_x, dst_y;
  uint64_t flags;
  int32_t motion_x, motion_y;
  uint16_t motion_scale;
};
_Static_assert(sizeof(struct orc_mv_layout) == MV_STRIDE, "AVMotionVector is 40 bytes");
_Static_assert(__builtin_offsetof(struct orc_mv_layout, src_x) == OFF_SRC_X, "src_x@6");
_Static_assert(__builtin_offsetof(struct orc_mv_layout, dst_y) == OFF_DST_Y, "dst_y@12");

static inline int rd16(const uint8_t* p) {
  int16_t v;
  memcpy(&v, p, 2);
  return (int)v;
}

void orc_geometry(int width, int height, int block_size, int block_shift, float vertical_mask,
                  int32_t* grid_w, int32_t* grid_h, int32_t* vertical_margin) {
  /* :189-192 — int arithmetic, then narrowed to int16_t members */
  int16_t gw = (int16_t)((width + block_size - 1) >> block_shift);
  int16_t gh = (int16_t)((height + block_size - 1) >> block_shift);
  /* :196 — int16 promoted to int, converted to float, float multiply, truncation */
  int margin = (int)((float)gh * vertical_mask);
  *grid_w = gw;
  *grid_h = gh;
  *vertical_margin = margin;
}

/* Phase 0 + Phase 1 (:229-268). Returns 0 when there is no side data (:219-221). */
static int vote(const orc_cfg* c, const uint8_t* recs, int64_t size_bytes, uint8_t* grid) {
  if (!recs) return 0;
  const int count = (int)(size_bytes / MV_STRIDE); /* :226 floor */
  const int gw = c->grid_w, gh = c->grid_h;
  memset(grid, 0, (size_t)gw * (size_t)gh); /* :229 */
  const double t2 = c->mv_threshold_sq;
  const int shift = c->block_shift;
  const int y_min = c->vertical_margin;
  const int y_max = gh - c->vertical_margin;
  for (int i = 0; i < count; ++i) {
    const uint8_t* r = recs + (size_t)i * MV_STRIDE;
    const int sx = rd16(r + OFF_SRC_X), sy = rd16(r + OFF_SRC_Y);
    const int tx = rd16(r + OFF_DST_X), ty = rd16(r + OFF_DST_Y);
    const int dx = tx - sx, dy = ty - sy;                                 /* :246-247 */
    const int mag = (int)((uint32_t)dx * (uint32_t)dx + (uint32_t)dy * (uint32_t)dy); /* :248 */
    if ((double)mag < t2) continue;                                       /* :251 int→double compare */
    const int gx = tx >> shift, gy = ty >> shift;                         /* :255-256 arithmetic shift */
    if (gx < 0 || gx >= gw || gy < y_min || gy >= y_max) continue;        /* :262 */
    uint8_t* cell = &grid[gy * gw + gx];
    if (*cell != 255) ++*cell;                                            /* :265-266 saturate */
  }
  return 1;
}

static inline int active_at(const orc_cfg* c, const uint8_t* grid, int x, int y) {
  if (y < 0 || y >= c->grid_h) return 0; /* defined behaviour for margin == 0 */
  return grid[y * c->grid_w + x] >= c->vectors_needed;
}

static inline int is_cluster_cell(const orc_cfg* c, const uint8_t* grid, int x, int y) {
  if (!active_at(c, grid, x, y)) return 0;                                /* :282 */
  return active_at(c, grid, x - 1, y) | active_at(c, grid, x + 1, y) |    /* :284-286 */
         active_at(c, grid, x, y - 1) | active_at(c, grid, x, y + 1);
}

int orc_check_frame(const orc_cfg* c, const void* recs, int64_t size_bytes, uint8_t* grid) {
  if (!vote(c, (const uint8_t*)recs, size_bytes, grid)) return 0;
  const int y_min = c->vertical_margin, y_max = c->grid_h - c->vertical_margin;
  int clusters = 0;
  for (int y = y_min; y < y_max; ++y)
    for (int x = 1; x < c->grid_w - 1; ++x) /* :279-280 columns 0 and gw-1 are never centres */
      if (is_cluster_cell(c, grid, x, y))
        if (++clusters >= c->clusters_needed) return 1; /* :288-289 */
  return 0;
}

uint32_t orc_full_count(const orc_cfg* c, const void* recs, int64_t size_bytes, uint8_t* grid) {
  if (!vote(c, (const uint8_t*)recs, size_bytes, grid)) return 0;
  const int y_min = c->vertical_margin, y_max = c->grid_h - c->vertical_margin;
  uint32_t clusters = 0;
  for (int y = y_min; y < y_max; ++y)
    for (int x = 1; x < c->grid_w - 1; ++x) clusters += (uint32_t)is_cluster_cell(c, grid, x, y);
  return clusters;
}

static void scan_range(const orc_cfg* c, const uint8_t* recs, const uint64_t* rec_off, uint32_t f0,
                       uint32_t f1, uint8_t* flags, uint32_t* counts, int early_exit, uint8_t* grid) {
  const uint32_t need = c->clusters_needed < 1 ? 1u : (uint32_t)c->clusters_needed;
  for (uint32_t f = f0; f < f1; ++f) {
    const uint64_t n = rec_off[f + 1] - rec_off[f];
    const uint8_t* p = n ? recs + rec_off[f] * MV_STRIDE : NULL; /* zero records ⇒ no side data */
    if (early_exit) {
      flags[f] = (uint8_t)orc_check_frame(c, p, (int64_t)n * MV_STRIDE, grid);
    } else {
      const uint32_t k = orc_full_count(c, p, (int64_t)n * MV_STRIDE, grid);
      if (counts) counts[f] = k;
      if (flags) flags[f] = (uint8_t)(k >= need);
    }
  }
}

void orc_scan_frames(const orc_cfg* c, const void* recs, const uint64_t* rec_off, uint32_t n_frames,
                     uint8_t* flags, uint32_t* counts, int early_exit) {
  uint8_t* grid = (uint8_t*)malloc((size_t)c->grid_w * (size_t)c->grid_h + 1);
  scan_range(c, (const uint8_t*)recs, rec_off, 0, n_frames, flags, counts, early_exit, grid);
  free(grid);
}

struct mt_job {
  const orc_cfg* c;
  const uint8_t* recs;
  const uint64_t* rec_off;
  uint32_t f0, f1;
  uint8_t* flags;
  uint32_t* counts;
  int early_exit;
};

static void* mt_main(void* arg) {
  struct mt_job* j = (struct mt_job*)arg;
  uint8_t* grid = (uint8_t*)malloc((size_t)j->c->grid_w * (size_t)j->c->grid_h + 1);
  scan_range(j->c, j->recs, j->rec_off, j->f0, j->f1, j->flags, j->counts, j->early_exit, grid);
  free(grid);
  return NULL;
}

void orc_scan_frames_mt(const orc_cfg* c, const void* recs, const uint64_t* rec_off, uint32_t n_frames,
                        uint8_t* flags, uint32_t* counts, int early_exit, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if ((uint32_t)n_threads > n_frames) n_threads = n_frames ? (int)n_frames : 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
  struct mt_job* jobs = (struct mt_job*)malloc(sizeof(struct mt_job) * (size_t)n_threads);
  /* contiguous ranges balanced by record count, not frame count */
  const uint64_t total = rec_off[n_frames] - rec_off[0];
  uint32_t f = 0;
  for (int t = 0; t < n_threads; ++t) {
    const uint64_t target = rec_off[0] + total * (uint64_t)(t + 1) / (uint64_t)n_threads;
    uint32_t e = f;
    if (t == n_threads - 1) e = n_frames;
    else
      while (e < n_frames && rec_off[e + 1] <= target) ++e;
    jobs[t] = (struct mt_job){c, (const uint8_t*)recs, rec_off, f, e, flags, counts, early_exit};
    f = e;
    pthread_create(&th[t], NULL, mt_main, &jobs[t]);
  }
  for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
  free(jobs);
  free(th);
}

uint32_t orc_select_range(const int64_t* pts_ticks, const uint8_t* is_key, uint32_t n_frames, double time_base,
                          double video_fps, double target_fps, double start, double end, uint32_t* out_idx) {
  const int frame_skip = (target_fps > 0 && target_fps < video_fps) ? (int)(video_fps / target_fps) : 1; /* :310-313 */
  int frame_count = 0;                                                                                   /* :314 */
  uint32_t i = 0;
  if (start > 0) {                                                                                       /* :321-325 */
    const int64_t seek_ts = (int64_t)(start / time_base);
    uint32_t best = 0;
    for (uint32_t k = 0; k < n_frames; ++k) {
      if (!is_key[k]) continue;
      if (pts_ticks[k] <= seek_ts) best = k;
      else break;
    }
    i = best;
  }
  uint32_t n = 0;
  for (; i < n_frames; ++i) {
    if (++frame_count % frame_skip != 0) continue;                                                       /* :357 */
    const double pts = (double)pts_ticks[i] * time_base;                                                 /* :361 */
    if (pts < start) continue;                                                                           /* :364 */
    if (pts >= end) break;                                                                               /* :368 */
    out_idx[n++] = i;                                                                                    /* :376 */
  }
  return n;
}

uint32_t orc_select_pipeline(const int64_t* pts_ticks, const uint8_t* is_key, uint32_t n_frames, double time_base,
                             double video_fps, double target_fps, double duration, double chunk_sec, uint32_t* out_idx) {
  uint32_t n = 0;
  for (double t = 0; t < duration; t += chunk_sec) {                                                     /* :163 */
    const double end = (t + chunk_sec < duration) ? t + chunk_sec : duration;                            /* :164 std::min */
    n += orc_select_range(pts_ticks, is_key, n_frames, time_base, video_fps, target_fps, t, end, out_idx + n);
  }
  return n;
}

static int cmp_double(const void* a, const void* b) {
  const double x = *(const double*)a, y = *(const double*)b;
  return (x < y) ? -1 : (y < x) ? 1 : 0;
}

uint32_t orc_merge_timestamps(double* ts, uint32_t n) {
  if (n == 0) return 0;
  qsort(ts, n, sizeof(double), cmp_double); /* :302 ascending by operator< */
  uint32_t w = 0;                           /* :303-304 unique with == keeps the first of a run */
  for (uint32_t i = 1; i < n; ++i)
    if (!(ts[w] == ts[i])) ts[++w] = ts[i];
  return w + 1;
}

static inline double max_like_std(double a, double b) { return (a < b) ? b : a; } /* std::max(a,b) */
static inline double min_like_std(double a, double b) { return (b < a) ? b : a; } /* std::min(a,b) */

uint32_t orc_build_segments(const double* ts, uint32_t n, double max_gap, double padding, orc_segment* out) {
  uint32_t k = 0;
  double curr_start = ts[0], last_act = ts[0]; /* :325-326 */
  for (uint32_t i = 1; i < n; ++i) {
    const double gap = ts[i] - last_act;       /* :329 */
    if (gap > max_gap) {                       /* :330 strict */
      out[k].start = max_like_std(0.0, curr_start - padding); /* :337 */
      out[k].end = last_act + padding;                        /* :338 */
      ++k;
      curr_start = ts[i];
    }
    last_act = ts[i];
  }
  out[k].start = max_like_std(0.0, curr_start - padding);     /* :343-344 */
  out[k].end = last_act + padding;
  return k + 1;
}

void orc_savings(orc_segment* segs, uint32_t n, double duration, double* out_dur, double* time_removed,
                 double* saved_pct) {
  double sum = 0;
  for (uint32_t i = 0; i < n; ++i) {          /* :350-354, in order */
    segs[i].end = min_like_std(segs[i].end, duration);
    segs[i].start = min_like_std(segs[i].start, segs[i].end);
    sum += (segs[i].end - segs[i].start);
  }
  *out_dur = sum;
  *time_removed = duration - sum;                                     /* :355 */
  *saved_pct = (duration > 0) ? *time_removed / duration * 100.0 : 0.0; /* :356 */
}

void orc_video_tail(const double* pts, const uint8_t* flags, uint32_t n_frames, double duration,
                    double max_gap, double padding, double min_savings_pct, orc_segment* segs,
                    orc_result* res) {
  memset(res, 0, sizeof(*res));
  double* ts = (double*)malloc(sizeof(double) * (n_frames ? n_frames : 1));
  uint32_t n = 0;
  for (uint32_t i = 0; i < n_frames; ++i)
    if (flags[i]) ts[n++] = pts[i];            /* motion_scanner.cpp:382-383 */
  n = orc_merge_timestamps(ts, n);
  res->n_motion_frames = n;
  if (n == 0) {                                /* :308-319 — no job at all */
    res->decision = 0;
    free(ts);
    return;
  }
  res->n_segments = orc_build_segments(ts, n, max_gap, padding, segs);
  orc_savings(segs, res->n_segments, duration, &res->out_dur, &res->time_removed, &res->saved_pct);
  res->decision = (res->saved_pct > min_savings_pct) ? 1 : 2; /* :358 strict */
  free(ts);
}
