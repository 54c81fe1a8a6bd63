/*
 * mvgen_core.h — deterministic synthetic AVMotionVector streams (measurement harness).
 *
 * One definition compiled both by gcc (tests, CPU baseline) and by nvcc (device-side generation of
 * the 10^9-record stream without a 40 GB host buffer). Integer-only and counter-based: every
 * record is a pure function of (spec, global frame index, macroblock, partition), so any frame
 * range can be regenerated anywhere, bit-identically.
 *
 * Structure follows what FFmpeg's export_mvs emits for H.264 P-frames:
 * macroblocks in raster order, 1 / 2 / 4 records per MB (16x16, 16x8|8x16, 8x8), dst = partition
 * centre, src = dst - displacement, I-frames carry no records. The reference never generates
 * streams; only the record layout (mscan_mv == AVMotionVector) is shared with it.
 */
#ifndef MVGEN_CORE_H
#define MVGEN_CORE_H

#include <stdint.h>

#if defined(__CUDACC__)
#define MVGEN_HD __host__ __device__ __forceinline__
#else
#define MVGEN_HD static inline
#endif

#define MVGEN_MAX_BLOBS 4

typedef struct mvgen_spec {
  uint64_t seed;
  int32_t width, height;    /* pixels                                                      */
  int32_t gop;              /* local frame % gop == 0 → I-frame, 0 records (0 = no I-frames) */
  int32_t window;           /* activity window, frames (>= 8)                              */
  int32_t frames_per_video; /* pts and windows restart every this many frames (0 = never)  */
  uint32_t p_window_active; /* /1024: a window contains moving blobs                       */
  int32_t max_blobs;        /* 1..MVGEN_MAX_BLOBS                                          */
  uint32_t p_split2;        /* /1024: static MB exported as two partitions                 */
  uint32_t p_noise;         /* /65536: isolated 8x8-split MB with |d| components in -2..2  */
  uint32_t p_single;        /* /65536: moving 16x16 MB (one record, d in [-8,8]^2)         */
  uint32_t p_oob;           /* /65536 per record: dst pushed outside the picture           */
  int32_t dense;            /* 1: every MB is 8x8-split (4K dense-field shape)             */
  uint32_t p_dense_move;    /* /1024: 4x4-MB tiles moving in dense mode                    */
  int32_t static_a0, static_a1, static_b0, static_b1; /* forced-static local frame ranges  */
  double fps;
  int32_t scatter;          /* 1: SURVEY §8(d) config-5 shape — every MB exports 2 records whose dst is a uniformly
                               random MB sub-centre of the picture (no raster order, no spatial coherence)      */
  uint32_t p_rec_move;      /* /65536 per record (scatter mode): moving, d uniform in [-8,8]^2               */
} mvgen_spec;

/* Per-frame state: which blobs are alive and where (MB units). */
typedef struct mvgen_frame {
  uint64_t gframe;  /* global frame index   */
  uint32_t video;   /* gframe / frames_per_video */
  uint32_t lframe;  /* local frame index    */
  int32_t mbw, mbh;
  int32_t iframe;   /* 1 → no records       */
  int32_t quiet;    /* 1 → forced static    */
  int32_t nb;
  int32_t cx[MVGEN_MAX_BLOBS], cy[MVGEN_MAX_BLOBS];
  int32_t rx[MVGEN_MAX_BLOBS], ry[MVGEN_MAX_BLOBS];
  int32_t dx[MVGEN_MAX_BLOBS], dy[MVGEN_MAX_BLOBS];
} mvgen_frame;

/* Per-macroblock decision. kind: 0 static 16x16, 1 static 2-part, 2 noise 8x8, 3 single moving,
 * 4 blob 8x8, 5 dense static 8x8, 6 dense moving 8x8, 7 scattered (position and motion drawn per record). */
typedef struct mvgen_mb {
  int32_t nrec;
  int32_t kind;
  int32_t dx, dy;
  uint64_t h;
} mvgen_mb;

MVGEN_HD uint64_t mvgen_mix(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

MVGEN_HD uint64_t mvgen_hash(uint64_t seed, uint64_t a, uint64_t b, uint64_t c) {
  uint64_t h = mvgen_mix(seed ^ 0x6D76675F62323030ull);
  h = mvgen_mix(h ^ a);
  h = mvgen_mix(h ^ (b * 0xD6E8FEB86659FD93ull));
  h = mvgen_mix(h ^ (c * 0xA24BAED4963EE407ull));
  return h;
}

MVGEN_HD double mvgen_pts(const mvgen_spec* s, uint64_t gframe) {
  uint64_t l = s->frames_per_video > 0 ? gframe % (uint64_t)s->frames_per_video : gframe;
  return (double)l / s->fps;
}

MVGEN_HD void mvgen_frame_init(const mvgen_spec* s, uint64_t gframe, mvgen_frame* fr) {
  fr->gframe = gframe;
  if (s->frames_per_video > 0) {
    fr->video = (uint32_t)(gframe / (uint64_t)s->frames_per_video);
    fr->lframe = (uint32_t)(gframe % (uint64_t)s->frames_per_video);
  } else {
    fr->video = 0;
    fr->lframe = (uint32_t)gframe;
  }
  fr->mbw = (s->width + 15) >> 4;
  fr->mbh = (s->height + 15) >> 4;
  fr->iframe = (s->gop > 0 && (fr->lframe % (uint32_t)s->gop) == 0) ? 1 : 0;
  const int32_t lf = (int32_t)fr->lframe;
  fr->quiet = ((lf >= s->static_a0 && lf < s->static_a1) || (lf >= s->static_b0 && lf < s->static_b1)) ? 1 : 0;
  fr->nb = 0;
  if (fr->iframe || fr->quiet || s->dense) return;
  const int32_t win = s->window < 8 ? 8 : s->window;
  const uint32_t wi = fr->lframe / (uint32_t)win;
  const int32_t t = (int32_t)(fr->lframe - wi * (uint32_t)win);
  const uint64_t hw = mvgen_hash(s->seed, fr->video, wi, 0xB10Bull);
  if ((uint32_t)(hw & 1023u) >= s->p_window_active) return;
  int32_t mb = s->max_blobs < 1 ? 1 : (s->max_blobs > MVGEN_MAX_BLOBS ? MVGEN_MAX_BLOBS : s->max_blobs);
  const int32_t nb = 1 + (int32_t)((hw >> 10) % (uint64_t)mb);
  for (int32_t b = 0; b < nb; ++b) {
    const uint64_t hb = mvgen_hash(s->seed, fr->video, wi, 100u + (uint64_t)b);
    const uint64_t h2 = mvgen_mix(hb ^ 0x51ED270B1ull);
    const int32_t t0 = (int32_t)((hb >> 52) % (uint64_t)(win / 2));
    const int32_t dur = win / 4 + (int32_t)((h2 >> 40) % (uint64_t)(win / 2));
    if (t < t0 || t >= t0 + dur) continue;
    const int32_t x0 = (int32_t)((hb >> 8) % (uint64_t)fr->mbw);
    const int32_t y0 = (int32_t)((hb >> 24) % (uint64_t)fr->mbh);
    const int32_t vx = (int32_t)((hb >> 40) & 63u) - 32; /* 1/256 MB per frame */
    const int32_t vy = (int32_t)((hb >> 46) & 63u) - 32;
    int32_t ddx = (int32_t)(h2 % 17u) - 8;
    int32_t ddy = (int32_t)((h2 >> 8) % 17u) - 8;
    if (ddx == 0 && ddy == 0) ddx = 3;
    const int32_t k = fr->nb++;
    fr->cx[k] = ((x0 << 8) + vx * (t - t0)) >> 8;
    fr->cy[k] = ((y0 << 8) + vy * (t - t0)) >> 8;
    fr->rx[k] = 1 + (int32_t)(hb & 3u);
    fr->ry[k] = 1 + (int32_t)((hb >> 2) & 3u);
    fr->dx[k] = ddx;
    fr->dy[k] = ddy;
  }
}

MVGEN_HD void mvgen_mb_eval(const mvgen_spec* s, const mvgen_frame* fr, int32_t mx, int32_t my, mvgen_mb* m) {
  m->dx = 0;
  m->dy = 0;
  m->h = 0;
  if (fr->iframe) {
    m->nrec = 0;
    m->kind = 0;
    return;
  }
  const uint64_t mbid = (uint64_t)my * (uint64_t)fr->mbw + (uint64_t)mx;
  const uint64_t h = mvgen_hash(s->seed, fr->gframe, mbid, 0x4D42ull);
  m->h = h;
  if (s->scatter) {
    m->nrec = 2;
    m->kind = 7;
    return;
  }
  if (s->dense) {
    m->nrec = 4;
    /* moving areas are 4x4-MB tiles that persist for 8 frames → spatially adjacent active cells */
    const uint64_t ht = mvgen_hash(s->seed, fr->gframe >> 3, (uint64_t)(my >> 2) * 4096u + (uint64_t)(mx >> 2), 0xDE45Eull);
    if (!fr->quiet && (uint32_t)(ht & 1023u) < s->p_dense_move) {
      m->kind = 6;
      int32_t ddx = (int32_t)((ht >> 10) % 17u) - 8;
      int32_t ddy = (int32_t)((ht >> 20) % 17u) - 8;
      if (ddx == 0 && ddy == 0) ddy = -3;
      m->dx = ddx;
      m->dy = ddy;
    } else {
      m->kind = 5;
    }
    return;
  }
  if (!fr->quiet) {
    for (int32_t b = 0; b < fr->nb; ++b) {
      int32_t ax = mx - fr->cx[b];
      int32_t ay = my - fr->cy[b];
      if (ax < 0) ax = -ax;
      if (ay < 0) ay = -ay;
      if (ax <= fr->rx[b] && ay <= fr->ry[b]) {
        m->nrec = 4;
        m->kind = 4;
        m->dx = fr->dx[b];
        m->dy = fr->dy[b];
        return;
      }
    }
    const uint32_t r16 = (uint32_t)(h & 0xFFFFu);
    if (r16 < s->p_noise) {
      m->nrec = 4;
      m->kind = 2;
      return;
    }
    if (r16 < s->p_noise + s->p_single) {
      m->nrec = 1;
      m->kind = 3;
      int32_t ddx = (int32_t)((h >> 16) % 17u) - 8;
      int32_t ddy = (int32_t)((h >> 24) % 17u) - 8;
      m->dx = ddx;
      m->dy = ddy;
      return;
    }
  }
  if ((uint32_t)((h >> 32) & 1023u) < s->p_split2) {
    m->nrec = 2;
    m->kind = 1;
  } else {
    m->nrec = 1;
    m->kind = 0;
  }
}

/* Fields of record k (0 <= k < m->nrec) of macroblock (mx,my). Layout-free: the caller stores
 * them into a 40-byte mscan_mv / AVMotionVector. */
typedef struct mvgen_rec {
  int32_t source;
  int32_t w, h;
  int32_t src_x, src_y, dst_x, dst_y;
  int32_t motion_x, motion_y;
} mvgen_rec;

MVGEN_HD void mvgen_record(const mvgen_spec* s, const mvgen_mb* m, int32_t mx, int32_t my, int32_t k, mvgen_rec* r) {
  int32_t ox = 8, oy = 8, w = 16, h = 16;
  if (m->nrec == 4) {
    ox = (k & 1) ? 12 : 4;
    oy = (k & 2) ? 12 : 4;
    w = 8;
    h = 8;
  } else if (m->nrec == 2) {
    if ((m->h >> 42) & 1u) { /* 16x8: two rows */
      oy = k ? 12 : 4;
      h = 8;
    } else { /* 8x16: two columns */
      ox = k ? 12 : 4;
      w = 8;
    }
  }
  int32_t dx = m->dx, dy = m->dy;
  const uint64_t hk = mvgen_mix(m->h ^ ((uint64_t)(k + 1) * 0x9FB21C651E98DF25ull));
  if (m->kind == 2) { /* noise: components in -2..2, straddles the MV_THRESHOLD_SQ test */
    dx = (int32_t)(hk % 5u) - 2;
    dy = (int32_t)((hk >> 8) % 5u) - 2;
  } else if (m->kind == 4 || m->kind == 6) { /* jitter -1..1 around the blob displacement */
    dx += (int32_t)(hk % 3u) - 1;
    dy += (int32_t)((hk >> 8) % 3u) - 1;
  }
  int32_t X = (mx << 4) + ox;
  int32_t Y = (my << 4) + oy;
  if (m->kind == 7) { /* scattered: dst uniform over the picture's 8x8 sub-centres, 10 % moving uniformly */
    const uint64_t hp = mvgen_mix(hk ^ 0x5CA77E2ull);
    const int32_t mbw = (s->width + 15) >> 4, mbh = (s->height + 15) >> 4;
    X = ((int32_t)(hp % (uint64_t)mbw) << 4) + (((hp >> 40) & 1u) ? 12 : 4);
    Y = ((int32_t)((hp >> 20) % (uint64_t)mbh) << 4) + (((hp >> 41) & 1u) ? 12 : 4);
    w = 8;
    h = 8;
    dx = 0;
    dy = 0;
    if ((uint32_t)((hp >> 44) & 0xFFFFu) < s->p_rec_move) {
      dx = (int32_t)((hk >> 32) % 17u) - 8;
      dy = (int32_t)((hk >> 40) % 17u) - 8;
    }
  }
  if ((uint32_t)((hk >> 16) & 0xFFFFu) < s->p_oob) {
    const uint32_t mode = (uint32_t)(hk >> 32) & 3u;
    const int32_t off = (int32_t)((hk >> 34) & 31u);
    if (mode == 0) X = -1 - off;
    else if (mode == 1) X = s->width + off;
    else if (mode == 2) Y = -1 - off;
    else Y = ((s->height + 15) & ~15) + off;
  }
  r->source = -1;
  r->w = w;
  r->h = h;
  r->dst_x = X;
  r->dst_y = Y;
  r->src_x = X - dx;
  r->src_y = Y - dy;
  r->motion_x = -dx * 4;
  r->motion_y = -dy * 4;
}

/* Presets for BASELINE.json's configs. */
MVGEN_HD void mvgen_preset(mvgen_spec* s, int config, uint64_t seed) {
  s->seed = seed;
  s->width = 1920;
  s->height = 1080;
  s->gop = 30;
  s->window = 90;
  s->frames_per_video = 0;
  s->p_window_active = 102; /* ~10 % of windows */
  s->max_blobs = 3;
  s->p_split2 = 256;
  s->p_noise = 328; /* 0.5 % */
  s->p_single = 131; /* 0.2 % */
  s->p_oob = 66;     /* 0.1 % */
  s->dense = 0;
  s->p_dense_move = 0;
  s->static_a0 = s->static_a1 = s->static_b0 = s->static_b1 = 0;
  s->fps = 30.0;
  s->scatter = 0;
  s->p_rec_move = 0;
  if (config == 5) { /* SURVEY §8(d) config 5 as specified: 16 320 records per frame, uniform dst, 10 % moving, 0.1 % OOB */
    s->gop = 0;
    s->frames_per_video = 18000;
    s->scatter = 1;
    s->p_rec_move = 6554; /* 10 % */
    return;
  }
  if (config == 0) { /* 60 s 1080p30, static spans [300,900) and [1200,1500) */
    s->frames_per_video = 1800;
    s->p_window_active = 1024;
    s->window = 60;
    s->static_a0 = 300;
    s->static_a1 = 900;
    s->static_b0 = 1200;
    s->static_b1 = 1500;
  } else if (config == 1) { /* 10 min 1080p30 CCTV */
    s->frames_per_video = 18000;
  } else if (config == 2) { /* 4K30 dense field, 2 min */
    s->width = 3840;
    s->height = 2160;
    s->frames_per_video = 3600;
    s->dense = 1;
    s->p_dense_move = 307; /* 30 % */
  } else if (config == 3) { /* batch clip: like config 0 without forced spans, busier */
    s->frames_per_video = 1800;
    s->p_window_active = 300;
    s->window = 60;
  } else { /* config 4: 10^9-record stream, CCTV mix cut into 10-minute videos */
    s->frames_per_video = 18000;
    s->p_window_active = 256;
  }
}

#endif /* MVGEN_CORE_H */
