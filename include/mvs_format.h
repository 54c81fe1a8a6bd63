/*
 * mvs_format.h — the MVS1 motion-vector stream file (shared by the product CLI and the test shim): what FFmpeg's export_mvs decode of one video
 * would hand to the scanner, frozen to disk. Used as the input container of the fake-libav shim
 * (oracle/_ref) and as the decode-free input of the product CLI.
 *
 *   MvsHeader | MvsFrameEntry[n_frames] | AVMotionVector records (40 B each), frames back to back
 *
 * All fields little-endian; records start at a 64-byte aligned offset.
 */
#ifndef MVS_FORMAT_H
#define MVS_FORMAT_H

#include <stdint.h>

#define MVS_MAGIC "MVSTRM01"

typedef struct MvsHeader {
  char magic[8];
  int32_t width, height;      /* AVCodecContext width/height (display size) */
  int32_t tb_num, tb_den;     /* stream time_base                            */
  int32_t fps_num, fps_den;   /* avg_frame_rate                              */
  int64_t duration_us;        /* AVFormatContext duration (AV_TIME_BASE units) */
  uint32_t n_frames;
  uint32_t reserved;
  uint64_t records_offset;    /* byte offset of the first record             */
  uint64_t n_records;
} MvsHeader;

#define MVS_FRAME_KEY 1u      /* seek target (I-frame)                       */
#define MVS_FRAME_HAS_MVS 2u  /* frame carries AV_FRAME_DATA_MOTION_VECTORS  */

typedef struct MvsFrameEntry {
  int64_t pts;                /* in time_base ticks                          */
  uint64_t first_record;      /* index of the frame's first record           */
  uint32_t n_records;
  uint32_t flags;
} MvsFrameEntry;

#endif
