/*
 * ffshim.h — a FAKE libavformat/libavcodec/libavutil, just large enough to compile the reference's
 * own sources (src/motion_scanner.cpp, src/memory_io.cpp, src/pipeline.cpp) unmodified from
 * /root/reference. TEST INFRASTRUCTURE ONLY (oracle/_ref); nothing here is FFmpeg code.
 *
 * The "container" it demuxes is the MVS1 stream file (include/mvs_format.h): per frame a pts,
 * a keyframe bit and the AVMotionVector records export_mvs would have attached. "Decoding" a packet
 * hands those records back as AV_FRAME_DATA_MOTION_VECTORS side data, so MotionScanner::scan_range →
 * check_frame and ProcessingPipeline::run execute exactly the reference's code on our MV streams.
 */
#ifndef FFSHIM_H
#define FFSHIM_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- libavutil ---------------------------------------------------------------------------- */
typedef struct AVRational {
  int num, den;
} AVRational;
static inline double av_q2d(AVRational a) { return a.num / (double)a.den; }

#define AV_NOPTS_VALUE ((int64_t)UINT64_C(0x8000000000000000))
#define AV_TIME_BASE 1000000
#define AVERROR_EOF (-541478725)
#define AVERROR(e) (-(e))

void* av_malloc(size_t size);
void av_free(void* ptr);

typedef struct AVDictionary AVDictionary;
int av_dict_set(AVDictionary** pm, const char* key, const char* value, int flags);
void av_dict_free(AVDictionary** m);

enum AVMediaType { AVMEDIA_TYPE_UNKNOWN = -1, AVMEDIA_TYPE_VIDEO, AVMEDIA_TYPE_AUDIO };

enum AVFrameSideDataType { AV_FRAME_DATA_PANSCAN = 0, AV_FRAME_DATA_MOTION_VECTORS = 8 };

typedef struct AVFrameSideData {
  enum AVFrameSideDataType type;
  uint8_t* data;
  size_t size;
} AVFrameSideData;

typedef struct AVFrame {
  int64_t pts;
  int width, height;
  int key_frame;
  AVFrameSideData* shim_sd; /* NULL when the frame carries no motion vectors */
  AVFrameSideData shim_sd_storage;
} AVFrame;

AVFrame* av_frame_alloc(void);
void av_frame_free(AVFrame** frame);
AVFrameSideData* av_frame_get_side_data(const AVFrame* frame, enum AVFrameSideDataType type);

/* the 40-byte record export_mvs attaches (layout is FFmpeg's public ABI; only src/dst are read) */
typedef struct AVMotionVector {
  int32_t source;
  uint8_t w, h;
  int16_t src_x, src_y;
  int16_t dst_x, dst_y;
  uint64_t flags;
  int32_t motion_x, motion_y;
  uint16_t motion_scale;
} AVMotionVector;

/* ---- libavformat/avio ---------------------------------------------------------------------- */
#define AVSEEK_SIZE 0x10000
#define AVSEEK_FLAG_BACKWARD 1
#define AVFMT_FLAG_CUSTOM_IO 0x0080

typedef struct AVIOContext {
  unsigned char* buffer;
  int buffer_size;
  void* opaque;
  int (*read_packet)(void* opaque, uint8_t* buf, int buf_size);
  int64_t (*seek)(void* opaque, int64_t offset, int whence);
} AVIOContext;

AVIOContext* avio_alloc_context(unsigned char* buffer, int buffer_size, int write_flag, void* opaque,
                                int (*read_packet)(void* opaque, uint8_t* buf, int buf_size),
                                int (*write_packet)(void* opaque, const uint8_t* buf, int buf_size),
                                int64_t (*seek)(void* opaque, int64_t offset, int whence));
void avio_context_free(AVIOContext** s);

/* ---- libavcodec ---------------------------------------------------------------------------- */
enum AVCodecID { AV_CODEC_ID_NONE = 0, AV_CODEC_ID_H264 = 27, AV_CODEC_ID_HEVC = 173 };
enum AVDiscard { AVDISCARD_NONE = -16, AVDISCARD_DEFAULT = 0, AVDISCARD_BIDIR = 16, AVDISCARD_ALL = 48 };

#define AV_CODEC_FLAG_GRAY (1 << 13)
#define AV_CODEC_FLAG2_FAST (1 << 0)
#define FF_THREAD_FRAME 1
#define FF_THREAD_SLICE 2

typedef struct AVCodec {
  const char* name;
  enum AVCodecID id;
} AVCodec;

typedef struct AVCodecParameters {
  enum AVMediaType codec_type;
  enum AVCodecID codec_id;
  int width, height;
} AVCodecParameters;

typedef struct AVPacket {
  int stream_index;
  int64_t pts;
  int shim_frame; /* frame index inside the MVS1 stream, -1 = empty */
} AVPacket;

typedef struct AVCodecContext {
  int width, height;
  enum AVDiscard skip_loop_filter, skip_idct, skip_frame;
  int flags, flags2;
  int thread_count, thread_type;
  /* shim state */
  const struct MvsStream* shim_stream;
  int shim_pending; /* frame index waiting in the "decoder", -1 = none */
  int shim_export_mvs;
  uint8_t* shim_buf; /* records of the pending frame, as a demuxer/decoder would own them */
  size_t shim_buf_cap;
} AVCodecContext;

AVPacket* av_packet_alloc(void);
void av_packet_free(AVPacket** pkt);
void av_packet_unref(AVPacket* pkt);

const AVCodec* avcodec_find_decoder(enum AVCodecID id);
const AVCodec* avcodec_find_decoder_by_name(const char* name);
AVCodecContext* avcodec_alloc_context3(const AVCodec* codec);
void avcodec_free_context(AVCodecContext** avctx);
int avcodec_parameters_to_context(AVCodecContext* codec, const AVCodecParameters* par);
int avcodec_open2(AVCodecContext* avctx, const AVCodec* codec, AVDictionary** options);
void avcodec_flush_buffers(AVCodecContext* avctx);
int avcodec_send_packet(AVCodecContext* avctx, const AVPacket* avpkt);
int avcodec_receive_frame(AVCodecContext* avctx, AVFrame* frame);

/* ---- libavformat ---------------------------------------------------------------------------- */
typedef struct AVStream {
  int index;
  AVRational time_base;
  AVRational avg_frame_rate;
  AVCodecParameters* codecpar;
  enum AVDiscard discard;
} AVStream;

typedef struct AVFormatContext {
  AVIOContext* pb;
  int flags;
  unsigned int nb_streams;
  AVStream** streams;
  int64_t duration;
  /* shim state */
  struct MvsStream* shim_stream;
  int shim_next; /* next frame index av_read_frame returns */
} AVFormatContext;

AVFormatContext* avformat_alloc_context(void);
int avformat_open_input(AVFormatContext** ps, const char* url, const void* fmt, AVDictionary** options);
int avformat_find_stream_info(AVFormatContext* ic, AVDictionary** options);
void avformat_close_input(AVFormatContext** s);
int av_find_best_stream(AVFormatContext* ic, enum AVMediaType type, int wanted_stream_nb, int related_stream,
                        const AVCodec** decoder_ret, int flags);
int av_read_frame(AVFormatContext* s, AVPacket* pkt);
int av_seek_frame(AVFormatContext* s, int stream_index, int64_t timestamp, int flags);

#ifdef __cplusplus
}
#endif
#endif /* FFSHIM_H */
