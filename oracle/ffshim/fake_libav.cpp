// fake_libav.cpp — implementation of the fake libav* declared in ffshim.h. TEST INFRASTRUCTURE ONLY.
//
// Demuxes the MVS1 stream file (mvs_format.h) through the caller's AVIO callbacks (the reference's
// MemoryLoader::read/seek over its mmap), and "decodes" by copying the frame's AVMotionVector
// records into a decoder-owned buffer and attaching them as AV_FRAME_DATA_MOTION_VECTORS side data —
// the same hand-off FFmpeg's export_mvs performs (motion_scanner.cpp:219-226 reads it back).
// Seeking follows AVSEEK_FLAG_BACKWARD: the last key frame whose pts <= the requested timestamp.
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "ffshim.h"
#include "mvs_format.h"

struct MvsStream {
  MvsHeader hdr;
  std::vector<MvsFrameEntry> frames;
  AVIOContext* io;  // not owned
  AVStream stream;
  AVStream* stream_ptr;
  AVCodecParameters par;
};

namespace {

bool io_read_at(AVIOContext* io, int64_t off, void* dst, size_t n) {
  if (io->seek(io->opaque, off, SEEK_SET) < 0) return false;
  uint8_t* p = static_cast<uint8_t*>(dst);
  while (n) {
    const int chunk = n > (1u << 30) ? (1 << 30) : (int)n;
    const int got = io->read_packet(io->opaque, p, chunk);
    if (got <= 0) return false;
    p += got;
    n -= (size_t)got;
  }
  return true;
}

const AVCodec kH264{"h264", AV_CODEC_ID_H264};
const AVCodec kHevc{"hevc", AV_CODEC_ID_HEVC};

}  // namespace

extern "C" {

void* av_malloc(size_t size) { return std::malloc(size); }
void av_free(void* p) { std::free(p); }

struct AVDictionary {
  int export_mvs;
};
int av_dict_set(AVDictionary** pm, const char* key, const char* value, int) {
  if (!*pm) *pm = new AVDictionary{0};
  if (key && value && !std::strcmp(key, "flags2") && std::strstr(value, "export_mvs")) (*pm)->export_mvs = 1;
  return 0;
}
void av_dict_free(AVDictionary** m) {
  delete *m;
  *m = nullptr;
}

AVFrame* av_frame_alloc(void) {
  AVFrame* f = new (std::nothrow) AVFrame();
  if (f) {
    f->pts = AV_NOPTS_VALUE;
    f->shim_sd = nullptr;
  }
  return f;
}
void av_frame_free(AVFrame** f) {
  delete *f;
  *f = nullptr;
}
AVFrameSideData* av_frame_get_side_data(const AVFrame* f, enum AVFrameSideDataType type) {
  return (type == AV_FRAME_DATA_MOTION_VECTORS) ? f->shim_sd : nullptr;
}

AVIOContext* avio_alloc_context(unsigned char* buffer, int buffer_size, int, void* opaque,
                                int (*read_packet)(void*, uint8_t*, int), int (*)(void*, const uint8_t*, int),
                                int64_t (*seek)(void*, int64_t, int)) {
  AVIOContext* c = new (std::nothrow) AVIOContext();
  if (!c) return nullptr;
  c->buffer = buffer;
  c->buffer_size = buffer_size;
  c->opaque = opaque;
  c->read_packet = read_packet;
  c->seek = seek;
  return c;
}
void avio_context_free(AVIOContext** s) {
  if (*s) {
    std::free((*s)->buffer);
    delete *s;
  }
  *s = nullptr;
}

AVPacket* av_packet_alloc(void) {
  AVPacket* p = new (std::nothrow) AVPacket();
  if (p) p->shim_frame = -1;
  return p;
}
void av_packet_free(AVPacket** p) {
  delete *p;
  *p = nullptr;
}
void av_packet_unref(AVPacket* p) { p->shim_frame = -1; }

const AVCodec* avcodec_find_decoder(enum AVCodecID id) {
  if (id == AV_CODEC_ID_H264) return &kH264;
  if (id == AV_CODEC_ID_HEVC) return &kHevc;
  return nullptr;
}
const AVCodec* avcodec_find_decoder_by_name(const char* name) {
  if (!std::strcmp(name, "h264")) return &kH264;
  if (!std::strcmp(name, "hevc")) return &kHevc;
  return nullptr;
}
AVCodecContext* avcodec_alloc_context3(const AVCodec*) {
  AVCodecContext* c = new (std::nothrow) AVCodecContext();
  if (c) {
    std::memset(c, 0, sizeof *c);
    c->shim_pending = -1;
  }
  return c;
}
void avcodec_free_context(AVCodecContext** c) {
  if (*c) {
    std::free((*c)->shim_buf);
    delete *c;
  }
  *c = nullptr;
}
int avcodec_parameters_to_context(AVCodecContext* c, const AVCodecParameters* par) {
  c->width = par->width;
  c->height = par->height;
  // the stream is recovered from the parameters block, which lives inside MvsStream
  c->shim_stream = reinterpret_cast<const MvsStream*>(reinterpret_cast<const char*>(par) - offsetof(MvsStream, par));
  return 0;
}
int avcodec_open2(AVCodecContext* c, const AVCodec*, AVDictionary** options) {
  c->shim_export_mvs = (options && *options) ? (*options)->export_mvs : 0;
  return 0;
}
void avcodec_flush_buffers(AVCodecContext* c) { c->shim_pending = -1; }

int avcodec_send_packet(AVCodecContext* c, const AVPacket* pkt) {
  if (c->shim_pending >= 0) return AVERROR(11);  // EAGAIN
  c->shim_pending = pkt->shim_frame;
  return 0;
}

int avcodec_receive_frame(AVCodecContext* c, AVFrame* frame) {
  if (c->shim_pending < 0) return AVERROR(11);
  const MvsStream* s = c->shim_stream;
  const MvsFrameEntry& e = s->frames[(size_t)c->shim_pending];
  c->shim_pending = -1;
  frame->pts = e.pts;
  frame->width = c->width;
  frame->height = c->height;
  frame->key_frame = (e.flags & MVS_FRAME_KEY) ? 1 : 0;
  frame->shim_sd = nullptr;
  // without flags2=+export_mvs no side data exists (motion_scanner.cpp:168-172)
  if (c->shim_export_mvs && (e.flags & MVS_FRAME_HAS_MVS)) {
    const size_t bytes = (size_t)e.n_records * sizeof(AVMotionVector);
    if (bytes > c->shim_buf_cap) {
      std::free(c->shim_buf);
      c->shim_buf = static_cast<uint8_t*>(std::malloc(bytes ? bytes : 1));
      c->shim_buf_cap = bytes;
    }
    if (bytes && !io_read_at(s->io, (int64_t)(s->hdr.records_offset + e.first_record * sizeof(AVMotionVector)),
                             c->shim_buf, bytes))
      return AVERROR(5);  // EIO
    frame->shim_sd_storage.type = AV_FRAME_DATA_MOTION_VECTORS;
    frame->shim_sd_storage.data = c->shim_buf;
    frame->shim_sd_storage.size = bytes;
    frame->shim_sd = &frame->shim_sd_storage;
  }
  return 0;
}

AVFormatContext* avformat_alloc_context(void) {
  AVFormatContext* f = new (std::nothrow) AVFormatContext();
  if (f) {
    std::memset(f, 0, sizeof *f);
    f->duration = AV_NOPTS_VALUE;
  }
  return f;
}

int avformat_open_input(AVFormatContext** ps, const char*, const void*, AVDictionary**) {
  AVFormatContext* f = *ps;
  auto bail = [&]() {
    delete f->shim_stream;  // like libavformat: a failed open frees the context, not the caller's pb
    delete f;
    *ps = nullptr;
    return -1;
  };
  if (!f || !f->pb) return -1;
  MvsStream* s = new (std::nothrow) MvsStream();
  if (!s) return bail();
  f->shim_stream = s;
  s->io = f->pb;
  if (!io_read_at(s->io, 0, &s->hdr, sizeof s->hdr) || std::memcmp(s->hdr.magic, MVS_MAGIC, 8) != 0) return bail();
  s->frames.resize(s->hdr.n_frames);
  if (s->hdr.n_frames &&
      !io_read_at(s->io, sizeof(MvsHeader), s->frames.data(), sizeof(MvsFrameEntry) * (size_t)s->hdr.n_frames))
    return bail();
  s->par.codec_type = AVMEDIA_TYPE_VIDEO;
  s->par.codec_id = AV_CODEC_ID_H264;
  s->par.width = s->hdr.width;
  s->par.height = s->hdr.height;
  s->stream.index = 0;
  s->stream.time_base = AVRational{s->hdr.tb_num, s->hdr.tb_den};
  s->stream.avg_frame_rate = AVRational{s->hdr.fps_num, s->hdr.fps_den};
  s->stream.codecpar = &s->par;
  s->stream.discard = AVDISCARD_DEFAULT;
  s->stream_ptr = &s->stream;
  f->nb_streams = 1;
  f->streams = &s->stream_ptr;
  f->duration = s->hdr.duration_us;
  f->shim_next = 0;
  return 0;
}

int avformat_find_stream_info(AVFormatContext*, AVDictionary**) { return 0; }

void avformat_close_input(AVFormatContext** ps) {
  AVFormatContext* f = *ps;
  if (!f) return;
  if (f->pb) avio_context_free(&f->pb);  // custom IO: the reference relies on this freeing its buffer
  delete f->shim_stream;
  delete f;
  *ps = nullptr;
}

int av_find_best_stream(AVFormatContext* f, enum AVMediaType type, int, int, const AVCodec**, int) {
  return (type == AVMEDIA_TYPE_VIDEO && f->nb_streams) ? 0 : -1;
}

int av_read_frame(AVFormatContext* f, AVPacket* pkt) {
  MvsStream* s = f->shim_stream;
  if (f->shim_next >= (int)s->frames.size()) return AVERROR_EOF;
  pkt->stream_index = 0;
  pkt->shim_frame = f->shim_next;
  pkt->pts = s->frames[(size_t)f->shim_next].pts;
  ++f->shim_next;
  return 0;
}

int av_seek_frame(AVFormatContext* f, int, int64_t ts, int flags) {
  MvsStream* s = f->shim_stream;
  int best = -1;
  for (int i = 0; i < (int)s->frames.size(); ++i) {
    const MvsFrameEntry& e = s->frames[(size_t)i];
    if (!(e.flags & MVS_FRAME_KEY)) continue;
    if (e.pts <= ts) best = i;
    else if (!(flags & AVSEEK_FLAG_BACKWARD) && best < 0) {
      best = i;
      break;
    } else break;
  }
  if (best < 0) best = 0;
  f->shim_next = best;
  return 0;
}

}  // extern "C"
