// memory_io.hpp — input side of the host mirror. In the reference this module mmaps the media file
// and feeds FFmpeg through AVIO callbacks (src/memory_io.cpp:73-166). This image has no FFmpeg, so
// the input is an MVS1 motion-vector stream file (include/mvs_format.h): the frames' pts and
// AVMotionVector records exactly as export_mvs would deliver them. MappedFile keeps the reference's
// RAII mmap role; MvsView is the parsed, zero-copy view the scanner iterates.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>

#include "motionscan.h"

namespace motion_trim {

class MappedFile {
 public:
  MappedFile() = default;
  ~MappedFile();
  MappedFile(const MappedFile&) = delete;
  MappedFile& operator=(const MappedFile&) = delete;
  MappedFile(MappedFile&& o) noexcept;
  MappedFile& operator=(MappedFile&& o) noexcept;

  const uint8_t* data() const { return data_; }
  size_t size() const { return size_; }
  bool is_valid() const { return data_ != nullptr; }
  void will_need() const;  // asynchronous read-ahead of the whole mapping (decode-fed inputs)

 private:
  friend class MemoryLoader;
  void reset();
  uint8_t* data_ = nullptr;
  size_t size_ = 0;
  int fd_ = -1;
};

class MemoryLoader {
 public:
  // mmap(PROT_READ, MAP_PRIVATE) + madvise(SEQUENTIAL); the reference's MAP_POPULATE (src/memory_io.cpp:103-115) only with MOTION_TRIM_POPULATE=1
  static bool load_file(const std::string& path, MappedFile& file);
};

struct MvsFrame {
  int64_t pts;
  uint64_t first_record;
  uint32_t n_records;
  uint32_t flags;  // bit0 key frame, bit1 carries motion vectors
};

struct MvsView {
  int width = 0, height = 0;
  int tb_num = 1, tb_den = 1;
  int fps_num = 0, fps_den = 1;
  int64_t duration_us = 0;
  uint32_t n_frames = 0;
  const MvsFrame* frames = nullptr;
  const mscan_mv* records = nullptr;
  uint64_t n_records = 0;

  // false if the mapping is not a well-formed MVS1 file
  bool parse(const MappedFile& file);
};

}  // namespace motion_trim
