// memory_io.cpp — mmap of the input and the zero-copy MVS1 view (see memory_io.hpp).
#include "motion_trim/memory_io.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>
#include <utility>

#include "mvs_format.h"

namespace motion_trim {

static_assert(sizeof(MvsFrame) == sizeof(MvsFrameEntry), "MvsFrame mirrors MvsFrameEntry");

void MappedFile::reset() {
  if (data_) munmap(data_, size_);
  if (fd_ != -1) close(fd_);
  data_ = nullptr;
  size_ = 0;
  fd_ = -1;
}

MappedFile::~MappedFile() { reset(); }

void MappedFile::will_need() const {
  if (data_) madvise(data_, size_, MADV_WILLNEED);
}

MappedFile::MappedFile(MappedFile&& o) noexcept : data_(o.data_), size_(o.size_), fd_(o.fd_) {
  o.data_ = nullptr;
  o.size_ = 0;
  o.fd_ = -1;
}

MappedFile& MappedFile::operator=(MappedFile&& o) noexcept {
  if (this != &o) {
    reset();
    std::swap(data_, o.data_);
    std::swap(size_, o.size_);
    std::swap(fd_, o.fd_);
  }
  return *this;
}

bool MemoryLoader::load_file(const std::string& path, MappedFile& file) {
  const int fd = open(path.c_str(), O_RDONLY);
  if (fd < 0) return false;
  struct stat sb;
  if (fstat(fd, &sb) != 0 || sb.st_size <= 0) {
    close(fd);
    return false;
  }
  // Lazy mapping: the reference populates it up front (MAP_POPULATE, src/memory_io.cpp:103-109) because one
  // FFmpeg thread then reads the file sequentially. MV-stream inputs are read once, by the many threads of the
  // projection pool, which fault pages in faster in parallel (16 × 720 MB clips: 0.75 s → 0.55 s per batch,
  // profiles/r02_batch_cli.log); decode-fed runs ask for read-ahead instead (MappedFile::will_need).
  const char* pop = std::getenv("MOTION_TRIM_POPULATE");
  const int flags = MAP_PRIVATE | ((pop && pop[0] == '1') ? MAP_POPULATE : 0);
  void* addr = mmap(nullptr, (size_t)sb.st_size, PROT_READ, flags, fd, 0);
  if (addr == MAP_FAILED) {
    close(fd);
    return false;
  }
  madvise(addr, (size_t)sb.st_size, MADV_SEQUENTIAL);
  file.reset();
  file.data_ = static_cast<uint8_t*>(addr);
  file.size_ = (size_t)sb.st_size;
  file.fd_ = fd;
  return true;
}

bool MvsView::parse(const MappedFile& file) {
  if (!file.is_valid() || file.size() < sizeof(MvsHeader)) return false;
  MvsHeader h;
  std::memcpy(&h, file.data(), sizeof h);
  if (std::memcmp(h.magic, MVS_MAGIC, 8) != 0) return false;
  // the header is untrusted (watch mode picks files up on its own): every bound is checked without arithmetic that
  // could wrap — a crafted n_records / first_record must not let a submit read or DMA outside the mapping
  const uint64_t fsize = file.size();
  if ((uint64_t)h.n_frames > (fsize - sizeof(MvsHeader)) / sizeof(MvsFrameEntry)) return false;
  const uint64_t table_end = sizeof(MvsHeader) + (uint64_t)h.n_frames * sizeof(MvsFrameEntry);
  if (h.records_offset < table_end || h.records_offset > fsize || (h.records_offset & 7u)) return false;
  if (h.n_records > (fsize - h.records_offset) / sizeof(mscan_mv)) return false;
  if (h.tb_den == 0 || h.fps_den == 0) return false;
  width = h.width;
  height = h.height;
  tb_num = h.tb_num;
  tb_den = h.tb_den;
  fps_num = h.fps_num;
  fps_den = h.fps_den;
  duration_us = h.duration_us;
  n_frames = h.n_frames;
  n_records = h.n_records;
  frames = reinterpret_cast<const MvsFrame*>(file.data() + sizeof(MvsHeader));
  records = reinterpret_cast<const mscan_mv*>(file.data() + h.records_offset);
  for (uint32_t i = 0; i < n_frames; ++i)  // every frame's records must lie inside the file
    if (frames[i].first_record > n_records || frames[i].n_records > n_records - frames[i].first_record) return false;
  return true;
}

}  // namespace motion_trim
