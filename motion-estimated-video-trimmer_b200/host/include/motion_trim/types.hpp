// types.hpp — result/value types shared with the reference (include/motion_trim/types.hpp:56-59,92-96).
#pragma once

#include <string>
#include <vector>

#include "motionscan.h"

namespace motion_trim {

// Same layout as the reference's TimeSegment and as mscan_segment (two doubles, 16-byte aligned).
struct alignas(16) TimeSegment {
  double start;
  double end;
};
static_assert(sizeof(TimeSegment) == sizeof(mscan_segment), "TimeSegment must match mscan_segment");

struct ScanTask {
  double start;
  double end;
  int id;
};

// What the scan hands to the muxing stage (reference include/motion_trim/ffmpeg_queue.hpp:32-38).
struct FFmpegJob {
  int stream_id = -1;
  std::string input_path;
  std::string output_path;
  std::vector<TimeSegment> segments;
  std::vector<int> cpu_set;
};

}  // namespace motion_trim
