// main.cpp — motion_trim_b200 <input> <output>: same two positionals and the same directory/file switch
// as the reference CLI (src/main.cpp:35-100). A directory ⇒ batch over the GPUs of the box, a file ⇒
// single pipeline on GPU 0. Inputs are MVS1 motion-vector stream files (.mvs) — what FFmpeg's
// export_mvs decode of the clip would deliver — and, in -DMT_WITH_FFMPEG builds, the media files the
// reference accepts (src/main.cpp:68-69), decoded by the FFmpeg front-end.
// `--print-segments` additionally prints the decision and the job's segments as hex doubles.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <string>
#include <vector>

#include "motion_trim/batch_processor.hpp"
#include "motion_trim/config.hpp"
#include "motion_trim/gpu_pool.hpp"
#include "motion_trim/pipeline.hpp"

namespace fs = std::filesystem;
using namespace motion_trim;

static bool accepted(const std::string& ext) {
  if (ext == ".mvs") return true;
#ifdef MT_WITH_FFMPEG
  for (const char* m : {".mp4", ".mkv", ".ts", ".mov", ".avi"})  // src/main.cpp:68-69
    if (ext == m) return true;
#endif
  return false;
}

int main(int argc, char** argv) {
  std::setvbuf(stdout, nullptr, _IONBF, 0);
  bool print_segments = false;
  std::vector<std::string> pos;
  for (int i = 1; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--print-segments")) print_segments = true;
    else pos.push_back(argv[i]);
  }
  if (pos.size() < 2) {
    std::printf("Usage: %s [--print-segments] <input> <output>\n  <input>: .mvs file or directory of .mvs files\n", argv[0]);
    return 1;
  }
  const std::string input = pos[0], output = pos[1];
  if (!Config::validate()) return 2;  // an unparsable knob is fatal (the reference's std::stod/stoi terminate)
  if (fs::is_directory(input)) {
    std::vector<std::string> files;
    for (const auto& e : fs::directory_iterator(input))
      if (e.is_regular_file() && accepted(e.path().extension().string())) files.push_back(e.path().string());
    std::sort(files.begin(), files.end());
    if (!fs::exists(output)) fs::create_directories(output);
    BatchProcessor batch(Config::parallel_streams());
    batch.set_input_filter(accepted);
    const int failed = batch.process(files, output, input);
    if (print_segments)
      for (const auto& r : batch.results())
        std::printf("RESULT %s gpu=%d rc=%d decision=%d segments=%zu saved_pct=%a\n", fs::path(r.file).filename().c_str(), r.gpu,
                    r.rc, r.decision, r.n_segments, r.saved_pct);
    return failed > 0 ? 1 : 0;
  }
  GpuPool pool;
  if (!pool.open(std::max(1, Config::split_gpus()))) {
    std::printf("[ERROR] %s\n", pool.error().c_str());
    return 1;
  }
  ProcessingPipeline p(input, output, -1, Config::threads_per_stream());
  std::vector<int> gpus;
  for (int g = 0; g < pool.size(); ++g) gpus.push_back(g);  // > 1 only with MOTION_TRIM_SPLIT_GPUS
  p.set_gpus(&pool, gpus);
  const int rc = p.run();
  if (print_segments) {
    std::printf("RESULT decision=%d duration=%a time_removed=%a saved_pct=%a segments=%zu\n", p.get_decision(), p.get_duration(),
                p.get_time_removed(), p.get_saved_pct(), p.get_segments().size());
    for (const auto& s : p.get_segments()) std::printf("SEGMENT %a %a\n", s.start, s.end);
  }
  return rc;
}
