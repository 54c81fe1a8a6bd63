# Builds oracle/_ref/ref_scan: the reference's OWN sources, compiled where they lie under
# /root/reference (never copied), against the fake-libav shim in oracle/ffshim. Outputs only into
# oracle/_ref/ (git-ignored, travels to the GPU box). The reference's build system (CMake + FFmpeg +
# fmt via pkg-config, CMakeLists.txt:105-113) is not used; fmt comes header-only from torch's tree.
REF      ?= /root/reference
CXX      ?= g++
FMT_INC  := $(shell python -c "import torch,os;print(os.path.join(os.path.dirname(torch.__file__),'include'))")
# reference Release flags are -O3 -march=native (CMakeLists.txt:61-68); x86-64-v3 because the
# binary travels to a different CPU; no -flto (this toolchain has no lto-wrapper). Logging compiled
# out; the per-frame analyze timer stays.
CXXFLAGS ?= -std=c++20 -O3 -march=x86-64-v3 -pthread -DFMT_HEADER_ONLY -DENABLE_LOGGING=0
INCS     := -I ffshim -I ../include -I $(REF)/include -I $(FMT_INC)
REF_SRCS := $(REF)/src/motion_scanner.cpp $(REF)/src/pipeline.cpp $(REF)/src/memory_io.cpp \
            $(REF)/src/task_queue.cpp $(REF)/src/ffmpeg_queue.cpp $(REF)/src/logging.cpp $(REF)/src/system.cpp

all: _ref/ref_scan

_ref/ref_scan: ref_harness.cpp ffshim/fake_libav.cpp ffshim/ffshim.h ../include/mvs_format.h $(REF_SRCS)
	mkdir -p _ref
	$(CXX) $(CXXFLAGS) $(INCS) -o $@ ref_harness.cpp ffshim/fake_libav.cpp $(REF_SRCS)

clean:
	rm -rf _ref

.PHONY: all clean
