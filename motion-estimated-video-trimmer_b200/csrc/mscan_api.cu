// mscan_api.cu — C ABI of libmotionscan.so (see include/motionscan.h for the reference citations).
//
// Host-side structure, per GPU context:
//   * frame log   — device arrays pts/flags/counts indexed by "every frame ever submitted, in
//                   order"; a video is a list of extents of the log (ResultCollector role,
//                   reference src/task_queue.cpp:43-57);
//   * slab ring   — 3 device record slabs, each with pinned host staging and its own stream:
//                   while slab k is being DMA'd and scanned, submit() fills slab k+1
//                   (memory_io.cpp role: the reference mmaps the file for FFmpeg; here the staged
//                   thing is the decoder's MV side data on its way to HBM);
//   * submit      — reserve / fill / commit: a call reserves its byte range, frame slots and log range
//                   under the context mutex (a few hundred ns), fills the reserved range of the pinned
//                   staging OUTSIDE the mutex, and commits with an atomic. Many decode threads therefore
//                   project their own cache-hot side data concurrently (one scanner per decode thread,
//                   include/motion_trim/motion_scanner.hpp:8-13, workers at src/pipeline.cpp:186-235);
//                   staged bytes go to the GPU in copy windows as soon as every writer of a window
//                   has committed, so the link works while the slab is still filling;
//   * projection  — native records in pageable memory are not memcpy'd into staging: only bytes 6..13
//                   of each 40-byte record (the four int16 the path reads) are written there (by the
//                   process-wide worker pool for large submits), so 8 B/record cross PCIe instead of 40;
//   * tails       — collect / segments / close wait only for the slabs that hold frames of the videos
//                   in the call, outside the mutex: one stream's tail does not stall the others;
//   * K-A per slab segment, K-C per segments call.
// There is no CPU implementation of the path behind this ABI.
#include <cuda_runtime.h>

#include <sched.h>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <shared_mutex>
#include <string>
#include <thread>
#include <vector>

#include "kernels.cuh"

using namespace mscan;

namespace {

constexpr int kSlabs = 3;
constexpr size_t kPendStride = 16;  // copy-window counters sit one per cache line (16 x int32)
constexpr int kWorkSlots = 16;  // frame queues: one per slab stream (launches on a stream are serialised) + a rotating
                                 // pool for launches on caller streams (mscan_scan_device)
constexpr uint32_t kMaxGeoms = 4096;

struct Extent {
  uint64_t start, n;  // frames [start, start+n) of the frame log
  uint64_t vpos;      // index of the extent's first frame in the video's submission order (a submit call reserves
                      // its index range up front, so concurrent submits to one video keep each call's frames together)
};

struct Video {
  // (what every submit writes comes first: it shares cache lines that travel between the decode threads' cores)
  uint32_t geom = 0;
  uint64_t n_frames = 0;
  // epoch of each slab when this video last put frames into it: equal to the slab's current epoch ⇔ results
  // of this video may still be pending there (what collect / segments / close have to wait for — nothing else)
  uint64_t slab_epoch[kSlabs] = {0, 0, 0};
  // host-pass counters of this video's compacted submits, kept here because the submit's critical section writes this
  // node anyway; folded into the context's stats by fold_stats / mscan_video_close
  uint64_t st_records = 0, st_wire_bytes = 0, st_ns = 0;
  std::vector<Extent> extents;
  uint64_t dev_seq = 0;  // mscan_submit_device: sequence number of this video's last launch on the context's dev_stream
};

struct Slab {
  // ---- set up once (read-mostly) ----
  uint8_t* d_recs = nullptr;
  uint8_t* h_recs = nullptr;  // pinned, allocated on first pageable submit
  uint64_t* d_rec_off = nullptr;
  uint64_t* h_rec_off = nullptr;
  uint32_t* d_geom = nullptr;
  uint32_t* h_geom = nullptr;
  double* h_pts = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  cudaEvent_t copied = nullptr;  // recorded after the slab's last H2D copy (mscan_host_fence)
  // mvz segments: tile directory (start of every tile in 16-byte units from the segment base, + the end of the last)
  // and each frame's first tile; a segment of k tiles uses k + 1 directory slots
  uint32_t *d_tile_dir = nullptr, *h_tile_dir = nullptr, *d_frame_tile0 = nullptr, *h_frame_tile0 = nullptr;
  std::unique_ptr<std::atomic<int32_t>[]> pend;  // per copy window: uncommitted reservations that touch it
  // ---- what every reservation reads and writes: one cache line (it travels between the decode threads' cores) ----
  // The open segment = frames [seg_frame0, frames): one record format, one K-A launch. A slab carries any
  // number of segments (chunk workers may feed different formats); they are launched in order on `stream`.
  alignas(64) uint64_t bytes = 0;        // fill level of d_recs (and of h_recs, which mirrors its offsets)
  uint64_t seg_recs = 0;
  std::atomic<uint64_t> reserved_end{0};  // == bytes, published for the copy pump
  uint64_t epoch = 1;                     // bumped by every recycle
  uint32_t frames = 0;                    // frames staged since the slab was recycled
  uint32_t seg_frame0 = 0;
  uint32_t seg_slot0 = 0;     // first rec_off slot of the open segment (a segment of k frames uses k+1 slots)
  uint32_t dir_slots = 0;     // directory slots used since the slab was recycled
  uint32_t seg_dir0 = 0;      // first directory slot of the open segment
  uint8_t fmt = 0;            // record layout of the open segment: kLayoutNative (40 B), kLayoutMv8 (8 B) or kLayoutMvz
  bool staged = false;        // the open segment's records pass through h_recs (else DMA'd from caller memory)
  bool in_flight = false;
  // ---- per segment / per launch ----
  alignas(64) uint64_t seg_byte0 = 0;  // 256-byte aligned offset of the open segment's first record
  uint64_t seg_log_base = 0;           // frame-log index of the open segment's first frame
  uint64_t launch_seq = 0;             // bumped by every K-A launch on this slab
  uint64_t copy_head = 0;              // (ctx->issue_mu) staged bytes of the open segment below this are on their way
  std::atomic<uint32_t> writers{0};    // in-place reservations of the open segment whose copy has not been enqueued
};
static_assert(offsetof(Slab, in_flight) / 64 == offsetof(Slab, bytes) / 64, "the reservation fields of a slab share one cache line");

struct EvPair {
  cudaEvent_t a, b;
  int kind;  // 0 = K-A, 1 = K-C
};

// ---- host projection AVMotionVector → mscan_mv8: mscan::project_records, csrc/host_project.cpp ----------------

constexpr uint64_t kPoolChunk = 32768;      // records per work item (1.3 MB of native records)
constexpr uint64_t kPoolMinRecs = 1 << 18;  // smaller jobs are projected by the calling thread alone

// The CPUs the PROCESS may run on, captured the first time the library is used (mscan_create, normally from the main
// thread): decode threads are often pinned to a few CPUs each (src/system.cpp:211-225), and the shared projection pool
// must not inherit the mask of whichever of them happens to start it.
const cpu_set_t& process_cpus() {
  static const cpu_set_t mask = [] {
    cpu_set_t m;
    CPU_ZERO(&m);
    if (sched_getaffinity(0, sizeof m, &m) != 0 || CPU_COUNT(&m) == 0)
      for (int i = 0; i < (int)std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), CPU_SETSIZE); ++i) CPU_SET(i, &m);
    return m;
  }();
  return mask;
}

// Parallel-for over one projection job; the calling thread takes part. ONE pool per process, shared by every
// context (a pool per GPU context sized for the whole box would oversubscribe it n-fold, one sized cores/n would
// leave cores idle whenever the GPUs are not all projecting): jobs are serialised, each may use up to `limit` threads.
class PackPool {
 public:
  // Starts up to `workers` threads; if the platform refuses some, the pool simply runs with fewer.
  explicit PackPool(int workers) noexcept {
    try {
      for (int i = 0; i < workers; ++i) th_.emplace_back([this, i] { worker(i); });
    } catch (...) {
    }
  }
  ~PackPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  int workers() const { return (int)th_.size(); }

  // `limit`: threads that may work on this job, the caller included (<= 0: all).
  void run(const uint8_t* src, uint64_t* dst, uint64_t n, int limit) {
    std::lock_guard<std::mutex> job(job_mu_);
    const uint64_t total = (n + kPoolChunk - 1) / kPoolChunk;
    {
      std::unique_lock<std::mutex> lk(mu_);
      idle_cv_.wait(lk, [this] { return active_ == 0; });  // stragglers of the previous job have left
      src_ = src;
      dst_ = dst;
      n_ = n;
      total_ = total;
      limit_ = limit <= 0 ? (int)th_.size() : std::min((int)th_.size(), limit - 1);
      next_.store(0, std::memory_order_relaxed);
      done_.store(0, std::memory_order_relaxed);
      ++gen_;
    }
    cv_.notify_all();
    drain(src, dst, n, total);
    std::unique_lock<std::mutex> lk(mu_);
    idle_cv_.wait(lk, [&] { return done_.load(std::memory_order_acquire) == total && active_ == 0; });
  }

 private:
  void drain(const uint8_t* src, uint64_t* dst, uint64_t n, uint64_t total) {
    for (uint64_t k = next_.fetch_add(1, std::memory_order_relaxed); k < total; k = next_.fetch_add(1, std::memory_order_relaxed)) {
      const uint64_t a = k * kPoolChunk, b = std::min(n, a + kPoolChunk);
      project_records(src + (size_t)kRecBytes * a, b - a, dst + a);
      done_.fetch_add(1, std::memory_order_release);
    }
  }
  void worker(int index) {
    sched_setaffinity(0, sizeof(cpu_set_t), &process_cpus());  // not the (possibly pinned) creator's mask
    uint64_t seen = 0;
    for (;;) {
      const uint8_t* src;
      uint64_t *dst, n, total;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_;
        if (index >= limit_) continue;  // this job runs with fewer threads
        src = src_;
        dst = dst_;
        n = n_;
        total = total_;
        ++active_;  // joined under the lock: run() will not post the next job until we have left
      }
      drain(src, dst, n, total);
      {
        std::lock_guard<std::mutex> lk(mu_);
        --active_;
      }
      idle_cv_.notify_all();
    }
  }

  std::vector<std::thread> th_;
  std::mutex mu_, job_mu_;
  std::condition_variable cv_, idle_cv_;
  bool stop_ = false;
  uint64_t gen_ = 0;
  int active_ = 0, limit_ = 0;
  const uint8_t* src_ = nullptr;
  uint64_t* dst_ = nullptr;
  uint64_t n_ = 0, total_ = 0;
  std::atomic<uint64_t> next_{0}, done_{0};
};

int default_pack_threads() {
  if (const char* e = std::getenv("MSCAN_PACK_THREADS")) return std::max(1, std::atoi(e));
  const int n = CPU_COUNT(&process_cpus());  // honours taskset / cgroup cpusets
  return std::max(1, std::min(n, 64));
}

// The process-wide pool, created at the first large projection (never destroyed: worker threads must not be
// joined from a static destructor of a dlopen'ed library).
PackPool* shared_pool() {
  static PackPool* pool = new (std::nothrow) PackPool(default_pack_threads() - 1);
  return pool;
}

// Host ranges this library pinned (mscan_host_alloc / mscan_host_register): mscan_submit looks a source pointer up
// here instead of asking the driver on every call (cudaPointerGetAttributes costs ~1 µs, a per-frame submit from a
// decode thread is worth less than that). Memory pinned behind the library's back (cudaHostRegister by the caller)
// is still recognised for submits of at least kAttrQueryBytes, where the driver query is noise.
constexpr uint64_t kAttrQueryBytes = 1ull << 20;
class PinnedRanges {
 public:
  void add(const void* p, size_t n) {
    std::unique_lock<std::shared_mutex> lk(mu_);
    r_[reinterpret_cast<uintptr_t>(p)] = n;
    gen_.fetch_add(1, std::memory_order_release);
  }
  void remove(const void* p) {
    std::unique_lock<std::shared_mutex> lk(mu_);
    r_.erase(reinterpret_cast<uintptr_t>(p));
    gen_.fetch_add(1, std::memory_order_release);
  }
  // Is [p, p+n) inside one registered range? Each thread remembers the last interval it asked about (a registered
  // range, or the gap between two of them) together with the registry's generation, so a decode thread that submits
  // from the same side-data buffer frame after frame never touches the shared lock.
  bool contains(const void* p, size_t n) {
    struct Last {
      uint64_t gen = ~0ull;
      uintptr_t lo = 0, hi = 0;
      bool pinned = false;
    };
    static thread_local Last last;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint64_t g = gen_.load(std::memory_order_acquire);
    if (last.gen == g && a >= last.lo && a + n <= last.hi) return last.pinned;
    std::shared_lock<std::shared_mutex> lk(mu_);
    Last now;
    now.gen = gen_.load(std::memory_order_acquire);
    auto it = r_.upper_bound(a);  // first range starting above a
    now.hi = it == r_.end() ? ~uintptr_t(0) : it->first;
    if (it != r_.begin()) {
      --it;
      if (a < it->first + it->second) {  // a lies inside this range
        now.lo = it->first;
        now.hi = it->first + it->second;
        now.pinned = true;
      } else {
        now.lo = it->first + it->second;
      }
    }
    const bool inside = a + n <= now.hi;
    if (inside || !now.pinned) last = now;  // (a request straddling the end of a range is simply not cached)
    return now.pinned && inside;
  }

 private:
  std::shared_mutex mu_;
  std::map<uintptr_t, size_t> r_;
  std::atomic<uint64_t> gen_{0};
};
PinnedRanges& pinned_ranges() {
  static PinnedRanges* r = new PinnedRanges();
  return *r;
}

}  // namespace

struct mscan_ctx {
  int device = 0;
  int num_sms = 0;
  uint32_t smem_optin = 0;
  mscan_params params{};
  // Lock order: tail_mu → mu → issue_mu. `mu` guards the bookkeeping (videos, frame log, slab fill levels) and is
  // never held while records are projected or copied; `issue_mu` orders the H2D copies of staged windows with the
  // K-A launch that consumes them; `tail_mu` serialises K-C (its scratch and main_stream) without blocking submits.
  // `mu` shares its cache line with what every reservation updates under it (taking the lock brings them along)
  alignas(64) std::mutex mu;
  int cur = 0;                           // slab of the ring that is being filled
  uint64_t log_head = 0, log_limit = 0;  // frames [log_head, log_limit) of the log are known to be free (space closed videos gave back is found on demand)
  alignas(64) std::mutex issue_mu;
  std::mutex tail_mu;
  std::string err;

  // kernel constants derived from params
  int32_t ithr = 0;
  int32_t keep_none = 0;
  uint32_t vec_need = 0;
  uint32_t clust_need = 1;

  // geometry table
  std::vector<DevGeom> geoms;
  DevGeom* d_geoms = nullptr;
  uint32_t max_cells = 0, max_bit_words = 0;
  ScanPlan plan{};         // native 40-byte slabs
  ScanPlan plan_packed{};  // mscan_mv8 slabs

  // frame log
  uint64_t log_cap = 0;
  double* d_pts = nullptr;
  uint8_t* d_flags = nullptr;
  uint32_t* d_counts = nullptr;

  // slabs
  uint64_t slab_bytes = 0;
  uint32_t slab_frames = 0;
  uint32_t slab_dir_cap = 0;  // tile-directory slots per slab
  Slab slabs[kSlabs];

  uint32_t* d_work = nullptr;  // kWorkSlots × {next, done}
  uint32_t work_rr = 0;
  // a rotating slot is reused only after the launch that last used it: the next user's stream waits for this event
  cudaEvent_t slot_ev[kWorkSlots] = {};
  bool slot_used[kWorkSlots] = {};
  // mscan_submit_device: launches on records that already lie in device memory
  cudaStream_t dev_stream = nullptr;
  cudaEvent_t dev_done = nullptr, dev_ready = nullptr;
  uint64_t dev_seq = 0, dev_done_seq = 0;  // launches issued on dev_stream / known complete
  uint32_t* d_cnt_scratch = nullptr;  // global vote counters for grids beyond shared memory (zeroed)
  uint64_t cnt_scratch_elems = 0;
  uint32_t adj8 = 0;

  cudaStream_t main_stream = nullptr;
  std::map<uint32_t, Video> videos;

  // K-C scratch (grown on demand)
  uint64_t ts_cap = 0;
  double *d_ts_a = nullptr, *d_ts_b = nullptr;
  uint64_t seg_cap = 0;
  mscan_segment* d_segs = nullptr;
  uint32_t job_cap = 0, ext_cap = 0;
  SegJob *d_jobs = nullptr, *h_jobs = nullptr;
  SegExtent *d_exts = nullptr, *h_exts = nullptr;
  mscan_video_result *d_res = nullptr, *h_res = nullptr;
  // geometry staging for the device API
  DevGeom* d_user_geoms = nullptr;
  uint32_t user_geoms_cap = 0;
  std::vector<DevGeom> user_geoms_cached;  // what d_user_geoms currently holds
  // last job table uploaded by mscan_segments_device (re-used when unchanged: no sync, no copy)
  std::vector<SegJob> dev_jobs_cached;

  // host projection (see project_records)
  std::atomic<int> staging_mode{MSCAN_STAGING_AUTO};  // (read before the mutex is taken by the per-frame fast path)
  int pack_threads = 0;  // threads of the shared pool one large submit of this context may use; 0 → all
  uint32_t win_shift = 21;  // log2 of the H2D copy window (2 MiB: 38 µs on a Gen5 x16 link against ~3 µs to enqueue a copy)
  // counters written outside `mu` (folded into `stats` by mscan_get_stats)
  std::atomic<uint64_t> a_h2d_bytes{0}, a_records_projected{0}, a_project_ns{0}, a_records_elided{0}, a_elided_bytes{0};

  // MSCAN_TRACE=1: wall time per ABI entry point (including time spent waiting for the context mutex),
  // printed to stderr by mscan_destroy — the role of the reference's TIMER_START/END + TimingCollector
  // (include/motion_trim/logging.hpp:137-148) for the calls that replaced its analyze phase.
  struct ApiTime {
    uint64_t calls = 0;
    double total_ms = 0, max_ms = 0;
  };
  bool trace = false;
  std::mutex trace_mu;
  std::map<std::string, ApiTime> api_times;

  mscan_stats stats{};
  bool profiling = false;
  std::vector<EvPair> ev_pending;
  std::vector<EvPair> ev_free;
};

namespace {

struct ApiTimer {
  mscan_ctx* c;
  const char* name;
  std::chrono::steady_clock::time_point t0;
  ApiTimer(mscan_ctx* ctx, const char* n) : c(ctx && ctx->trace ? ctx : nullptr), name(n) {
    if (c) t0 = std::chrono::steady_clock::now();
  }
  ~ApiTimer() {
    if (!c) return;
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::lock_guard<std::mutex> lk(c->trace_mu);
    auto& t = c->api_times[name];
    t.calls += 1;
    t.total_ms += ms;
    t.max_ms = std::max(t.max_ms, ms);
  }
};

int fail(mscan_ctx* c, int code, const char* fmt, ...) {
  if (c) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    std::lock_guard<std::mutex> lk(c->trace_mu);  // errors are reported from under different locks
    c->err = buf;
  }
  return code;
}

// folds the counters that are written outside `mu` into the public stats (caller holds mu)
void fold_video_stats(mscan_ctx* c, Video& v) {
  c->stats.records_projected += v.st_records;
  c->stats.records_elided += v.st_records;
  c->stats.elided_bytes += v.st_wire_bytes;
  c->stats.project_ms += (double)v.st_ns * 1e-6;
  v.st_records = v.st_wire_bytes = v.st_ns = 0;
}

void fold_stats(mscan_ctx* c) {
  for (auto& kv : c->videos) fold_video_stats(c, kv.second);
  c->stats.h2d_bytes += c->a_h2d_bytes.exchange(0, std::memory_order_relaxed);
  c->stats.records_projected += c->a_records_projected.exchange(0, std::memory_order_relaxed);
  c->stats.project_ms += (double)c->a_project_ns.exchange(0, std::memory_order_relaxed) * 1e-6;
  c->stats.records_elided += c->a_records_elided.exchange(0, std::memory_order_relaxed);
  c->stats.elided_bytes += c->a_elided_bytes.exchange(0, std::memory_order_relaxed);
}

// No exception crosses the ABI: every entry point that can allocate is a function-try-block ending here.
int on_exception(mscan_ctx* c) noexcept {
  try {
    throw;
  } catch (const std::bad_alloc&) {
    return c ? fail(c, MSCAN_ERR_NOMEM, "out of host memory") : MSCAN_ERR_NOMEM;
  } catch (const std::exception& e) {
    return c ? fail(c, MSCAN_ERR_INVALID, "internal error: %s", e.what()) : MSCAN_ERR_INVALID;
  } catch (...) {
    return c ? fail(c, MSCAN_ERR_INVALID, "internal error") : MSCAN_ERR_INVALID;
  }
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(c, MSCAN_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// keep iff !((double)mag_sq < T²)  (motion_scanner.cpp:251)  ⇔  mag_sq >= ceil(T²) for finite T²
void threshold_to_int(double t2, int32_t* ithr, int32_t* keep_none) {
  *keep_none = 0;
  if (std::isnan(t2)) {  // every `<` against NaN is false → everything is kept
    *ithr = std::numeric_limits<int32_t>::min();
    return;
  }
  const double c = std::ceil(t2);
  if (c > 2147483647.0) {  // also +inf: no int32 magnitude reaches it
    *ithr = std::numeric_limits<int32_t>::max();
    *keep_none = 1;
  } else if (c <= -2147483648.0) {
    *ithr = std::numeric_limits<int32_t>::min();
  } else {
    *ithr = (int32_t)c;
  }
}

DevGeom to_dev_geom(const mscan_geometry& g) {
  DevGeom d;
  d.gw = g.grid_w;
  d.gh = g.grid_h;
  // rows [margin, gh-margin) are live (:237-238); clamp to the grid (margin < 0 or 2*margin > gh is
  // out-of-bounds behaviour in the reference, outside the parity domain)
  int y0 = g.vertical_margin, y1 = g.grid_h - g.vertical_margin;
  y0 = std::max(0, std::min(y0, g.grid_h));
  y1 = std::max(y0, std::min(y1, g.grid_h));
  d.y_min = y0;
  d.y_max = y1;
  return d;
}

bool geom_ok(const mscan_geometry& g) { return g.grid_w >= 0 && g.grid_h >= 0 && g.grid_w <= 32767 && g.grid_h <= 32767; }

void geom_need(const DevGeom& g, uint32_t* cells, uint32_t* bit_words) {
  *cells = (uint32_t)g.gw * (uint32_t)g.gh;
  *bit_words = (uint32_t)g.gh * (((uint32_t)g.gw + 31u) >> 5);
}

EvPair get_events(mscan_ctx* c, int kind) {
  EvPair p{};
  if (!c->ev_free.empty()) {
    p = c->ev_free.back();
    c->ev_free.pop_back();
  } else {
    cudaEventCreate(&p.a);
    cudaEventCreate(&p.b);
  }
  p.kind = kind;
  return p;
}

void drain_events(mscan_ctx* c) {
  for (auto& p : c->ev_pending) {
    float ms = 0.f;
    if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      if (p.kind == 0) c->stats.scan_ms += ms;
      else c->stats.segment_ms += ms;
    }
    c->ev_free.push_back(p);
  }
  c->ev_pending.clear();
}

int run_scan(mscan_ctx* c, const ScanArgs& args_in, const ScanPlan& plan, cudaStream_t st, uint64_t n_recs, int slab_index = -1) {
  ScanArgs a = args_in;
  // two kernels must never share a frame queue while both run: a slab's launches are ordered by its stream, so
  // the slab index is a safe slot; launches on caller streams rotate through the remaining slots
  // slab_index >= 0: that slab's stream; kSlabs: the context's dev_stream; -1: a caller stream → rotating slots, each
  // guarded by the event of its previous user (two kernels sharing a queue while both run would split the frames)
  uint32_t slot;
  if (slab_index >= 0) {
    slot = (uint32_t)slab_index;
  } else {
    slot = (uint32_t)kSlabs + 1u + (c->work_rr++ % (uint32_t)(kWorkSlots - kSlabs - 1));
    if (c->slot_used[slot]) CU(cudaStreamWaitEvent(st, c->slot_ev[slot], 0));
  }
  a.work = c->d_work + 2 * slot;
  a.adj8 = c->adj8;
  a.cnt_scratch = nullptr;
  if (plan.global_cnt) {
    // one scratch per context: launches that need it are serialised behind each other
    const uint64_t need = (uint64_t)c->num_sms * plan.ctas_per_sm * a.max_cells;
    if (need * sizeof(uint32_t) > (16ull << 30))
      return fail(c, MSCAN_ERR_UNSUPPORTED, "block grid of %u cells needs more than 16 GiB of counter scratch", a.max_cells);
    if (need > c->cnt_scratch_elems) {
      CU(cudaDeviceSynchronize());
      cudaFree(c->d_cnt_scratch);
      c->d_cnt_scratch = nullptr;
      c->cnt_scratch_elems = 0;
      if (cudaMalloc((void**)&c->d_cnt_scratch, need * sizeof(uint32_t)) != cudaSuccess) {
        cudaGetLastError();
        return fail(c, MSCAN_ERR_NOMEM, "cannot allocate %llu bytes of vote-counter scratch", (unsigned long long)(need * 4));
      }
      CU(cudaMemset(c->d_cnt_scratch, 0, need * sizeof(uint32_t)));
      c->cnt_scratch_elems = need;
    } else {
      CU(cudaDeviceSynchronize());
    }
    a.cnt_scratch = c->d_cnt_scratch;
  }
  EvPair ev{};
  if (c->profiling) {
    ev = get_events(c, 0);
    cudaEventRecord(ev.a, st);
  }
  cudaError_t le = scan_launch(a, plan, c->num_sms, st);
  if (le != cudaSuccess && plan.cluster) {
    // the device would not place the cluster (partitioned GPU, co-tenants): scan with per-CTA counters in global memory
    cudaGetLastError();
    ScanPlan fb;
    if (!scan_plan(plan.full_cells, plan.full_bit_words, c->smem_optin, a.packed != 0, &fb) || !fb.global_cnt)
      return fail(c, MSCAN_ERR_CUDA, "cluster launch failed (%s) and the grid has no single-CTA plan", cudaGetErrorString(le));
    fb.cluster = 0;
    fb.cells = plan.full_cells;
    fb.bit_words = plan.full_bit_words;
    ScanArgs b = args_in;
    b.stages = fb.stages;
    b.max_cells = fb.cells;
    b.max_bit_words = fb.bit_words;
    if (c->profiling) c->ev_free.push_back(ev);
    return run_scan(c, b, fb, st, n_recs, slab_index);
  }
  CU(le);
  if (c->profiling) {
    cudaEventRecord(ev.b, st);
    c->ev_pending.push_back(ev);
  }
  if (slab_index < 0) {
    CU(cudaEventRecord(c->slot_ev[slot], st));
    c->slot_used[slot] = true;
  }
  c->stats.scan_launches += 1;
  c->stats.frames_scanned += a.n_frames;
  c->stats.records_scanned += n_recs;
  return MSCAN_OK;
}

ScanArgs base_args(mscan_ctx* c) {
  ScanArgs a{};
  a.ithr = c->ithr;
  a.keep_none = c->keep_none;
  a.shift = c->params.block_shift;
  a.vec_need = c->vec_need;
  a.clust_need = c->clust_need;
  return a;
}

// Takes `mu` for a short critical section: spins briefly before blocking. With one lock acquisition per submitted frame
// from every decode thread, a contended std::mutex::lock() puts the loser to sleep in the kernel, and the wake-up costs far
// more (tens of µs in a VM) than the few hundred ns the holder needs.
void lock_briefly(std::unique_lock<std::mutex>& lk) {
  for (int round = 0, wait = 1; round < 48; ++round) {  // exponential back-off: 30 threads retrying at full rate would
    if (lk.try_lock()) return;                          // keep the lock's cache line away from its holder
    for (int i = 0; i < wait; ++i) {
#if defined(__x86_64__)
      _mm_pause();
#endif
    }
    if (wait < 128) wait *= 2;
  }
  lk.lock();
}

// CUDA calls need the context's device current in the calling thread; the submit fast path (reserve / project / commit)
// makes none, so this is only paid where something is enqueued.
inline cudaError_t use_device(mscan_ctx* c) {
  int d = -1;
  if (cudaGetDevice(&d) == cudaSuccess && d == c->device) return cudaSuccess;
  return cudaSetDevice(c->device);
}

// Copy pump (caller holds issue_mu): enqueues the H2D copy of every staged window of the open segment that is
// complete — closed (the reservation pointer has moved past its end) and without uncommitted writers. Windows go out
// in order, so copy_head is the only state; `seal` also sends the partial last window (launch_segment has already
// waited for the segment's writers). Anyone may pump at any time: the decision depends on the slab's state only.
int pump_locked(mscan_ctx* c, Slab& s, bool seal) {
  if (!s.staged) return MSCAN_OK;
  const uint64_t end_res = s.reserved_end.load(std::memory_order_acquire);
  while (s.copy_head < end_res) {
    const uint64_t k = s.copy_head >> c->win_shift, win_end = (k + 1) << c->win_shift;
    if (!seal && end_res < win_end) break;
    if (s.pend[k * kPendStride].load(std::memory_order_acquire) != 0) break;
    const uint64_t e = std::min(win_end, end_res);
    CU(cudaMemcpyAsync(s.d_recs + s.copy_head, s.h_recs + s.copy_head, e - s.copy_head, cudaMemcpyHostToDevice, s.stream));
    c->a_h2d_bytes.fetch_add(e - s.copy_head, std::memory_order_relaxed);
    s.copy_head = e;
  }
  return MSCAN_OK;
}

void try_pump(mscan_ctx* c, Slab& s) {
  std::unique_lock<std::mutex> lk(c->issue_mu, std::try_to_lock);
  if (!lk.owns_lock()) return;  // somebody is pumping or launching; launch_segment is the backstop
  if (use_device(c) != cudaSuccess) return;
  pump_locked(c, s, false);
}

// Every reservation of the open segment has been committed: the in-place ones are counted in `writers`, the staged
// ones in the copy-window counters of the segment's byte range (no separate count: one contended atomic less per frame).
bool segment_quiet(const mscan_ctx* c, const Slab& s) {
  if (s.writers.load(std::memory_order_acquire) != 0) return false;
  if (s.staged && s.bytes > s.seg_byte0)
    for (uint64_t k = s.seg_byte0 >> c->win_shift; k <= (s.bytes - 1) >> c->win_shift; ++k)
      if (s.pend[k * kPendStride].load(std::memory_order_acquire) != 0) return false;
  return true;
}

void wait_writers(const mscan_ctx* c, const Slab& s) {
  for (uint32_t spin = 0; !segment_quiet(c, s); ++spin) {
#if defined(__x86_64__)
    if (spin < 256) _mm_pause();
    else std::this_thread::yield();
#else
    std::this_thread::yield();
#endif
  }
}

// launch K-A on the open segment of a slab and open the next one (caller holds mu)
int launch_segment(mscan_ctx* c, Slab& s) {
  const uint32_t n = s.frames - s.seg_frame0;
  if (n == 0) return MSCAN_OK;
  wait_writers(c, s);  // fills run outside mu and never need it to commit: this wait is bounded by one frame's projection
  std::lock_guard<std::mutex> issue(c->issue_mu);
  CU(use_device(c));
  int rc = pump_locked(c, s, true);
  if (rc) return rc;
  s.h_rec_off[s.seg_slot0 + n] = s.seg_recs;
  CU(cudaMemcpyAsync(s.d_rec_off + s.seg_slot0, s.h_rec_off + s.seg_slot0, sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, s.stream));
  CU(cudaMemcpyAsync(s.d_geom + s.seg_frame0, s.h_geom + s.seg_frame0, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, s.stream));
  CU(cudaMemcpyAsync(c->d_pts + s.seg_log_base, s.h_pts + s.seg_frame0, sizeof(double) * n, cudaMemcpyHostToDevice, s.stream));
  c->a_h2d_bytes.fetch_add(sizeof(uint64_t) * (n + 1) + 12ull * n, std::memory_order_relaxed);
  CU(cudaEventRecord(s.copied, s.stream));  // every H2D copy of the slab so far precedes this point
  ScanArgs a = base_args(c);
  a.recs = s.d_recs + s.seg_byte0;
  a.packed = s.fmt;
  if (s.fmt == kLayoutMvz) {
    const uint32_t n_dir = s.dir_slots - s.seg_dir0;  // tiles + 1
    CU(cudaMemcpyAsync(s.d_tile_dir + s.seg_dir0, s.h_tile_dir + s.seg_dir0, sizeof(uint32_t) * n_dir, cudaMemcpyHostToDevice, s.stream));
    CU(cudaMemcpyAsync(s.d_frame_tile0 + s.seg_frame0, s.h_frame_tile0 + s.seg_frame0, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, s.stream));
    c->a_h2d_bytes.fetch_add(sizeof(uint32_t) * ((uint64_t)n_dir + n), std::memory_order_relaxed);
    a.tile_dir = s.d_tile_dir + s.seg_dir0;
    a.frame_tile0 = s.d_frame_tile0 + s.seg_frame0;
  }
  a.rec_off = s.d_rec_off + s.seg_slot0;
  a.frame_geom = s.d_geom + s.seg_frame0;
  a.geoms = c->d_geoms;
  a.flags = c->d_flags + s.seg_log_base;
  a.counts = c->d_counts + s.seg_log_base;
  a.n_frames = n;
  const ScanPlan& plan = s.fmt != kLayoutNative ? c->plan_packed : c->plan;
  a.stages = plan.stages;
  a.max_cells = plan.cells;
  a.max_bit_words = plan.bit_words;
  rc = run_scan(c, a, plan, s.stream, s.seg_recs, (int)(&s - c->slabs));
  if (rc) return rc;
  CU(cudaEventRecord(s.done, s.stream));
  s.in_flight = true;
  s.launch_seq += 1;
  s.seg_frame0 = s.frames;
  s.seg_slot0 += n + 1;
  s.bytes = (s.bytes + 255) & ~255ull;
  s.seg_byte0 = s.bytes;
  s.seg_recs = 0;
  s.seg_dir0 = s.dir_slots;
  s.staged = false;
  s.reserved_end.store(s.bytes, std::memory_order_release);
  s.copy_head = s.bytes;
  return MSCAN_OK;
}

void recycle_slab(mscan_ctx* c, Slab& s) {
  std::lock_guard<std::mutex> issue(c->issue_mu);  // a late try_pump may be looking at copy_head
  s.bytes = 0;
  s.frames = 0;
  s.seg_frame0 = 0;
  s.seg_slot0 = 0;
  s.seg_byte0 = 0;
  s.seg_recs = 0;
  s.dir_slots = 0;
  s.seg_dir0 = 0;
  s.staged = false;
  s.copy_head = 0;
  s.reserved_end.store(0, std::memory_order_release);
  s.epoch += 1;
}

int wait_slab(mscan_ctx* c, Slab& s) {
  if (s.in_flight) {
    CU(use_device(c));
    CU(cudaEventSynchronize(s.done));
    s.in_flight = false;
  }
  recycle_slab(c, s);
  return MSCAN_OK;
}

// launch what the current slab holds and move on to the next slab of the ring
int flush_locked(mscan_ctx* c) {
  Slab& s = c->slabs[c->cur];
  if (s.frames == 0) return MSCAN_OK;
  int rc = launch_segment(c, s);
  if (rc) return rc;
  c->cur = (c->cur + 1) % kSlabs;
  return wait_slab(c, c->slabs[c->cur]);
}

int sync_scans_locked(mscan_ctx* c) {
  int rc = flush_locked(c);
  if (rc) return rc;
  if (c->dev_seq != c->dev_done_seq) {
    CU(cudaStreamSynchronize(c->dev_stream));
    c->dev_done_seq = c->dev_seq;
  }
  for (auto& s : c->slabs)
    if (s.in_flight) {
      CU(cudaEventSynchronize(s.done));
      s.in_flight = false;
      if (&s != &c->slabs[c->cur]) recycle_slab(c, s);
    }
  return MSCAN_OK;
}

// Waits until every frame the listed videos have submitted so far has its results in the frame log — and for nothing
// else: only slabs that hold frames of these videos are launched / waited for, and the waiting happens with `mu`
// released, so the decode threads of other videos keep submitting (the reference's workers never wait for each
// other either, src/pipeline.cpp:186-235). `lk` holds c->mu on entry and on return; iterators into c->videos are
// invalid afterwards.
int sync_videos(mscan_ctx* c, std::unique_lock<std::mutex>& lk, const uint32_t* ids, uint32_t n_ids) {
  struct Wait {
    int k;
    uint64_t epoch, seq;
    cudaEvent_t ev;
  };
  Wait waits[kSlabs];
  int n_wait = 0;
  for (int k = 0; k < kSlabs; ++k) {
    Slab& s = c->slabs[k];
    bool mine = false;
    for (uint32_t i = 0; i < n_ids && !mine; ++i) {
      auto it = c->videos.find(ids[i]);
      mine = it != c->videos.end() && it->second.slab_epoch[k] == s.epoch;
    }
    if (!mine) continue;
    if (s.frames > s.seg_frame0) {  // an open segment (only the current slab has one): launch it, keep filling the slab
      int rc = launch_segment(c, s);
      if (rc) return rc;
    }
    if (s.in_flight) waits[n_wait++] = Wait{k, s.epoch, s.launch_seq, s.done};
  }
  uint64_t dev_need = 0;  // frames handed over with mscan_submit_device
  for (uint32_t i = 0; i < n_ids; ++i) {
    auto it = c->videos.find(ids[i]);
    if (it != c->videos.end()) dev_need = std::max(dev_need, it->second.dev_seq);
  }
  const bool dev_wait = dev_need > c->dev_done_seq;
  const uint64_t dev_seq_now = c->dev_seq;  // what dev_done covers when we wait on it
  if (n_wait == 0 && !dev_wait) return MSCAN_OK;
  lk.unlock();
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < n_wait && e == cudaSuccess; ++i) e = cudaEventSynchronize(waits[i].ev);
  if (dev_wait && e == cudaSuccess) e = cudaEventSynchronize(c->dev_done);
  lk.lock();
  if (e != cudaSuccess) return fail(c, MSCAN_ERR_CUDA, "cudaEventSynchronize failed: %s", cudaGetErrorString(e));
  if (dev_wait) c->dev_done_seq = std::max(c->dev_done_seq, dev_seq_now);
  for (int i = 0; i < n_wait; ++i) {
    Slab& s = c->slabs[waits[i].k];
    if (s.epoch == waits[i].epoch && s.launch_seq == waits[i].seq && s.in_flight) {  // nothing was launched on it meanwhile
      s.in_flight = false;
      if (waits[i].k != c->cur) recycle_slab(c, s);
    }
  }
  return MSCAN_OK;
}

int replan(mscan_ctx* c, const std::vector<DevGeom>& geoms) {
  ScanPlan p, pp;
  if (!scan_plan_for(geoms.data(), (uint32_t)geoms.size(), c->smem_optin, false, &p) ||
      !scan_plan_for(geoms.data(), (uint32_t)geoms.size(), c->smem_optin, true, &pp))
    return fail(c, MSCAN_ERR_UNSUPPORTED, "block grid of %u cells: bit-rows do not fit shared memory", c->max_cells);
  if (p.global_cnt && (uint64_t)c->num_sms * p.ctas_per_sm * c->max_cells * sizeof(uint32_t) > (16ull << 30))
    return fail(c, MSCAN_ERR_UNSUPPORTED, "block grid of %u cells needs more than 16 GiB of counter scratch", c->max_cells);
  c->plan = p;
  c->plan_packed = pp;
  return MSCAN_OK;
}

int add_geometry(mscan_ctx* c, const mscan_geometry& g, uint32_t* idx) {
  if (!geom_ok(g)) return fail(c, MSCAN_ERR_INVALID, "bad geometry %dx%d", g.grid_w, g.grid_h);
  const DevGeom d = to_dev_geom(g);
  for (uint32_t i = 0; i < c->geoms.size(); ++i) {
    const DevGeom& e = c->geoms[i];
    if (e.gw == d.gw && e.gh == d.gh && e.y_min == d.y_min && e.y_max == d.y_max) {
      *idx = i;
      return MSCAN_OK;
    }
  }
  if (c->geoms.size() >= kMaxGeoms) return fail(c, MSCAN_ERR_CAPACITY, "too many distinct geometries");
  uint32_t cells, words;
  geom_need(d, &cells, &words);
  const uint32_t old_cells = c->max_cells, old_words = c->max_bit_words;
  c->max_cells = std::max(c->max_cells, cells);
  c->max_bit_words = std::max(c->max_bit_words, words);
  if (c->max_cells != old_cells || c->max_bit_words != old_words) {
    std::vector<DevGeom> all(c->geoms);
    all.push_back(d);
    int rc = replan(c, all);
    if (rc) {
      c->max_cells = old_cells;
      c->max_bit_words = old_words;
      return rc;
    }
  }
  *idx = (uint32_t)c->geoms.size();
  c->geoms.push_back(d);
  // the table is only appended to; in-flight kernels never read the new slot
  CU(cudaMemcpyAsync(c->d_geoms + *idx, &c->geoms[*idx], sizeof(DevGeom), cudaMemcpyHostToDevice, c->main_stream));
  CU(cudaStreamSynchronize(c->main_stream));
  return MSCAN_OK;
}

// The frame log is reused: when the bump pointer runs out of known-free frames, the gaps between the extents of the
// videos that are still open are searched first-fit, cyclically from the current head. Sets [log_head, log_limit) to
// the gap found; false when every frame of the log belongs to an open video. The caller has closed the open
// segment (its frames must be contiguous in the log) and frees nothing that a running kernel still writes:
// mscan_video_close waits for the scans in flight.
bool log_find_space(mscan_ctx* c) {
  std::vector<Extent> live;
  for (const auto& kv : c->videos)
    for (const Extent& e : kv.second.extents)
      if (e.n) live.push_back(e);
  std::sort(live.begin(), live.end(), [](const Extent& x, const Extent& y) { return x.start < y.start; });
  struct Gap {
    uint64_t a, b;
  };
  std::vector<Gap> gaps;
  uint64_t at = 0;
  for (const Extent& e : live) {
    if (e.start > at) gaps.push_back(Gap{at, e.start});
    at = std::max(at, e.start + e.n);
  }
  if (at < c->log_cap) gaps.push_back(Gap{at, c->log_cap});
  if (gaps.empty()) return false;
  // first gap that ends after the current head (continue forward), else wrap to the first gap of the log
  for (const Gap& g : gaps)
    if (g.b > c->log_head) {
      c->log_head = std::max(c->log_head, g.a);
      c->log_limit = g.b;
      return true;
    }
  c->log_head = gaps.front().a;
  c->log_limit = gaps.front().b;
  return true;
}

template <typename T>
int grow(mscan_ctx* c, T** d, uint64_t* cap, uint64_t need) {
  if (need <= *cap) return MSCAN_OK;
  uint64_t n = std::max<uint64_t>(need, *cap * 2);
  if (*d) cudaFree(*d);
  *d = nullptr;
  *cap = 0;
  if (cudaMalloc((void**)d, n * sizeof(T)) != cudaSuccess) return fail(c, MSCAN_ERR_NOMEM, "cudaMalloc of %llu bytes failed", (unsigned long long)(n * sizeof(T)));
  *cap = n;
  return MSCAN_OK;
}

uint64_t pow2_ge(uint64_t n) {
  uint64_t p = 1;
  while (p < n) p <<= 1;
  return p;
}

bool is_pinned(const void* p) {
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

}  // namespace

// =================================================================================================
extern "C" {

int mscan_abi_version(void) { return MSCAN_ABI_VERSION; }

const char* mscan_status_string(int s) {
  switch (s) {
    case MSCAN_OK: return "ok";
    case MSCAN_ERR_INVALID: return "invalid argument or state";
    case MSCAN_ERR_CUDA: return "CUDA error / no usable GPU";
    case MSCAN_ERR_NOMEM: return "out of memory";
    case MSCAN_ERR_CAPACITY: return "capacity exceeded";
    case MSCAN_ERR_UNSUPPORTED: return "unsupported geometry";
    default: return "unknown status";
  }
}

int mscan_device_count(int* n_out) {
  if (!n_out) return MSCAN_ERR_INVALID;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    *n_out = 0;
    return MSCAN_ERR_CUDA;
  }
  *n_out = n;
  return n > 0 ? MSCAN_OK : MSCAN_ERR_CUDA;
}

int mscan_params_default(mscan_params* p) {
  if (!p) return MSCAN_ERR_INVALID;
  p->mv_threshold_sq = 16.0;  // config.hpp:57
  p->block_size = 16;         // :63
  p->block_shift = 4;         // :69
  p->vectors_needed = 2;      // :75
  p->clusters_needed = 2;     // :81
  p->vertical_mask = 0.05f;   // :87
  p->adjacency = 4;           // the reference's 4-connectivity (motion_scanner.cpp:284-286)
  p->max_gap_sec = 5.0;       // :93
  p->padding_sec = 0.5;       // :99
  p->min_savings_pct = 5.0;   // :123
  return MSCAN_OK;
}

static bool env_double(const char* name, double* v) {
  const char* s = std::getenv(name);
  if (!s) return true;
  char* end = nullptr;
  const double x = std::strtod(s, &end);  // std::stod: leading whitespace ok, trailing junk ignored
  if (end == s) return false;
  *v = x;
  return true;
}
static bool env_int(const char* name, int32_t* v) {
  const char* s = std::getenv(name);
  if (!s) return true;
  char* end = nullptr;
  const long x = std::strtol(s, &end, 10);
  if (end == s) return false;
  *v = (int32_t)x;
  return true;
}
static bool env_float(const char* name, float* v) {
  const char* s = std::getenv(name);
  if (!s) return true;
  char* end = nullptr;
  const float x = std::strtof(s, &end);
  if (end == s) return false;
  *v = x;
  return true;
}

int mscan_params_from_env(mscan_params* p) {
  int rc = mscan_params_default(p);
  if (rc) return rc;
  bool ok = true;
  ok &= env_double("MV_THRESHOLD_SQ", &p->mv_threshold_sq);
  ok &= env_int("BLOCK_SIZE", &p->block_size);
  ok &= env_int("BLOCK_SHIFT", &p->block_shift);
  ok &= env_int("VECTORS_NEEDED", &p->vectors_needed);
  ok &= env_int("CLUSTERS_NEEDED", &p->clusters_needed);
  ok &= env_float("VERTICAL_MASK", &p->vertical_mask);
  ok &= env_int("CLUSTER_ADJACENCY", &p->adjacency);  // extension knob, not in the reference
  ok &= (p->adjacency == 4 || p->adjacency == 8);
  ok &= env_double("MAX_GAP_SEC", &p->max_gap_sec);
  ok &= env_double("PADDING_SEC", &p->padding_sec);
  ok &= env_double("MIN_SAVINGS_PCT", &p->min_savings_pct);
  return ok ? MSCAN_OK : MSCAN_ERR_INVALID;
}

int mscan_geometry_from_dims(const mscan_params* p, int width, int height, mscan_geometry* g) {
  if (!p || !g) return MSCAN_ERR_INVALID;
  if (p->block_shift < 0 || p->block_shift > 30) return MSCAN_ERR_INVALID;
  // motion_scanner.cpp:189-192: int arithmetic narrowed to int16_t
  const int16_t gw = (int16_t)((width + p->block_size - 1) >> p->block_shift);
  const int16_t gh = (int16_t)((height + p->block_size - 1) >> p->block_shift);
  g->grid_w = gw;
  g->grid_h = gh;
  // :196 — int16 → float, float32 multiply, truncate toward zero
  volatile float prod = (float)gh * p->vertical_mask;
  g->vertical_margin = (int)prod;
  g->reserved = 0;
  return MSCAN_OK;
}

int mscan_create(int device, const mscan_params* p, uint64_t max_log_frames, uint64_t slab_bytes, mscan_ctx** out) {
  if (!p || !out) return MSCAN_ERR_INVALID;
  *out = nullptr;
  if (p->block_shift < 0 || p->block_shift > 30) return MSCAN_ERR_INVALID;
  if (p->adjacency != 0 && p->adjacency != 4 && p->adjacency != 8) return MSCAN_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    return MSCAN_ERR_CUDA;  // no CPU fallback
  }
  process_cpus();  // capture the process's CPU mask before the caller starts pinning threads
  mscan_ctx* c = new (std::nothrow) mscan_ctx();
  if (!c) return MSCAN_ERR_NOMEM;
  c->device = device;
  c->params = *p;
  if (const char* t = std::getenv("MSCAN_TRACE")) c->trace = t[0] && t[0] != '0';
  threshold_to_int(p->mv_threshold_sq, &c->ithr, &c->keep_none);
  c->vec_need = (uint32_t)(uint8_t)p->vectors_needed;  // config.hpp:75 static_cast<uint8_t>
  c->clust_need = p->clusters_needed < 1 ? 1u : (uint32_t)p->clusters_needed;
  c->adj8 = p->adjacency == 8 ? 1u : 0u;
  c->log_cap = max_log_frames ? max_log_frames : (16ull << 20);
  c->log_limit = c->log_cap;
  c->slab_bytes = slab_bytes ? ((slab_bytes + 255) & ~255ull) : (64ull << 20);
  c->slab_frames = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(c->slab_bytes / 1024, 4096), 1u << 22);
  // every frame with records has at least one tile and a tile is at most ~8.4 KB; + one closing slot per segment
  c->slab_dir_cap = (uint32_t)std::min<uint64_t>(2ull * c->slab_frames + c->slab_bytes / 2048 + 16, 1u << 24);
  if (const char* w = std::getenv("MSCAN_COPY_WINDOW_KB")) {  // experiments: H2D copy window, rounded down to a power of two
    const long kb = std::atol(w);
    uint32_t sh = 16;
    while (sh < 30 && (2ull << sh) <= (uint64_t)std::max(64l, kb) * 1024ull) ++sh;
    c->win_shift = sh;
  }

  auto bail = [&](int code) {
    mscan_destroy(c);
    return code;
  };
#define CUB_(call)                                         \
  do {                                                     \
    if ((call) != cudaSuccess) {                           \
      cudaGetLastError();                                  \
      return bail(MSCAN_ERR_CUDA);                         \
    }                                                      \
  } while (0)
  CUB_(cudaSetDevice(device));
  int v = 0;
  CUB_(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
  c->num_sms = v;
  CUB_(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  c->smem_optin = (uint32_t)v;
  CUB_(scan_configure(c->smem_optin));
  CUB_(cudaStreamCreateWithFlags(&c->main_stream, cudaStreamNonBlocking));
  CUB_(cudaStreamCreateWithFlags(&c->dev_stream, cudaStreamNonBlocking));
  CUB_(cudaEventCreateWithFlags(&c->dev_done, cudaEventDisableTiming));
  CUB_(cudaEventCreateWithFlags(&c->dev_ready, cudaEventDisableTiming));
  for (auto& e : c->slot_ev) CUB_(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CUB_(cudaMalloc((void**)&c->d_geoms, sizeof(DevGeom) * kMaxGeoms));
  CUB_(cudaMalloc((void**)&c->d_work, sizeof(uint32_t) * 2 * kWorkSlots));
  CUB_(cudaMemset(c->d_work, 0, sizeof(uint32_t) * 2 * kWorkSlots));
  CUB_(cudaMalloc((void**)&c->d_pts, sizeof(double) * c->log_cap));
  CUB_(cudaMalloc((void**)&c->d_flags, c->log_cap));
  CUB_(cudaMalloc((void**)&c->d_counts, sizeof(uint32_t) * c->log_cap));
  for (auto& s : c->slabs) {
    const size_t n_win = (size_t)((c->slab_bytes + 256) >> c->win_shift) + 2;
    s.pend.reset(new (std::nothrow) std::atomic<int32_t>[n_win * kPendStride]);
    if (!s.pend) return bail(MSCAN_ERR_NOMEM);
    for (size_t k = 0; k < n_win * kPendStride; ++k) s.pend[k].store(0, std::memory_order_relaxed);
    CUB_(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CUB_(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    CUB_(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
    CUB_(cudaMalloc((void**)&s.d_recs, c->slab_bytes + 256));
    CUB_(cudaMalloc((void**)&s.d_rec_off, sizeof(uint64_t) * (2 * (size_t)c->slab_frames + 2)));  // k+1 slots per segment
    CUB_(cudaMalloc((void**)&s.d_geom, sizeof(uint32_t) * c->slab_frames));
    CUB_(cudaHostAlloc((void**)&s.h_rec_off, sizeof(uint64_t) * (2 * (size_t)c->slab_frames + 2), cudaHostAllocDefault));
    CUB_(cudaHostAlloc((void**)&s.h_geom, sizeof(uint32_t) * c->slab_frames, cudaHostAllocDefault));
    CUB_(cudaHostAlloc((void**)&s.h_pts, sizeof(double) * c->slab_frames, cudaHostAllocDefault));
    CUB_(cudaMalloc((void**)&s.d_tile_dir, sizeof(uint32_t) * c->slab_dir_cap));
    CUB_(cudaMalloc((void**)&s.d_frame_tile0, sizeof(uint32_t) * c->slab_frames));
    CUB_(cudaHostAlloc((void**)&s.h_tile_dir, sizeof(uint32_t) * c->slab_dir_cap, cudaHostAllocDefault));
    CUB_(cudaHostAlloc((void**)&s.h_frame_tile0, sizeof(uint32_t) * c->slab_frames, cudaHostAllocDefault));
  }
#undef CUB_
  *out = c;
  return MSCAN_OK;
}

int mscan_destroy(mscan_ctx* c) {
  if (!c) return MSCAN_OK;
  fold_stats(c);
  if (c->trace && !c->api_times.empty()) {
    std::fprintf(stderr, "[mscan trace] device %d: wall time per entry point (lock waits included)\n", c->device);
    for (const auto& kv : c->api_times)
      std::fprintf(stderr, "[mscan trace]   %-28s calls %8llu  total %10.3f ms  max %9.3f ms\n", kv.first.c_str(),
                   (unsigned long long)kv.second.calls, kv.second.total_ms, kv.second.max_ms);
    std::fprintf(stderr, "[mscan trace]   projected %llu records in %.3f ms; H2D %llu B, D2H %llu B; K-A launches %llu, K-C launches %llu\n",
                 (unsigned long long)c->stats.records_projected, c->stats.project_ms, (unsigned long long)c->stats.h2d_bytes,
                 (unsigned long long)c->stats.d2h_bytes, (unsigned long long)c->stats.scan_launches,
                 (unsigned long long)c->stats.segment_launches);
  }
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (auto& p : c->ev_pending) {
    cudaEventDestroy(p.a);
    cudaEventDestroy(p.b);
  }
  for (auto& p : c->ev_free) {
    cudaEventDestroy(p.a);
    cudaEventDestroy(p.b);
  }
  for (auto& s : c->slabs) {
    if (s.d_recs) cudaFree(s.d_recs);
    if (s.h_recs) cudaFreeHost(s.h_recs);
    if (s.d_rec_off) cudaFree(s.d_rec_off);
    if (s.h_rec_off) cudaFreeHost(s.h_rec_off);
    if (s.d_geom) cudaFree(s.d_geom);
    if (s.h_geom) cudaFreeHost(s.h_geom);
    if (s.h_pts) cudaFreeHost(s.h_pts);
    if (s.d_tile_dir) cudaFree(s.d_tile_dir);
    if (s.d_frame_tile0) cudaFree(s.d_frame_tile0);
    if (s.h_tile_dir) cudaFreeHost(s.h_tile_dir);
    if (s.h_frame_tile0) cudaFreeHost(s.h_frame_tile0);
    if (s.done) cudaEventDestroy(s.done);
    if (s.copied) cudaEventDestroy(s.copied);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  cudaFree(c->d_geoms);
  cudaFree(c->d_work);
  cudaFree(c->d_cnt_scratch);
  cudaFree(c->d_pts);
  cudaFree(c->d_flags);
  cudaFree(c->d_counts);
  cudaFree(c->d_ts_a);
  cudaFree(c->d_ts_b);
  cudaFree(c->d_segs);
  cudaFree(c->d_jobs);
  cudaFree(c->d_exts);
  cudaFree(c->d_res);
  cudaFree(c->d_user_geoms);
  if (c->h_jobs) cudaFreeHost(c->h_jobs);
  if (c->h_exts) cudaFreeHost(c->h_exts);
  if (c->h_res) cudaFreeHost(c->h_res);
  if (c->main_stream) cudaStreamDestroy(c->main_stream);
  if (c->dev_stream) cudaStreamDestroy(c->dev_stream);
  if (c->dev_done) cudaEventDestroy(c->dev_done);
  if (c->dev_ready) cudaEventDestroy(c->dev_ready);
  for (auto& e : c->slot_ev)
    if (e) cudaEventDestroy(e);
  cudaGetLastError();
  delete c;
  return MSCAN_OK;
}

const char* mscan_last_error(mscan_ctx* c) { return c ? c->err.c_str() : "null context"; }

int mscan_get_params(mscan_ctx* c, mscan_params* p) {
  if (!c || !p) return MSCAN_ERR_INVALID;
  *p = c->params;
  return MSCAN_OK;
}

int mscan_sync(mscan_ctx* c) try {
  ApiTimer trace_(c, "mscan_sync");
  if (!c) return MSCAN_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->device));
  int rc = sync_scans_locked(c);
  if (rc) return rc;
  CU(cudaStreamSynchronize(c->main_stream));
  return MSCAN_OK;
} catch (...) {
  return on_exception(c);
}

int mscan_get_stats(mscan_ctx* c, mscan_stats* s) try {
  if (!c || !s) return MSCAN_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  cudaSetDevice(c->device);
  drain_events(c);
  fold_stats(c);
  *s = c->stats;
  return MSCAN_OK;
} catch (...) {
  return on_exception(c);
}

int mscan_reset_stats(mscan_ctx* c) try {
  if (!c) return MSCAN_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  cudaSetDevice(c->device);
  drain_events(c);
  fold_stats(c);
  c->stats = mscan_stats{};
  return MSCAN_OK;
} catch (...) {
  return on_exception(c);
}

int mscan_set_profiling(mscan_ctx* c, int enabled) {
  if (!c) return MSCAN_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  c->profiling = enabled != 0;
  return MSCAN_OK;
}

// ---- host-fed path ------------------------------------------------------------------------------
int mscan_video_open_geometry(mscan_ctx* c, uint32_t video_id, const mscan_geometry* g) try {
  ApiTimer trace_(c, "mscan_video_open");
  if (!c || !g) return MSCAN_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->device));
  if (c->videos.count(video_id)) return fail(c, MSCAN_ERR_INVALID, "video %u already open", video_id);
  // a larger grid changes the kernel's shared-memory plan: finish what is staged under the old one
  uint32_t idx = 0;
  const uint32_t old_cells = c->max_cells;
  {
    uint32_t cells, words;
    if (!geom_ok(*g)) return fail(c, MSCAN_ERR_INVALID, "bad geometry %dx%d", g->grid_w, g->grid_h);
    geom_need(to_dev_geom(*g), &cells, &words);
    if (cells > old_cells) {
      int rc = flush_locked(c);
      if (rc) return rc;
    }
  }
  int rc = add_geometry(c, *g, &idx);
  if (rc) return rc;
  Video v;
  v.geom = idx;
  c->videos.emplace(video_id, std::move(v));
  return MSCAN_OK;
} catch (...) {
  return on_exception(c);
}

int mscan_video_open(mscan_ctx* c, uint32_t video_id, int width, int height) {
  if (!c) return MSCAN_ERR_INVALID;
  mscan_geometry g;
  int rc = mscan_geometry_from_dims(&c->params, width, height, &g);
  if (rc) return fail(c, rc, "bad dimensions %dx%d", width, height);
  return mscan_video_open_geometry(c, video_id, &g);
}

// How one reservation gets its records into the slab (decided under mu, executed outside it).
enum FillKind { kFillNone = 0, kFillProject, kFillMemcpy, kFillInPlace };

// Commits a reservation whose bytes are in place: the copy windows first (staged reservations), then the segment's
// writer count (launch_segment waits for the latter).
static void commit_fill(mscan_ctx* c, Slab& s, bool staged, uint64_t off, uint64_t nbytes) {
  bool completed_window = false;
  if (staged) {
    std::atomic_thread_fence(std::memory_order_release);
    for (uint64_t k = off >> c->win_shift; k <= (off + nbytes - 1) >> c->win_shift; ++k)
      if (s.pend[k * kPendStride].fetch_sub(1, std::memory_order_acq_rel) == 1) completed_window = true;
  } else s.writers.fetch_sub(1, std::memory_order_release);
  // after the last decrement `s` may be launched and recycled by another thread at any moment; pumping is state-based
  // and therefore still safe (it then finds nothing, or windows of the slab's next life)
  if (completed_window && (s.reserved_end.load(std::memory_order_acquire) >> c->win_shift) > (off >> c->win_shift)) try_pump(c, s);
}

// Executes one reservation's fill outside the context mutex, then commits it.
static int fill_and_commit(mscan_ctx* c, Slab& s, FillKind kind, uint64_t off, uint64_t nbytes, const uint8_t* from,
                           uint64_t n_recs) {
  int rc = MSCAN_OK;
  if (kind == kFillProject) {
    const auto t0 = std::chrono::steady_clock::now();
    uint64_t* to = reinterpret_cast<uint64_t*>(s.h_recs + off);
    PackPool* pool = (n_recs > kPoolMinRecs && c->pack_threads != 1) ? shared_pool() : nullptr;
    if (pool && pool->workers() > 0) pool->run(from, to, n_recs, c->pack_threads);
    else project_records(from, n_recs, to);
    c->a_project_ns.fetch_add((uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count(),
                              std::memory_order_relaxed);
    c->a_records_projected.fetch_add(n_recs, std::memory_order_relaxed);
  } else if (kind == kFillMemcpy) {
    stream_copy(from, s.h_recs + off, nbytes);  // (the ring is read next by the DMA engine, not by this core)
  } else if (kind == kFillInPlace) {
    std::lock_guard<std::mutex> issue(c->issue_mu);
    cudaError_t e = use_device(c);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s.d_recs + off, from, nbytes, cudaMemcpyHostToDevice, s.stream);
    if (e != cudaSuccess) rc = MSCAN_ERR_CUDA;
    c->a_h2d_bytes.fetch_add(nbytes, std::memory_order_relaxed);
  }
  commit_fill(c, s, kind == kFillProject || kind == kFillMemcpy, off, nbytes);
  return rc;
}

// Places frames that are already in the static-elided form (mvz, host_project.cpp) into the slab ring: whole frames,
// in as many reservations as slab / log / directory room demands; the bytes are moved outside the mutex (streaming copy
// into the pinned ring, or DMA'd in place out of pinned caller memory) and committed.
struct ElidedPlacer {
  mscan_ctx* c;
  uint32_t video_id;
  uint32_t n_frames;         // frames of the whole submit call
  uint64_t* first_frame_out;
  // The call's video-local indices [vbase, vbase + n_frames) are reserved in the SAME critical section that places its
  // first frames: per-frame submits from many decode threads then get video order == log order, and the video stays
  // one extent instead of one per frame.
  uint64_t vbase = 0;
  bool have_vbase = false;
  std::unique_lock<std::mutex> lk;

  ElidedPlacer(mscan_ctx* ctx, uint32_t vid, uint32_t n, uint64_t* first) : c(ctx), video_id(vid), n_frames(n), first_frame_out(first), lk(ctx->mu, std::defer_lock) {}

  void give_back(uint32_t placed) {  // like submit_impl's Rollback
    if (!have_vbase) return;
    if (!lk.owns_lock()) lk.lock();
    auto it = c->videos.find(video_id);
    if (it != c->videos.end() && it->second.n_frames == vbase + n_frames) it->second.n_frames = vbase + placed;
    lk.unlock();
  }

  // One piece: nf frames starting at frame f0 of the call. data: the piece's encoded bytes; frame_end[i + 1]: end byte
  // offset of frame i in data (frame_end[0] == 0); frame_tile[i]: index of frame i's first tile in tile_end (size nf + 1);
  // tile_end[t]: end of tile t in 16-byte units from the start of data. in_place: data is pinned caller memory.
  int place(uint32_t f0, uint32_t nf, const double* pts, const uint32_t* rec_count, const uint8_t* data, const uint64_t* frame_end,
            const uint32_t* frame_tile, const uint32_t* tile_end, bool in_place) {
    uint32_t i = 0;
    while (i < nf) {
      lock_briefly(lk);
      auto it = c->videos.find(video_id);
      if (it == c->videos.end()) {
        lk.unlock();
        return fail(c, MSCAN_ERR_INVALID, "video %u is not open", video_id);
      }
      Video& v = it->second;
      if (!have_vbase) {
        vbase = v.n_frames;
        v.n_frames += n_frames;
        have_vbase = true;
        if (first_frame_out) *first_frame_out = vbase;
      }
      Slab* s = &c->slabs[c->cur];
      int rc = MSCAN_OK;
      if (s->frames > s->seg_frame0 && (s->fmt != kLayoutMvz || s->staged == in_place)) rc = launch_segment(c, *s);
      if (!rc && c->log_head >= c->log_limit) {
        rc = launch_segment(c, *s);
        if (!rc && !log_find_space(c))
          rc = fail(c, MSCAN_ERR_CAPACITY, "frame log full (%llu frames, all owned by open videos); close videos or create a larger context",
                    (unsigned long long)c->log_cap);
      }
      if (rc) {
        lk.unlock();
        give_back(f0 + i);
        return rc;
      }
      const bool opens = s->frames == s->seg_frame0;
      const uint64_t b0 = frame_end[i], log_room = c->log_limit - c->log_head;
      uint32_t take = 0;
      while (i + take < nf && s->frames + take < c->slab_frames && take < log_room) {
        const uint64_t nb = frame_end[(size_t)i + take + 1] - b0;
        const uint32_t tiles = frame_tile[(size_t)i + take + 1] - frame_tile[i];
        if (s->bytes + nb > c->slab_bytes) break;
        if ((uint64_t)s->dir_slots + (opens ? 1u : 0u) + tiles + 1u > c->slab_dir_cap) break;
        ++take;
      }
      if (take == 0) {
        if (s->frames == 0) rc = fail(c, MSCAN_ERR_CAPACITY, "frame with %u records exceeds the slab size (%llu bytes)", rec_count[i],
                                      (unsigned long long)c->slab_bytes);
        else rc = flush_locked(c);
        lk.unlock();
        if (rc) {
          give_back(f0 + i);
          return rc;
        }
        continue;
      }
      const uint64_t nbytes = frame_end[(size_t)i + take] - b0;
      if (!in_place && nbytes && !s->h_recs) {
        cudaError_t e = use_device(c);
        if (e == cudaSuccess) e = cudaHostAlloc((void**)&s->h_recs, c->slab_bytes, cudaHostAllocDefault);
        if (e != cudaSuccess) {
          lk.unlock();
          give_back(f0 + i);
          return fail(c, MSCAN_ERR_NOMEM, "cudaHostAlloc of the staging slab failed: %s", cudaGetErrorString(e));
        }
      }
      if (opens) {
        s->seg_log_base = c->log_head;
        s->fmt = (uint8_t)kLayoutMvz;
        s->seg_dir0 = s->dir_slots;
        s->h_tile_dir[s->dir_slots++] = 0;  // tile 0 of the segment starts at its base
        std::lock_guard<std::mutex> issue(c->issue_mu);
        s->staged = !in_place;
        s->copy_head = s->seg_byte0;
      }
      const uint64_t off = s->bytes;
      // piece-relative 16-byte units → segment-relative: frame i's first tile starts where the segment's last one ended
      const uint32_t rel16 = (uint32_t)((off - s->seg_byte0) >> 4) - (uint32_t)(b0 >> 4);
      const uint32_t tiles_before = s->dir_slots - s->seg_dir0 - 1;
      uint64_t r = s->seg_recs, take_recs = 0;
      const uint32_t slot = s->seg_slot0 + (s->frames - s->seg_frame0);
      for (uint32_t k = 0; k < take; ++k) {
        s->h_rec_off[slot + k] = r;
        r += rec_count[i + k];
        take_recs += rec_count[i + k];
        s->h_geom[s->frames + k] = v.geom;
        s->h_pts[s->frames + k] = pts[i + k];
        s->h_frame_tile0[s->frames + k] = tiles_before + (frame_tile[(size_t)i + k] - frame_tile[i]);
      }
      for (uint32_t t = frame_tile[i]; t < frame_tile[(size_t)i + take]; ++t) s->h_tile_dir[s->dir_slots++] = tile_end[t] + rel16;
      const uint64_t log_at = c->log_head, vpos = vbase + f0 + i;
      if (!v.extents.empty() && v.extents.back().start + v.extents.back().n == log_at && v.extents.back().vpos + v.extents.back().n == vpos)
        v.extents.back().n += take;
      else v.extents.push_back(Extent{log_at, take, vpos});
      v.slab_epoch[c->cur] = s->epoch;
      c->log_head += take;
      s->frames += take;
      s->seg_recs += take_recs;
      s->bytes += nbytes;
      const bool slab_full = s->frames == c->slab_frames || s->bytes + 64 * 1024 > c->slab_bytes;
      const uint64_t my_epoch = s->epoch;
      if (nbytes) {
        if (in_place) s->writers.fetch_add(1, std::memory_order_relaxed);
        else
          for (uint64_t k = off >> c->win_shift; k <= (off + nbytes - 1) >> c->win_shift; ++k) s->pend[k * kPendStride].fetch_add(1, std::memory_order_relaxed);
        s->reserved_end.store(s->bytes, std::memory_order_release);
      }
      lk.unlock();
      if (nbytes) {
        if (in_place) {
          std::lock_guard<std::mutex> issue(c->issue_mu);
          cudaError_t e = use_device(c);
          if (e == cudaSuccess) e = cudaMemcpyAsync(s->d_recs + off, data + b0, nbytes, cudaMemcpyHostToDevice, s->stream);
          if (e != cudaSuccess) rc = MSCAN_ERR_CUDA;
          c->a_h2d_bytes.fetch_add(nbytes, std::memory_order_relaxed);
        } else {
          stream_copy(data + b0, s->h_recs + off, nbytes);
        }
        commit_fill(c, *s, !in_place, off, nbytes);
        if (rc) {
          give_back(f0 + i);
          return fail(c, rc, "cudaMemcpyAsync of pinned elided records failed");
        }
      }
      i += take;
      if (slab_full) {
        lock_briefly(lk);
        if (&c->slabs[c->cur] == s && s->epoch == my_epoch) rc = flush_locked(c);
        lk.unlock();
        if (rc) {
          give_back(f0 + i);
          return rc;
        }
      }
    }
    return MSCAN_OK;
  }
};

// mscan_submit under MSCAN_STAGING_ELIDE (and AUTO for small pageable submits): the calling thread encodes its native
// records as mvz into a private scratch, piece by piece, and places each piece. The caller (submit_impl) has released
// the mutex.
static int submit_elide(mscan_ctx* c, uint32_t video_id, uint32_t n_frames, const double* pts, const uint32_t* rec_count,
                        const uint8_t* src, uint64_t* first_frame_out) {
  thread_local std::vector<uint8_t> scratch;       // one encoded piece
  thread_local std::vector<uint32_t> tile_end;     // per tile of the piece: end, 16-byte units from the piece start
  thread_local std::vector<uint32_t> frame_tile;   // per frame of the piece (+1): index of its first tile
  thread_local std::vector<uint64_t> frame_end;    // per frame of the piece (+1): end byte offset in the piece
  ElidedPlacer placer(c, video_id, n_frames, first_frame_out);
  uint32_t f = 0;
  uint64_t src_rec = 0;
  while (f < n_frames) {
    // encode a piece outside the mutex: frames [f, g) with at most kPoolMinRecs records (one frame at least)
    uint32_t g = f;
    uint64_t piece_recs = 0;
    while (g < n_frames && (g == f || piece_recs + rec_count[g] <= kPoolMinRecs)) piece_recs += rec_count[g++];
    const uint32_t nf = g - f;
    if (piece_recs && !src) {
      placer.give_back(f);
      return fail(c, MSCAN_ERR_INVALID, "null recs with non-zero rec_count");
    }
    const auto t0 = std::chrono::steady_clock::now();
    scratch.resize((size_t)mvz_bound(piece_recs, nf));
    tile_end.clear();
    frame_tile.resize((size_t)nf + 1);
    frame_end.resize((size_t)nf + 1);
    uint64_t at = 0, done = 0;
    uint32_t nt = 0;
    frame_end[0] = 0;
    for (uint32_t i = 0; i < nf; ++i) {
      const uint32_t n = rec_count[f + i], tiles = (n + kMvzTileRecs - 1) / kMvzTileRecs;
      frame_tile[i] = nt;
      tile_end.resize((size_t)nt + tiles);
      if (n) {
        const uint64_t bytes = mvz_encode_frame(src + (size_t)kRecBytes * (src_rec + done), n, scratch.data() + at, tile_end.data() + nt);
        for (uint32_t t = 0; t < tiles; ++t) tile_end[nt + t] += (uint32_t)(at >> 4);
        at += bytes;
      }
      nt += tiles;
      done += n;
      frame_end[(size_t)i + 1] = at;
    }
    frame_tile[nf] = nt;
    c->a_project_ns.fetch_add((uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count(),
                              std::memory_order_relaxed);
    c->a_records_projected.fetch_add(piece_recs, std::memory_order_relaxed);
    c->a_records_elided.fetch_add(piece_recs, std::memory_order_relaxed);
    c->a_elided_bytes.fetch_add(at, std::memory_order_relaxed);
    int rc = placer.place(f, nf, pts + f, rec_count + f, scratch.data(), frame_end.data(), frame_tile.data(), tile_end.data(), false);
    if (rc) return rc;
    f = g;
    src_rec += piece_recs;
  }
  return MSCAN_OK;
}

// A submit call that hands its frames over in several pieces (submit_compact): the video-local index range of the WHOLE
// call is reserved by the piece that is placed first — in the same critical section as its first frames, so that
// per-frame submits from many decode threads keep video order == log order — and the later pieces place into it.
struct CallSpan {
  uint32_t total_frames;  // frames of the whole call
  uint32_t f0;            // index, in the call, of this piece's first frame
  uint64_t vbase;         // set by the first piece
  bool have;
  // what the piece's host pass did: added to the context's counters inside the piece's own critical section (four
  // shared atomics per frame from every decode thread would cost more than the bookkeeping they count)
  uint64_t st_records = 0, st_wire_bytes = 0, st_ns = 0;
};

static int submit_impl(mscan_ctx* c, uint32_t video_id, uint32_t n_frames, const double* pts, const uint32_t* rec_count,
                       const void* recs, bool src_packed, uint64_t* first_frame_out, CallSpan* span);

// mscan_submit under MSCAN_STAGING_COMPACT (and AUTO for a decode thread's own frame while the threshold is positive):
// the calling thread writes the projections of the MOVING records of its frames into a private scratch, piece by piece,
// and each piece goes through the packed submit (pageable mscan_mv8 → streaming copy into the pinned ring). A frame's
// record count on the GPU side is its number of moving records; results are unchanged because a static record cannot
// pass motion_scanner.cpp:251 while T² > 0 (host_project.cpp, compact_moving). The caller has released the mutex.
static int submit_compact(mscan_ctx* c, uint32_t video_id, uint32_t n_frames, const double* pts, const uint32_t* rec_count,
                          const uint8_t* src, uint64_t* first_frame_out) {
  thread_local std::vector<uint64_t> scratch;    // one piece: mscan_mv8 of its moving records
  thread_local std::vector<uint32_t> moving;     // per frame of the piece: how many
  CallSpan span{n_frames, 0, 0, false};
  uint32_t f = 0;
  uint64_t src_rec = 0;
  while (f < n_frames) {
    uint32_t g = f;
    uint64_t piece_recs = 0;
    while (g < n_frames && (g == f || piece_recs + rec_count[g] <= kPoolMinRecs)) piece_recs += rec_count[g++];
    const uint32_t nf = g - f;
    if (piece_recs && !src) {
      if (span.have) {  // give the unplaced rest of the reserved range back (like submit_impl's own rollback)
        std::lock_guard<std::mutex> lk(c->mu);
        auto it = c->videos.find(video_id);
        if (it != c->videos.end() && it->second.n_frames == span.vbase + n_frames) it->second.n_frames = span.vbase + f;
      }
      return fail(c, MSCAN_ERR_INVALID, "null recs with non-zero rec_count");
    }
    const auto t0 = std::chrono::steady_clock::now();
    scratch.resize((size_t)piece_recs + 8);
    moving.resize(nf);
    uint64_t at = 0, done = 0;
    for (uint32_t i = 0; i < nf; ++i) {
      const uint32_t n = rec_count[f + i];
      const uint64_t m = n ? compact_moving(src + (size_t)kRecBytes * (src_rec + done), n, scratch.data() + at) : 0;
      moving[i] = (uint32_t)m;
      at += m;
      done += n;
    }
    span.st_ns = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    span.st_records = piece_recs;
    span.st_wire_bytes = 8 * at;
    span.f0 = f;
    int rc = submit_impl(c, video_id, nf, pts + f, moving.data(), scratch.data(), true, f == 0 ? first_frame_out : nullptr, &span);
    if (rc) return rc;
    f = g;
    src_rec += piece_recs;
  }
  return MSCAN_OK;
}

// Shared body of mscan_submit (native 40-byte records) and mscan_submit_packed (mscan_mv8).
static int submit_impl(mscan_ctx* c, uint32_t video_id, uint32_t n_frames, const double* pts, const uint32_t* rec_count,
                       const void* recs, bool src_packed, uint64_t* first_frame_out, CallSpan* span) try {
  ApiTimer trace_(c, "mscan_submit[_packed]");
  if (!c) return MSCAN_ERR_INVALID;
  if (n_frames == 0) {
    if (first_frame_out) {
      std::lock_guard<std::mutex> lk(c->mu);
      auto it0 = c->videos.find(video_id);
      *first_frame_out = it0 == c->videos.end() ? 0 : it0->second.n_frames;
    }
    return MSCAN_OK;
  }
  if (!pts || !rec_count) return fail(c, MSCAN_ERR_INVALID, "null pts/rec_count");
  const uint8_t* src = reinterpret_cast<const uint8_t*>(recs);
  const uint64_t in_stride = src_packed ? (uint64_t)kPackedBytes : (uint64_t)kRecBytes;
  // is the source pinned? (before taking the mutex; see PinnedRanges)
  bool pinned = false;
  uint64_t total = 0;
  for (uint32_t i = 0; i < n_frames; ++i) total += rec_count[i];
  if (recs) {
    const uint64_t src_bytes = total * in_stride;
    if (src_bytes) {
      pinned = pinned_ranges().contains(recs, src_bytes);
      if (!pinned && src_bytes >= kAttrQueryBytes) pinned = is_pinned(recs);
    }
  }
  // Static macroblocks (src == dst) are not sent as 8 bytes. While the threshold is positive they cannot vote and the
  // calling thread sends the moving records only (MSCAN_STAGING_COMPACT; decided without the mutex: the pieces take it
  // themselves); otherwise they go as their 4 dst bytes (mvz, MSCAN_STAGING_ELIDE — the cluster kernel reads native and
  // mv8 records only).
  const int mode = c->staging_mode.load(std::memory_order_relaxed);
  const bool small_auto = mode == MSCAN_STAGING_AUTO && !pinned && total <= kPoolMinRecs;
  if (!src_packed && !span && (mode == MSCAN_STAGING_COMPACT || small_auto) && c->ithr > 0)
    return submit_compact(c, video_id, n_frames, pts, rec_count, src, first_frame_out);
  std::unique_lock<std::mutex> lk(c->mu, std::defer_lock);
  lock_briefly(lk);
  auto it = c->videos.find(video_id);
  if (it == c->videos.end()) return fail(c, MSCAN_ERR_INVALID, "video %u is not open", video_id);
  // static-elided transport (mvz): native records, encoded by the calling thread before it reserves anything; the
  // cluster kernel (grids beyond one CTA's shared memory) reads native and mv8 records only
  // (also what AUTO does with a pageable submit small enough that the caller projects alone — a decode thread handing
  // over its own frame: the encoding costs about what the projection costs and nearly halves the bytes on the link)
  const bool elide = mode == MSCAN_STAGING_ELIDE || mode == MSCAN_STAGING_COMPACT || small_auto;
  if (!src_packed && elide && !c->plan_packed.cluster && !c->plan_packed.global_cnt) {
    lk.unlock();
    return submit_elide(c, video_id, n_frames, pts, rec_count, src, first_frame_out);
  }
  // this call owns video-local indices [vbase, vbase + n_frames) (a piece of a larger call: its part of that call's range)
  CallSpan own{n_frames, 0, 0, false};
  if (!span) span = &own;
  if (!span->have) {
    span->vbase = it->second.n_frames;
    it->second.n_frames += span->total_frames;
    span->have = true;
  }
  if (span->st_records) {
    it->second.st_records += span->st_records;
    it->second.st_wire_bytes += span->st_wire_bytes;
    it->second.st_ns += span->st_ns;
    span->st_records = 0;
  }
  const uint64_t vbase = span->vbase + span->f0;
  if (first_frame_out) *first_frame_out = vbase;
  // on failure the indices not yet backed by log frames are given back (when no later call has reserved behind them)
  struct Rollback {
    mscan_ctx* c;
    std::unique_lock<std::mutex>& lk;
    uint32_t video_id, n_frames;
    const uint32_t& f;
    uint64_t vbase;       // of this piece
    uint64_t call_end;    // end of the whole call's reserved range
    bool armed = true;
    ~Rollback() {
      if (!armed || f >= n_frames) return;
      if (!lk.owns_lock()) lk.lock();
      auto it = c->videos.find(video_id);
      if (it != c->videos.end() && it->second.n_frames == call_end) it->second.n_frames = vbase + f;
    }
  };
  // How the records reach the slab: native+pinned → DMA in place (40 B/record over PCIe, no host work);
  // native+pageable → the staging pass writes only the 8 bytes the path reads (8 B/record over PCIe);
  // packed → as is (DMA in place when pinned, memcpy into staging otherwise).
  bool project = false;
  if (!src_packed) {
    if (mode == MSCAN_STAGING_PACK || mode == MSCAN_STAGING_ELIDE || mode == MSCAN_STAGING_COMPACT) project = true;
    else if (mode == MSCAN_STAGING_NATIVE) project = false;
    else project = !pinned;
  }
  const uint8_t slab_fmt = (src_packed || project) ? (uint8_t)kLayoutMv8 : (uint8_t)kLayoutNative;
  const bool slab_packed = slab_fmt != kLayoutNative;
  const bool staged = project || !pinned;
  const FillKind kind = project ? kFillProject : (pinned ? kFillInPlace : kFillMemcpy);
  const uint64_t out_stride = slab_packed ? (uint64_t)kPackedBytes : (uint64_t)kRecBytes;
  // Large projected submits go out in pieces. A piece is what one reservation's writer holds while it projects, and a
  // slab's launch waits for its writers: when the calling thread projects alone (pool off, or several decode workers
  // each feeding whole chunks) pieces stay small — 256 Ki records ≈ 1 ms — so that a slab flip never waits long for a
  // neighbour; a pool job (all the process's cores on one piece) takes 4 Mi records at a time.
  const bool pool_job = project && total > kPoolMinRecs && c->pack_threads != 1 && shared_pool() && shared_pool()->workers() > 0;
  const uint64_t max_take_recs = project ? (pool_job ? (4ull << 20) : kPoolMinRecs) : (src_packed && !pinned ? (1ull << 20) : ~0ull);
  uint32_t f = 0;
  uint64_t src_rec = 0;
  int result = MSCAN_OK;
  Rollback rollback{c, lk, video_id, n_frames, f, vbase, span->vbase + span->total_frames};
  while (f < n_frames) {
    // (re)validate the video: the mutex was released while the previous piece was being filled
    it = c->videos.find(video_id);
    if (it == c->videos.end()) return fail(c, MSCAN_ERR_INVALID, "video %u was closed during a submit", video_id);
    Video& v = it->second;
    Slab* s = &c->slabs[c->cur];
    if (s->frames > s->seg_frame0 && (s->fmt != slab_fmt || s->staged != staged)) {
      // one record format and one route (staging / in-place DMA) per segment (= one K-A launch)
      int rc = launch_segment(c, *s);
      if (rc) return rc;
    }
    if (c->log_head >= c->log_limit) {  // out of known-free log frames: look for space closed videos gave back
      int rc = launch_segment(c, *s);    // a segment's frames are contiguous in the log
      if (rc) return rc;
      if (!log_find_space(c))
        return fail(c, MSCAN_ERR_CAPACITY, "frame log full (%llu frames, all owned by open videos); close videos or create a larger context",
                    (unsigned long long)c->log_cap);
    }
    // how many whole frames fit into the current slab (and into the free run of the log)
    uint32_t take = 0;
    uint64_t take_recs = 0;
    const uint64_t log_room = c->log_limit - c->log_head;
    while (f + take < n_frames && s->frames + take < c->slab_frames && take < log_room) {
      const uint64_t nb = (take_recs + rec_count[f + take]) * out_stride;
      if (s->bytes + nb > c->slab_bytes) break;
      if (take && take_recs + rec_count[f + take] > max_take_recs) break;
      take_recs += rec_count[f + take];
      ++take;
    }
    if (take == 0) {
      if (s->frames == 0) {
        return fail(c, MSCAN_ERR_CAPACITY, "frame with %u records exceeds the slab size (%llu bytes)", rec_count[f],
                    (unsigned long long)c->slab_bytes);
      }
      int rc = flush_locked(c);
      if (rc) return rc;
      continue;
    }
    const uint64_t nbytes = take_recs * out_stride;
    if (nbytes && !recs) return fail(c, MSCAN_ERR_INVALID, "null recs with non-zero rec_count");
    if (s->frames == s->seg_frame0) {  // this reservation opens the segment
      s->seg_log_base = c->log_head;
      s->fmt = slab_fmt;
      std::lock_guard<std::mutex> issue(c->issue_mu);
      s->staged = staged;
      s->copy_head = s->seg_byte0;
    }
    if (staged && nbytes && !s->h_recs) {
      CU(use_device(c));
      CU(cudaHostAlloc((void**)&s->h_recs, c->slab_bytes, cudaHostAllocDefault));
    }
    // ---- reserve: byte range, frame slots, log range --------------------------------------------------
    const uint64_t off = s->bytes;
    uint64_t r = s->seg_recs;
    const uint32_t slot = s->seg_slot0 + (s->frames - s->seg_frame0);
    for (uint32_t i = 0; i < take; ++i) {
      s->h_rec_off[slot + i] = r;
      r += rec_count[f + i];
      s->h_geom[s->frames + i] = v.geom;
      s->h_pts[s->frames + i] = pts[f + i];
    }
    // extend the video's last extent when contiguous in the log
    const uint64_t at = c->log_head, vpos = vbase + f;
    if (!v.extents.empty() && v.extents.back().start + v.extents.back().n == at && v.extents.back().vpos + v.extents.back().n == vpos)
      v.extents.back().n += take;
    else v.extents.push_back(Extent{at, take, vpos});
    v.slab_epoch[c->cur] = s->epoch;
    c->log_head += take;
    s->frames += take;
    s->seg_recs += take_recs;
    s->bytes += nbytes;
    const uint8_t* from = src + src_rec * in_stride;
    src_rec += take_recs;
    f += take;
    const bool slab_full = s->frames == c->slab_frames || s->bytes + 64 * 1024 > c->slab_bytes;
    if (nbytes) {
      if (staged)
        for (uint64_t k = off >> c->win_shift; k <= (off + nbytes - 1) >> c->win_shift; ++k)
          s->pend[k * kPendStride].fetch_add(1, std::memory_order_relaxed);
      else s->writers.fetch_add(1, std::memory_order_relaxed);
      s->reserved_end.store(s->bytes, std::memory_order_release);
      // ---- fill + commit outside the mutex -------------------------------------------------------------
      const uint64_t my_epoch = s->epoch;
      lk.unlock();
      const int rc = fill_and_commit(c, *s, kind, off, nbytes, from, take_recs);
      if (rc && !result) result = rc;
      if (f >= n_frames && !slab_full) break;  // common case: done without taking the mutex again
      lock_briefly(lk);
      if (slab_full && &c->slabs[c->cur] == s && s->epoch == my_epoch) {
        int rc2 = flush_locked(c);
        if (rc2) return rc2;
      }
    } else if (slab_full) {
      int rc = flush_locked(c);
      if (rc) return rc;
    }
  }
  if (result == MSCAN_ERR_CUDA) {
    if (!lk.owns_lock()) lk.lock();
    return fail(c, result, "cudaMemcpyAsync of pinned caller records failed");
  }
  return result;
} catch (...) {
  return on_exception(c);
}

int mscan_submit(mscan_ctx* c, uint32_t video_id, uint32_t n_frames, const double* pts, const uint32_t* rec_count,
                 const mscan_mv* recs, uint64_t* first_frame_out) {
  return submit_impl(c, video_id, n_frames, pts, rec_count, recs, false, first_frame_out, nullptr);
}

int mscan_submit_packed(mscan_ctx* c, uint32_t video_id, uint32_t n_frames, const double* pts, const uint32_t* rec_count,
                        const mscan_mv8* recs, uint64_t* first_frame_out) {
  if (recs && (reinterpret_cast<uintptr_t>(recs) & 7u)) return fail(c, MSCAN_ERR_INVALID, "packed records must be 8-byte aligned");
  return submit_impl(c, video_id, n_frames, pts, rec_count, recs, true, first_frame_out, nullptr);
}

// Frames whose records already lie in this GPU's memory: appended to the video's frame log like a host submit and
// scanned IN PLACE by K-A on the context's device stream — no staging, no copy of the records; only the per-frame
// metadata (8 B offset + 8 B pts per frame) crosses PCIe.
int mscan_submit_device(mscan_ctx* c, uint32_t video_id, uint32_t n_frames, const double* pts, const uint32_t* rec_count,
                        const void* d_recs, int packed, void* ready_stream, uint64_t* first_frame_out) try {
  ApiTimer trace_(c, "mscan_submit_device");
  if (!c) return MSCAN_ERR_INVALID;
  if (n_frames && (!pts || !rec_count)) return fail(c, MSCAN_ERR_INVALID, "null pts/rec_count");
  if (reinterpret_cast<uintptr_t>(d_recs) & 15u) return fail(c, MSCAN_ERR_INVALID, "d_recs must be 16-byte aligned");
  std::unique_lock<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->device));
  auto it = c->videos.find(video_id);
  if (it == c->videos.end()) return fail(c, MSCAN_ERR_INVALID, "video %u is not open", video_id);
  Video& v = it->second;
  const uint64_t vbase = v.n_frames;
  if (first_frame_out) *first_frame_out = vbase;
  if (n_frames == 0) return MSCAN_OK;
  uint64_t total = 0;
  for (uint32_t i = 0; i < n_frames; ++i) total += rec_count[i];
  if (total && !d_recs) return fail(c, MSCAN_ERR_INVALID, "null d_recs with non-zero rec_count");
  if (ready_stream) {  // the records become ready on the producer's stream
    CU(cudaEventRecord(c->dev_ready, (cudaStream_t)ready_stream));
    CU(cudaStreamWaitEvent(c->dev_stream, c->dev_ready, 0));
  }
  {  // the frames of an open staged segment are contiguous in the log: close it before this call takes log frames
    int rc = launch_segment(c, c->slabs[c->cur]);
    if (rc) return rc;
  }
  std::vector<uint64_t> off;
  uint32_t f = 0;
  uint64_t rec_at = 0;
  while (f < n_frames) {
    if (c->log_head >= c->log_limit) {
      int rc = launch_segment(c, c->slabs[c->cur]);  // (a staged segment's frames must stay contiguous in the log)
      if (rc) return rc;
      if (!log_find_space(c)) {
        v.n_frames = vbase + f;
        return fail(c, MSCAN_ERR_CAPACITY, "frame log full (%llu frames, all owned by open videos)", (unsigned long long)c->log_cap);
      }
    }
    const uint32_t take = (uint32_t)std::min<uint64_t>(n_frames - f, c->log_limit - c->log_head);
    off.resize((size_t)take + 1);
    uint64_t r = rec_at;
    for (uint32_t i = 0; i < take; ++i) {
      off[i] = r;  // record indices from d_recs: K-A reads frames at any record offset of a 16-byte aligned base
      r += rec_count[f + i];
    }
    off[take] = r;
    const uint64_t at = c->log_head;
    uint64_t* d_off = nullptr;
    CU(cudaMallocAsync((void**)&d_off, sizeof(uint64_t) * ((size_t)take + 1), c->dev_stream));
    // pageable sources: both copies are staged by the runtime before the calls return
    CU(cudaMemcpyAsync(d_off, off.data(), sizeof(uint64_t) * ((size_t)take + 1), cudaMemcpyHostToDevice, c->dev_stream));
    CU(cudaMemcpyAsync(c->d_pts + at, pts + f, sizeof(double) * take, cudaMemcpyHostToDevice, c->dev_stream));
    c->a_h2d_bytes.fetch_add(16ull * take + 8, std::memory_order_relaxed);
    ScanArgs a = base_args(c);
    a.recs = reinterpret_cast<const uint8_t*>(d_recs);
    a.packed = packed ? 1u : 0u;
    a.rec_off = d_off;
    a.frame_geom = nullptr;            // every frame of the call has the video's geometry:
    a.geoms = c->d_geoms + v.geom;     // index 0 of a table that starts at it
    a.flags = c->d_flags + at;
    a.counts = c->d_counts + at;
    a.n_frames = take;
    const ScanPlan& plan = packed ? c->plan_packed : c->plan;
    a.stages = plan.stages;
    a.max_cells = plan.cells;
    a.max_bit_words = plan.bit_words;
    int rc = run_scan(c, a, plan, c->dev_stream, r - rec_at, kSlabs);
    CU(cudaFreeAsync(d_off, c->dev_stream));
    if (rc) {
      v.n_frames = vbase + f;
      return rc;
    }
    CU(cudaEventRecord(c->dev_done, c->dev_stream));
    c->dev_seq += 1;
    v.dev_seq = c->dev_seq;
    const uint64_t vpos = vbase + f;
    if (!v.extents.empty() && v.extents.back().start + v.extents.back().n == at && v.extents.back().vpos + v.extents.back().n == vpos)
      v.extents.back().n += take;
    else v.extents.push_back(Extent{at, take, vpos});
    v.n_frames = vpos + take;
    c->log_head += take;
    rec_at = r;
    f += take;
  }
  return MSCAN_OK;
} catch (...) {
  return on_exception(c);
}

int mscan_device_pci_bus_id(int device, char* buf, int len) {
  if (!buf || len < 13) return MSCAN_ERR_INVALID;
  if (cudaDeviceGetPCIBusId(buf, len, device) != cudaSuccess) {
    cudaGetLastError();
    buf[0] = 0;
    return MSCAN_ERR_CUDA;
  }
  return MSCAN_OK;
}

int mscan_pack_records(const mscan_mv* recs, uint64_t n, mscan_mv8* out) {
  if (n && (!recs || !out)) return MSCAN_ERR_INVALID;
  if (reinterpret_cast<uintptr_t>(out) & 7u) return MSCAN_ERR_INVALID;
  project_records(reinterpret_cast<const uint8_t*>(recs), n, reinterpret_cast<uint64_t*>(out));
  return MSCAN_OK;
}

int mscan_compact_records(const mscan_mv* recs, uint64_t n, mscan_mv8* out, uint64_t* n_out) {
  if ((n && (!recs || !out)) || !n_out) return MSCAN_ERR_INVALID;
  if (reinterpret_cast<uintptr_t>(out) & 7u) return MSCAN_ERR_INVALID;
  *n_out = n ? compact_moving(reinterpret_cast<const uint8_t*>(recs), n, reinterpret_cast<uint64_t*>(out)) : 0;
  return MSCAN_OK;
}

int mscan_submit_elided(mscan_ctx* c, uint32_t video_id, uint32_t n_frames, const double* pts, const uint32_t* rec_count,
                        const void* enc, const uint64_t* enc_off, const uint32_t* tile_end16, uint64_t* first_frame_out) try {
  ApiTimer trace_(c, "mscan_submit_elided");
  if (!c) return MSCAN_ERR_INVALID;
  if (n_frames == 0) {
    if (first_frame_out) {
      std::lock_guard<std::mutex> lk(c->mu);
      auto it0 = c->videos.find(video_id);
      *first_frame_out = it0 == c->videos.end() ? 0 : it0->second.n_frames;
    }
    return MSCAN_OK;
  }
  if (!pts || !rec_count || !enc_off) return fail(c, MSCAN_ERR_INVALID, "null pts/rec_count/enc_off");
  const uint8_t* data = static_cast<const uint8_t*>(enc);
  const uint64_t total_bytes = enc_off[n_frames] - enc_off[0];
  if (total_bytes && (!enc || !tile_end16)) return fail(c, MSCAN_ERR_INVALID, "null enc/tile_end16 with non-empty frames");
  if ((reinterpret_cast<uintptr_t>(data) + enc_off[0]) & 15u) return fail(c, MSCAN_ERR_INVALID, "elided frames must be 16-byte aligned");
  {
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->videos.count(video_id)) return fail(c, MSCAN_ERR_INVALID, "video %u is not open", video_id);
    if (c->plan_packed.cluster || c->plan_packed.global_cnt)
      return fail(c, MSCAN_ERR_UNSUPPORTED, "the static-elided form is not read by the cluster kernel (grids beyond one CTA's shared memory)");
  }
  // piece-relative tables (one piece = the whole call)
  std::vector<uint64_t> frame_end((size_t)n_frames + 1);
  std::vector<uint32_t> frame_tile((size_t)n_frames + 1), tile_end;
  uint32_t nt = 0;
  frame_end[0] = 0;
  for (uint32_t i = 0; i < n_frames; ++i) {
    if (enc_off[i + 1] < enc_off[i] || ((enc_off[i] - enc_off[0]) & 15u)) return fail(c, MSCAN_ERR_INVALID, "enc_off must be non-decreasing multiples of 16");
    const uint32_t tiles = (rec_count[i] + kMvzTileRecs - 1) / kMvzTileRecs;
    const uint64_t rel = enc_off[i] - enc_off[0], bytes = enc_off[i + 1] - enc_off[i];
    frame_tile[i] = nt;
    for (uint32_t t = 0; t < tiles; ++t) {
      const uint64_t end = (uint64_t)tile_end16[nt + t] << 4;  // from the start of the frame's encoding
      if (end > bytes || (t && tile_end16[nt + t] < tile_end16[nt + t - 1])) return fail(c, MSCAN_ERR_INVALID, "tile ends of frame %u do not fit its encoding", i);
      tile_end.push_back((uint32_t)((rel + end) >> 4));
    }
    if (tiles && ((uint64_t)tile_end16[nt + tiles - 1] << 4) != bytes) return fail(c, MSCAN_ERR_INVALID, "frame %u: the last tile must end at the frame's end", i);
    if (!tiles && bytes) return fail(c, MSCAN_ERR_INVALID, "frame %u has no records but %llu encoded bytes", i, (unsigned long long)bytes);
    nt += tiles;
    frame_end[(size_t)i + 1] = rel + bytes;
  }
  frame_tile[n_frames] = nt;
  bool pinned = false;
  if (total_bytes) {
    pinned = pinned_ranges().contains(data + enc_off[0], total_bytes);
    if (!pinned && total_bytes >= kAttrQueryBytes) pinned = is_pinned(data + enc_off[0]);
  }
  uint64_t recs = 0;
  for (uint32_t i = 0; i < n_frames; ++i) recs += rec_count[i];
  c->a_records_elided.fetch_add(recs, std::memory_order_relaxed);
  c->a_elided_bytes.fetch_add(total_bytes, std::memory_order_relaxed);
  ElidedPlacer placer(c, video_id, n_frames, first_frame_out);
  return placer.place(0, n_frames, pts, rec_count, data + enc_off[0], frame_end.data(), frame_tile.data(), tile_end.data(), pinned);
} catch (...) {
  return on_exception(c);
}

int mscan_elide_records(const mscan_mv* recs, uint32_t n, void* out, size_t cap, uint32_t* tile_end16, uint32_t tile_cap,
                        size_t* bytes_out) {
  if ((n && !recs) || !out || !bytes_out) return MSCAN_ERR_INVALID;
  if (reinterpret_cast<uintptr_t>(out) & 15u) return MSCAN_ERR_INVALID;
  const uint32_t tiles = (n + kMvzTileRecs - 1) / kMvzTileRecs;
  if (cap < mvz_bound(n, 1) || (tiles && (!tile_end16 || tile_cap < tiles))) return MSCAN_ERR_CAPACITY;
  *bytes_out = (size_t)mvz_encode_frame(reinterpret_cast<const uint8_t*>(recs), n, static_cast<uint8_t*>(out), tile_end16);
  return MSCAN_OK;
}

size_t mscan_elide_bound(uint32_t n) { return (size_t)mvz_bound(n, 1); }

int mscan_set_staging_mode(mscan_ctx* c, int mode) {
  if (!c) return MSCAN_ERR_INVALID;
  if (mode != MSCAN_STAGING_AUTO && mode != MSCAN_STAGING_PACK && mode != MSCAN_STAGING_NATIVE && mode != MSCAN_STAGING_ELIDE &&
      mode != MSCAN_STAGING_COMPACT)
    return fail(c, MSCAN_ERR_INVALID, "unknown staging mode %d", mode);
  std::lock_guard<std::mutex> lk(c->mu);
  c->staging_mode.store(mode, std::memory_order_relaxed);
  return MSCAN_OK;
}

// Allocates the pinned staging of every slab now instead of at the first staged submit that needs it. Pinning tens of
// MiB takes the process's address-space lock for ~0.4 ms per MiB: paid lazily it stalls every thread that is faulting in
// file pages at that moment (8 contexts x 3 slabs x 64 MiB on an 8-GPU box cost a 64-clip batch 0.6 s of its 1.3 s);
// a host that knows it will feed pageable records calls this right after mscan_create, before its workers start.
int mscan_reserve_staging(mscan_ctx* c) try {
  if (!c) return MSCAN_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->device));
  for (auto& s : c->slabs)
    if (!s.h_recs) CU(cudaHostAlloc((void**)&s.h_recs, c->slab_bytes, cudaHostAllocDefault));
  return MSCAN_OK;
} catch (...) {
  return on_exception(c);
}

int mscan_set_pack_threads(mscan_ctx* c, int n_threads) try {
  if (!c || n_threads < 0) return MSCAN_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  c->pack_threads = n_threads;  // the pool itself is shared by every context of the process
  return MSCAN_OK;
} catch (...) {
  return on_exception(c);
}

int mscan_flush(mscan_ctx* c) try {
  ApiTimer trace_(c, "mscan_flush");
  if (!c) return MSCAN_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->device));
  return flush_locked(c);
} catch (...) {
  return on_exception(c);
}

// Copies pieces of the frame log (flags / counts) to the caller's arrays. A video fed frame by frame by many decode
// threads next to other videos owns thousands of short extents: instead of two tiny D2H copies per extent (≈ 6 µs each —
// 74 ms for three interleaved videos of 2 000 frames) the covering range of the log comes over in one copy and is
// scattered on the host, as long as that range is not much larger than what is asked for.
namespace {
struct LogPiece {
  uint64_t src, n, dst;
};
int copy_log_pieces(mscan_ctx* c, const std::vector<LogPiece>& pieces, uint8_t* flags, uint32_t* counts) {
  if (pieces.empty() || (!flags && !counts)) return MSCAN_OK;
  uint64_t lo = ~0ull, hi = 0, need = 0;
  for (const LogPiece& p : pieces) {
    lo = std::min(lo, p.src);
    hi = std::max(hi, p.src + p.n);
    need += p.n;
  }
  const uint64_t span = hi - lo;
  if (pieces.size() > 8 && span <= 4 * need + (1u << 16)) {
    std::vector<uint8_t> tf;
    std::vector<uint32_t> tc;
    if (flags) {
      tf.resize(span);
      CU(cudaMemcpyAsync(tf.data(), c->d_flags + lo, span, cudaMemcpyDeviceToHost, c->main_stream));
    }
    if (counts) {
      tc.resize(span);
      CU(cudaMemcpyAsync(tc.data(), c->d_counts + lo, span * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->main_stream));
    }
    CU(cudaStreamSynchronize(c->main_stream));
    for (const LogPiece& p : pieces) {
      if (flags) std::memcpy(flags + p.dst, tf.data() + (p.src - lo), p.n);
      if (counts) std::memcpy(counts + p.dst, tc.data() + (p.src - lo), p.n * sizeof(uint32_t));
    }
    c->stats.d2h_bytes += (flags ? span : 0) + (counts ? 4 * span : 0);
    return MSCAN_OK;
  }
  for (const LogPiece& p : pieces) {
    if (flags) CU(cudaMemcpyAsync(flags + p.dst, c->d_flags + p.src, p.n, cudaMemcpyDeviceToHost, c->main_stream));
    if (counts) CU(cudaMemcpyAsync(counts + p.dst, c->d_counts + p.src, p.n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->main_stream));
    c->stats.d2h_bytes += (flags ? p.n : 0) + (counts ? 4 * p.n : 0);
  }
  CU(cudaStreamSynchronize(c->main_stream));
  return MSCAN_OK;
}
}  // namespace

int mscan_collect(mscan_ctx* c, uint32_t video_id, uint8_t* flags, uint32_t* counts, uint32_t cap, uint32_t* n_out) try {
  ApiTimer trace_(c, "mscan_collect");
  if (!c) return MSCAN_ERR_INVALID;
  std::unique_lock<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->device));
  auto it = c->videos.find(video_id);
  if (it == c->videos.end()) return fail(c, MSCAN_ERR_INVALID, "video %u is not open", video_id);
  if (n_out) *n_out = (uint32_t)it->second.n_frames;
  if (it->second.n_frames > cap && (flags || counts))
    return fail(c, MSCAN_ERR_CAPACITY, "need room for %llu frames", (unsigned long long)it->second.n_frames);
  if (!flags && !counts) return MSCAN_OK;  // size query
  int rc = sync_videos(c, lk, &video_id, 1);
  if (rc) return rc;
  it = c->videos.find(video_id);
  if (it == c->videos.end()) return fail(c, MSCAN_ERR_INVALID, "video %u was closed during the call", video_id);
  const Video& v = it->second;
  std::vector<LogPiece> pieces;
  pieces.reserve(v.extents.size());
  for (const Extent& e : v.extents) {
    if (e.vpos + e.n > cap) return fail(c, MSCAN_ERR_CAPACITY, "need room for %llu frames", (unsigned long long)v.n_frames);
    pieces.push_back(LogPiece{e.start, e.n, e.vpos});
  }
  return copy_log_pieces(c, pieces, flags, counts);
} catch (...) {
  return on_exception(c);
}

int mscan_collect_range(mscan_ctx* c, uint32_t video_id, uint64_t first, uint32_t n, uint8_t* flags, uint32_t* counts) try {
  ApiTimer trace_(c, "mscan_collect_range");
  if (!c) return MSCAN_ERR_INVALID;
  std::unique_lock<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->device));
  auto it = c->videos.find(video_id);
  if (it == c->videos.end()) return fail(c, MSCAN_ERR_INVALID, "video %u is not open", video_id);
  if (first + n > it->second.n_frames)
    return fail(c, MSCAN_ERR_INVALID, "range [%llu,+%u) exceeds the video's %llu frames", (unsigned long long)first, n,
                (unsigned long long)it->second.n_frames);
  int rc = sync_videos(c, lk, &video_id, 1);
  if (rc) return rc;
  it = c->videos.find(video_id);
  if (it == c->videos.end()) return fail(c, MSCAN_ERR_INVALID, "video %u was closed during the call", video_id);
  const Video& v = it->second;
  std::vector<LogPiece> pieces;
  for (const Extent& e : v.extents) {
    const uint64_t a = std::max<uint64_t>(first, e.vpos), b = std::min<uint64_t>(first + n, e.vpos + e.n);
    if (a < b) pieces.push_back(LogPiece{e.start + (a - e.vpos), b - a, a - first});
  }
  return copy_log_pieces(c, pieces, flags, counts);
} catch (...) {
  return on_exception(c);
}

// Runs K-C for a list of open videos; leaves results in c->h_res and segments on the device. Caller holds tail_mu
// (K-C scratch, main_stream) and `lk` = mu; mu is released while the scans of these videos finish and while K-C runs.
static int run_segments_locked(mscan_ctx* c, std::unique_lock<std::mutex>& lk, uint32_t n_videos, const uint32_t* ids,
                               const double* durations, std::vector<uint64_t>* seg_base_out) {
  for (uint32_t i = 0; i < n_videos; ++i)
    if (!c->videos.count(ids[i])) return fail(c, MSCAN_ERR_INVALID, "video %u is not open", ids[i]);
  int rc = sync_videos(c, lk, ids, n_videos);
  if (rc) return rc;
  uint64_t n_ext = 0, ts_total = 0, seg_total = 0;
  for (uint32_t i = 0; i < n_videos; ++i) {
    auto it = c->videos.find(ids[i]);
    if (it == c->videos.end()) return fail(c, MSCAN_ERR_INVALID, "video %u is not open", ids[i]);
    n_ext += it->second.extents.size();
    ts_total += pow2_ge(std::max<uint64_t>(it->second.n_frames, 1));
    seg_total += std::max<uint64_t>(it->second.n_frames, 1);
  }
  if (n_videos > c->job_cap) {
    if (c->h_jobs) cudaFreeHost(c->h_jobs);
    if (c->h_res) cudaFreeHost(c->h_res);
    cudaFree(c->d_jobs);
    cudaFree(c->d_res);
    c->job_cap = 0;
    const uint32_t n = std::max(n_videos, 64u);
    CU(cudaHostAlloc((void**)&c->h_jobs, sizeof(SegJob) * n, cudaHostAllocDefault));
    CU(cudaHostAlloc((void**)&c->h_res, sizeof(mscan_video_result) * n, cudaHostAllocDefault));
    CU(cudaMalloc((void**)&c->d_jobs, sizeof(SegJob) * n));
    CU(cudaMalloc((void**)&c->d_res, sizeof(mscan_video_result) * n));
    c->job_cap = n;
  }
  if (n_ext > c->ext_cap) {
    if (c->h_exts) cudaFreeHost(c->h_exts);
    cudaFree(c->d_exts);
    c->ext_cap = 0;
    const uint32_t n = (uint32_t)std::max<uint64_t>(n_ext, 256);
    CU(cudaHostAlloc((void**)&c->h_exts, sizeof(SegExtent) * n, cudaHostAllocDefault));
    CU(cudaMalloc((void**)&c->d_exts, sizeof(SegExtent) * n));
    c->ext_cap = n;
  }
  {
    uint64_t cap = c->ts_cap;
    rc = grow(c, &c->d_ts_a, &cap, ts_total);
    if (rc) return rc;
    uint64_t cap_b = c->ts_cap;
    rc = grow(c, &c->d_ts_b, &cap_b, ts_total);
    if (rc) return rc;
    c->ts_cap = std::min(cap, cap_b);
    rc = grow(c, &c->d_segs, &c->seg_cap, seg_total);
    if (rc) return rc;
  }
  c->dev_jobs_cached.clear();
  seg_base_out->resize(n_videos);
  uint64_t e = 0, ts_at = 0, seg_at = 0;
  std::vector<Extent> ordered;
  for (uint32_t i = 0; i < n_videos; ++i) {
    const Video& v = c->videos.find(ids[i])->second;
    SegJob j{};
    j.ext_begin = (uint32_t)e;
    // in submission order: chunks that arrived in order then compact to an increasing list and K-C skips its sort
    ordered = v.extents;
    std::sort(ordered.begin(), ordered.end(), [](const Extent& x, const Extent& y) { return x.vpos < y.vpos; });
    for (const Extent& x : ordered) c->h_exts[e++] = SegExtent{x.start, x.n};
    j.ext_end = (uint32_t)e;
    j.ts_base = ts_at;
    j.ts_cap = pow2_ge(std::max<uint64_t>(v.n_frames, 1));
    j.seg_base = seg_at;
    j.duration = durations[i];
    (*seg_base_out)[i] = seg_at;
    ts_at += j.ts_cap;
    seg_at += std::max<uint64_t>(v.n_frames, 1);
    c->h_jobs[i] = j;
  }
  const bool profiling = c->profiling;
  EvPair ev{};
  if (profiling) ev = get_events(c, 1);
  c->stats.h2d_bytes += sizeof(SegJob) * n_videos + sizeof(SegExtent) * n_ext;
  c->stats.segment_launches += 1;
  c->stats.d2h_bytes += sizeof(mscan_video_result) * n_videos;
  // K-C reads the videos' (finished) frame-log extents and its own scratch: the bookkeeping mutex is not needed
  lk.unlock();
  cudaStream_t st = c->main_stream;
  cudaError_t ce = cudaMemcpyAsync(c->d_jobs, c->h_jobs, sizeof(SegJob) * n_videos, cudaMemcpyHostToDevice, st);
  if (ce == cudaSuccess && n_ext) ce = cudaMemcpyAsync(c->d_exts, c->h_exts, sizeof(SegExtent) * n_ext, cudaMemcpyHostToDevice, st);
  SegArgs a{};
  a.jobs = c->d_jobs;
  a.extents = c->d_exts;
  a.pts = c->d_pts;
  a.flags = c->d_flags;
  a.ts_a = c->d_ts_a;
  a.ts_b = c->d_ts_b;
  a.segs = c->d_segs;
  a.results = c->d_res;
  a.max_gap = c->params.max_gap_sec;
  a.padding = c->params.padding_sec;
  a.min_savings_pct = c->params.min_savings_pct;
  if (profiling) cudaEventRecord(ev.a, st);
  if (ce == cudaSuccess) ce = segments_launch(a, n_videos, st);
  if (profiling) cudaEventRecord(ev.b, st);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(c->h_res, c->d_res, sizeof(mscan_video_result) * n_videos, cudaMemcpyDeviceToHost, st);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
  lk.lock();
  if (profiling) c->ev_pending.push_back(ev);
  if (ce != cudaSuccess) return fail(c, MSCAN_ERR_CUDA, "K-C launch failed: %s", cudaGetErrorString(ce));
  return MSCAN_OK;
}

static int segments_impl(mscan_ctx* c, uint32_t n_videos, const uint32_t* ids, const double* durations,
                         mscan_segment* out, uint64_t cap, uint64_t* seg_off_out, mscan_video_result* res_out,
                         bool job_semantics) try {
  ApiTimer trace_(c, "mscan_segments[_batch]");
  if (!c || (n_videos && (!ids || !durations))) return MSCAN_ERR_INVALID;
  std::lock_guard<std::mutex> tail(c->tail_mu);
  std::unique_lock<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->device));
  if (n_videos == 0) {
    if (seg_off_out) seg_off_out[0] = 0;
    return MSCAN_OK;
  }
  std::vector<uint64_t> seg_base;
  int rc = run_segments_locked(c, lk, n_videos, ids, durations, &seg_base);
  if (rc) return rc;
  uint64_t at = 0;
  bool overflow = false;
  for (uint32_t i = 0; i < n_videos; ++i) {
    const mscan_video_result& r = c->h_res[i];
    if (res_out) res_out[i] = r;
    if (seg_off_out) seg_off_out[i] = at;
    uint64_t n = r.n_segments;
    const bool full_copy = job_semantics && r.decision == MSCAN_FULL_COPY;
    if (job_semantics && r.decision == MSCAN_NO_MOTION) n = 0;
    if (full_copy) n = 1;
    if (at + n > cap || (!out && n)) {
      overflow = true;
    } else if (full_copy) {
      out[at] = mscan_segment{0.0, durations[i]};  // pipeline.cpp:386-387
    } else if (n) {
      CU(cudaMemcpyAsync(out + at, c->d_segs + seg_base[i], sizeof(mscan_segment) * n, cudaMemcpyDeviceToHost, c->main_stream));
      c->stats.d2h_bytes += sizeof(mscan_segment) * n;
    }
    at += n;
  }
  if (seg_off_out) seg_off_out[n_videos] = at;
  CU(cudaStreamSynchronize(c->main_stream));
  if (overflow) return fail(c, MSCAN_ERR_CAPACITY, "segment buffer too small: need %llu", (unsigned long long)at);
  return MSCAN_OK;
} catch (...) {
  return on_exception(c);
}

int mscan_segments_batch(mscan_ctx* c, uint32_t n_videos, const uint32_t* ids, const double* durations,
                         mscan_segment* out, uint64_t cap, uint64_t* seg_off_out, mscan_video_result* res_out) {
  return segments_impl(c, n_videos, ids, durations, out, cap, seg_off_out, res_out, true);
}

int mscan_segments(mscan_ctx* c, uint32_t video_id, double duration, mscan_segment* out, uint32_t cap, uint32_t* n_out,
                   mscan_video_result* res_out) {
  uint64_t off[2] = {0, 0};
  int rc = segments_impl(c, 1, &video_id, &duration, out, cap, off, res_out, true);
  if (n_out) *n_out = (uint32_t)off[1];
  return rc;
}

int mscan_motion_segments(mscan_ctx* c, uint32_t video_id, double duration, mscan_segment* out, uint32_t cap,
                          uint32_t* n_out, mscan_video_result* res_out) {
  uint64_t off[2] = {0, 0};
  int rc = segments_impl(c, 1, &video_id, &duration, out, cap, off, res_out, false);
  if (n_out) *n_out = (uint32_t)off[1];
  return rc;
}

int mscan_video_close(mscan_ctx* c, uint32_t video_id) try {
  ApiTimer trace_(c, "mscan_video_close");
  if (!c) return MSCAN_ERR_INVALID;
  std::unique_lock<std::mutex> lk(c->mu);
  if (!c->videos.count(video_id)) return fail(c, MSCAN_ERR_INVALID, "video %u is not open", video_id);
  // the video's log frames become reusable: no scan in flight may still write them (only this video's scans are
  // waited for — the other streams of the GPU keep running)
  CU(cudaSetDevice(c->device));
  int rc = sync_videos(c, lk, &video_id, 1);
  if (rc) return rc;
  auto it = c->videos.find(video_id);
  if (it == c->videos.end()) return MSCAN_OK;  // closed by another thread while we waited
  fold_video_stats(c, it->second);
  c->videos.erase(it);
  if (c->videos.empty()) {  // nothing live any more: rewind
    c->log_head = 0;
    c->log_limit = c->log_cap;
  }
  return MSCAN_OK;
} catch (...) {
  return on_exception(c);
}

// Cross-GPU stitch: the frames another context scanned for the same video join this
// context's log by a peer copy of their 13 B/frame (NVLink when the GPUs are peers), so that K-C sees the
// whole video. The union/sort/unique of chunk results (pipeline.cpp:268,302-304) then happens in K-C as usual.
int mscan_video_append_from(mscan_ctx* dst, uint32_t dst_video, mscan_ctx* src, uint32_t src_video) try {
  if (!dst || !src) return MSCAN_ERR_INVALID;
  mscan_ctx* c = dst;  // errors are reported on the destination context
  if (dst == src && dst_video == src_video) return fail(c, MSCAN_ERR_INVALID, "cannot append a video to itself");
  std::unique_lock<std::mutex> l1(dst->mu, std::defer_lock), l2(src->mu, std::defer_lock);
  if (dst == src) l1.lock();
  else std::lock(l1, l2);
  auto sit = src->videos.find(src_video);
  if (sit == src->videos.end()) return fail(c, MSCAN_ERR_INVALID, "source video %u is not open", src_video);
  auto dit = dst->videos.find(dst_video);
  if (dit == dst->videos.end()) return fail(c, MSCAN_ERR_INVALID, "video %u is not open", dst_video);
  const Video& sv = sit->second;
  Video& dv = dit->second;
  if (sv.n_frames == 0) return MSCAN_OK;
  // the source's results must exist; the destination's open segment must not straddle the imported range
  if (cudaSetDevice(src->device) != cudaSuccess) return fail(c, MSCAN_ERR_CUDA, "cudaSetDevice(%d) failed", src->device);
  {
    mscan_ctx* c = src;  // CU() reports on `c`
    int rc = sync_scans_locked(c);
    if (rc) return fail(dst, rc, "source context: %s", src->err.c_str());
  }
  CU(cudaSetDevice(dst->device));
  if (dst != src) {
    int rc = launch_segment(dst, dst->slabs[dst->cur]);
    if (rc) return rc;
    if (dst->device != src->device) {  // best effort: direct NVLink path for the peer copies below
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, dst->device, src->device) == cudaSuccess && can) cudaDeviceEnablePeerAccess(src->device, 0);
      cudaGetLastError();  // "already enabled" is fine
    }
  } else {
    int rc = sync_scans_locked(dst);
    if (rc) return rc;
  }
  const std::vector<Extent> src_extents = sv.extents;  // dst == src: dv.extents grows below
  const uint64_t dv_base = dv.n_frames;                 // the adopted frames keep their order behind the video's own
  const uint64_t sv_frames = sv.n_frames;
  uint64_t done_frames = 0;
  for (const Extent& e : src_extents) {
    uint64_t off = 0;
    while (off < e.n) {  // in pieces, as the free runs of the destination log allow
      if (dst->log_head >= dst->log_limit && !log_find_space(dst)) {
        CU(cudaStreamSynchronize(dst->main_stream));
        return fail(c, MSCAN_ERR_CAPACITY, "frame log full (%llu frames, all owned by open videos)", (unsigned long long)dst->log_cap);
      }
      const uint64_t m = std::min(e.n - off, dst->log_limit - dst->log_head), at = dst->log_head, from = e.start + off;
      CU(cudaMemcpyPeerAsync(dst->d_pts + at, dst->device, src->d_pts + from, src->device, sizeof(double) * m, dst->main_stream));
      CU(cudaMemcpyPeerAsync(dst->d_flags + at, dst->device, src->d_flags + from, src->device, m, dst->main_stream));
      CU(cudaMemcpyPeerAsync(dst->d_counts + at, dst->device, src->d_counts + from, src->device, sizeof(uint32_t) * m, dst->main_stream));
      const uint64_t vpos = dv_base + e.vpos + off;
      if (!dv.extents.empty() && dv.extents.back().start + dv.extents.back().n == at && dv.extents.back().vpos + dv.extents.back().n == vpos)
        dv.extents.back().n += m;
      else dv.extents.push_back(Extent{at, m, vpos});
      dst->log_head += m;
      off += m;
      done_frames += m;
    }
  }
  dv.n_frames = dv_base + sv_frames;
  CU(cudaStreamSynchronize(dst->main_stream));
  dst->stats.peer_bytes += 13ull * done_frames;
  return MSCAN_OK;
} catch (...) {
  return on_exception(dst);
}

int mscan_host_alloc(mscan_ctx* c, size_t bytes, void** p) {
  if (!c || !p) return MSCAN_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  if (cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    return fail(c, MSCAN_ERR_NOMEM, "cudaHostAlloc(%zu) failed", bytes);
  }
  pinned_ranges().add(*p, bytes ? bytes : 1);
  return MSCAN_OK;
}

int mscan_host_free(mscan_ctx* c, void* p) {
  if (!c) return MSCAN_ERR_INVALID;
  if (p) {
    pinned_ranges().remove(p);
    CU(cudaFreeHost(p));
  }
  return MSCAN_OK;
}

int mscan_host_register(mscan_ctx* c, void* p, size_t bytes, int read_only) {
  ApiTimer trace_(c, "mscan_host_register");
  if (!c || !p || !bytes) return MSCAN_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  unsigned flags = cudaHostRegisterPortable;
  if (read_only) flags |= cudaHostRegisterReadOnly;
  const cudaError_t e = cudaHostRegister(p, bytes, flags);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(c, MSCAN_ERR_CUDA, "cudaHostRegister(%zu bytes%s) refused: %s", bytes, read_only ? ", read-only" : "",
                cudaGetErrorString(e));
  }
  pinned_ranges().add(p, bytes);
  return MSCAN_OK;
}

int mscan_host_unregister(mscan_ctx* c, void* p) {
  if (!c || !p) return MSCAN_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  pinned_ranges().remove(p);
  CU(cudaHostUnregister(p));
  return MSCAN_OK;
}

int mscan_host_fence(mscan_ctx* c) try {
  ApiTimer trace_(c, "mscan_host_fence");
  if (!c) return MSCAN_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->device));
  int rc = flush_locked(c);  // the open slab's copies get their `copied` event at launch
  if (rc) return rc;
  for (auto& s : c->slabs)
    if (s.in_flight) CU(cudaEventSynchronize(s.copied));
  return MSCAN_OK;
} catch (...) {
  return on_exception(c);
}

// ---- device-resident path -----------------------------------------------------------------------
int mscan_dev_alloc(mscan_ctx* c, size_t bytes, void** d) {
  if (!c || !d) return MSCAN_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  if (cudaMalloc(d, bytes ? bytes : 1) != cudaSuccess) {
    cudaGetLastError();
    return fail(c, MSCAN_ERR_NOMEM, "cudaMalloc(%zu) failed", bytes);
  }
  return MSCAN_OK;
}

int mscan_dev_free(mscan_ctx* c, void* d) {
  if (!c) return MSCAN_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  if (d) CU(cudaFree(d));
  return MSCAN_OK;
}

int mscan_memcpy_h2d(mscan_ctx* c, void* d, const void* h, size_t bytes) {
  if (!c) return MSCAN_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice));
  return MSCAN_OK;
}

int mscan_memcpy_d2h(mscan_ctx* c, void* h, const void* d, size_t bytes) {
  if (!c) return MSCAN_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost));
  return MSCAN_OK;
}

int mscan_offsets_from_counts(mscan_ctx* c, const uint32_t* d_cnt, uint32_t n, uint64_t* d_off, void* stream) {
  if (!c || !d_off || (n && !d_cnt)) return MSCAN_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : c->main_stream;
  CU(offsets_launch(d_cnt, n, d_off, nullptr, st));
  c->stats.aux_launches += 1;
  return MSCAN_OK;
}

static int scan_device_impl(mscan_ctx* c, const void* d_recs, bool packed, const uint64_t* d_rec_off,
                            const uint32_t* d_frame_geom, const mscan_geometry* geoms, uint32_t n_geoms, uint32_t n_frames,
                            uint8_t* d_flags, uint32_t* d_counts, void* stream) try {
  if (!c || !d_rec_off || !geoms || n_geoms == 0 || !d_flags || !d_counts) return MSCAN_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(d_recs) & 15u) != 0) return fail(c, MSCAN_ERR_INVALID, "d_recs must be 16-byte aligned");
  if (n_geoms > kMaxGeoms) return fail(c, MSCAN_ERR_CAPACITY, "too many geometries");
  std::lock_guard<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : c->main_stream;
  std::vector<DevGeom> dg(n_geoms);
  uint32_t cells = 0, words = 0;
  for (uint32_t i = 0; i < n_geoms; ++i) {
    if (!geom_ok(geoms[i])) return fail(c, MSCAN_ERR_INVALID, "bad geometry %u", i);
    dg[i] = to_dev_geom(geoms[i]);
    uint32_t ce, wo;
    geom_need(dg[i], &ce, &wo);
    cells = std::max(cells, ce);
    words = std::max(words, wo);
  }
  ScanPlan plan;
  if (!scan_plan_for(dg.data(), n_geoms, c->smem_optin, packed, &plan))
    return fail(c, MSCAN_ERR_UNSUPPORTED, "block grid of %u cells does not fit shared memory", cells);
  if (n_geoms > c->user_geoms_cap) {
    cudaFree(c->d_user_geoms);
    c->user_geoms_cap = 0;
    c->user_geoms_cached.clear();
    CU(cudaMalloc((void**)&c->d_user_geoms, sizeof(DevGeom) * std::max(n_geoms, 16u)));
    c->user_geoms_cap = std::max(n_geoms, 16u);
  }
  // upload only when the table changed: a pageable-source copy would serialise the stream
  if (c->user_geoms_cached.size() != n_geoms ||
      std::memcmp(c->user_geoms_cached.data(), dg.data(), sizeof(DevGeom) * n_geoms) != 0) {
    // earlier launches on OTHER caller streams may still read the old table: drain the device before rewriting it
    // (the table only changes when a caller switches geometry sets, so this is rare)
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(c->d_user_geoms, dg.data(), sizeof(DevGeom) * n_geoms, cudaMemcpyHostToDevice));
    c->user_geoms_cached = dg;
  }
  ScanArgs a = base_args(c);
  a.recs = reinterpret_cast<const uint8_t*>(d_recs);
  a.packed = packed ? 1u : 0u;
  a.rec_off = d_rec_off;
  a.frame_geom = d_frame_geom;
  a.geoms = c->d_user_geoms;
  a.flags = d_flags;
  a.counts = d_counts;
  a.n_frames = n_frames;
  a.stages = plan.stages;
  a.max_cells = plan.cells;
  a.max_bit_words = plan.bit_words;
  return run_scan(c, a, plan, st, 0);
} catch (...) {
  return on_exception(c);
}

int mscan_scan_device(mscan_ctx* c, const mscan_mv* d_recs, const uint64_t* d_rec_off, const uint32_t* d_frame_geom,
                      const mscan_geometry* geoms, uint32_t n_geoms, uint32_t n_frames, uint8_t* d_flags,
                      uint32_t* d_counts, void* stream) {
  return scan_device_impl(c, d_recs, false, d_rec_off, d_frame_geom, geoms, n_geoms, n_frames, d_flags, d_counts, stream);
}

int mscan_scan_device_packed(mscan_ctx* c, const mscan_mv8* d_recs, const uint64_t* d_rec_off, const uint32_t* d_frame_geom,
                             const mscan_geometry* geoms, uint32_t n_geoms, uint32_t n_frames, uint8_t* d_flags,
                             uint32_t* d_counts, void* stream) {
  return scan_device_impl(c, d_recs, true, d_rec_off, d_frame_geom, geoms, n_geoms, n_frames, d_flags, d_counts, stream);
}

int mscan_pack_records_device(mscan_ctx* c, const mscan_mv* d_recs, uint64_t n, mscan_mv8* d_out, void* stream) {
  if (!c || (n && (!d_recs || !d_out))) return MSCAN_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(d_recs) & 7u) || (reinterpret_cast<uintptr_t>(d_out) & 7u))
    return fail(c, MSCAN_ERR_INVALID, "device record buffers must be 8-byte aligned");
  std::lock_guard<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->device));
  CU(project_launch(d_recs, n, d_out, stream ? (cudaStream_t)stream : c->main_stream));
  c->stats.aux_launches += 1;
  return MSCAN_OK;
}

int mscan_segments_device(mscan_ctx* c, uint32_t n_videos, const uint64_t* h_video_off, const double* h_durations,
                          const double* d_pts, const uint8_t* d_flags, mscan_segment* d_segments,
                          mscan_video_result* d_results, void* stream) try {
  if (!c || !h_video_off || !h_durations || !d_pts || !d_flags || !d_segments || !d_results) return MSCAN_ERR_INVALID;
  if (n_videos == 0) return MSCAN_OK;
  std::lock_guard<std::mutex> tail(c->tail_mu);  // K-C scratch is shared with mscan_segments*
  std::lock_guard<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : c->main_stream;
  std::vector<SegJob> jobs(n_videos);
  uint64_t ts_total = 0;
  for (uint32_t v = 0; v < n_videos; ++v) {
    if (h_video_off[v + 1] < h_video_off[v]) return fail(c, MSCAN_ERR_INVALID, "video offsets must be non-decreasing");
    const uint64_t n = h_video_off[v + 1] - h_video_off[v];
    SegJob j{};
    j.ext_begin = v;
    j.ext_end = v + 1;
    j.ts_base = ts_total;
    j.ts_cap = pow2_ge(std::max<uint64_t>(n, 1));
    j.seg_base = h_video_off[v];
    j.duration = h_durations[v];
    ts_total += j.ts_cap;
    jobs[v] = j;
  }
  const bool same = c->dev_jobs_cached.size() == n_videos && ts_total <= c->ts_cap &&
                    std::memcmp(c->dev_jobs_cached.data(), jobs.data(), sizeof(SegJob) * n_videos) == 0;
  if (!same) {
    // tables and scratch may still be read by an earlier launch: drain before touching them
    CU(cudaDeviceSynchronize());
    c->dev_jobs_cached.clear();
    if (n_videos > c->job_cap) {
      if (c->h_jobs) cudaFreeHost(c->h_jobs);
      if (c->h_res) cudaFreeHost(c->h_res);
      cudaFree(c->d_jobs);
      cudaFree(c->d_res);
      c->job_cap = 0;
      const uint32_t n = std::max(n_videos, 64u);
      CU(cudaHostAlloc((void**)&c->h_jobs, sizeof(SegJob) * n, cudaHostAllocDefault));
      CU(cudaHostAlloc((void**)&c->h_res, sizeof(mscan_video_result) * n, cudaHostAllocDefault));
      CU(cudaMalloc((void**)&c->d_jobs, sizeof(SegJob) * n));
      CU(cudaMalloc((void**)&c->d_res, sizeof(mscan_video_result) * n));
      c->job_cap = n;
    }
    if (n_videos > c->ext_cap) {
      if (c->h_exts) cudaFreeHost(c->h_exts);
      cudaFree(c->d_exts);
      c->ext_cap = 0;
      const uint32_t n = std::max(n_videos, 256u);
      CU(cudaHostAlloc((void**)&c->h_exts, sizeof(SegExtent) * n, cudaHostAllocDefault));
      CU(cudaMalloc((void**)&c->d_exts, sizeof(SegExtent) * n));
      c->ext_cap = n;
    }
    {
      uint64_t cap = c->ts_cap;
      int rc = grow(c, &c->d_ts_a, &cap, ts_total);
      if (rc) return rc;
      uint64_t cap_b = c->ts_cap;
      rc = grow(c, &c->d_ts_b, &cap_b, ts_total);
      if (rc) return rc;
      c->ts_cap = std::min(cap, cap_b);
    }
    for (uint32_t v = 0; v < n_videos; ++v) {
      c->h_exts[v] = SegExtent{h_video_off[v], h_video_off[v + 1] - h_video_off[v]};
      c->h_jobs[v] = jobs[v];
    }
    CU(cudaMemcpy(c->d_jobs, c->h_jobs, sizeof(SegJob) * n_videos, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->d_exts, c->h_exts, sizeof(SegExtent) * n_videos, cudaMemcpyHostToDevice));
    c->dev_jobs_cached = jobs;
  }
  SegArgs a{};
  a.jobs = c->d_jobs;
  a.extents = c->d_exts;
  a.pts = d_pts;
  a.flags = d_flags;
  a.ts_a = c->d_ts_a;
  a.ts_b = c->d_ts_b;
  a.segs = d_segments;
  a.results = d_results;
  a.max_gap = c->params.max_gap_sec;
  a.padding = c->params.padding_sec;
  a.min_savings_pct = c->params.min_savings_pct;
  EvPair ev{};
  if (c->profiling) {
    ev = get_events(c, 1);
    cudaEventRecord(ev.a, st);
  }
  CU(segments_launch(a, n_videos, st));
  if (c->profiling) {
    cudaEventRecord(ev.b, st);
    c->ev_pending.push_back(ev);
  }
  c->stats.segment_launches += 1;
  return MSCAN_OK;
} catch (...) {
  return on_exception(c);
}

// ---- measurement harness ------------------------------------------------------------------------
int mscan_synth_counts(mscan_ctx* c, const mvgen_spec* spec, uint64_t frame0, uint32_t n_frames, uint32_t* d_cnt,
                       void* stream) {
  if (!c || !spec || !d_cnt) return MSCAN_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->device));
  CU(synth_counts_launch(*spec, frame0, n_frames, d_cnt, stream ? (cudaStream_t)stream : c->main_stream));
  c->stats.aux_launches += 1;
  return MSCAN_OK;
}

int mscan_synth_fill(mscan_ctx* c, const mvgen_spec* spec, uint64_t frame0, uint32_t n_frames, const uint64_t* d_off,
                     mscan_mv* d_recs, double* d_pts, void* stream) {
  if (!c || !spec || !d_off || !d_recs) return MSCAN_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  CU(cudaSetDevice(c->device));
  CU(synth_fill_launch(*spec, frame0, n_frames, d_off, d_recs, d_pts, stream ? (cudaStream_t)stream : c->main_stream));
  c->stats.aux_launches += 1;
  return MSCAN_OK;
}

}  // extern "C"
