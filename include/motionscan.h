/*
 * motionscan.h — C ABI of libmotionscan.so (B200 / sm_100a motion-scan hot path).
 *
 * Drop-in boundary for the motion-scan path of Motion-Estimated-Video-Trimmer.
 * The reference exposes no FFI; the seam is cut where SURVEY.md §8(b) puts it.
 * Citations are relative to the reference tree (/root/reference):
 *
 *   mscan_mv                <- AVMotionVector (FFmpeg libavutil/motion_vector.h), read at
 *                              src/motion_scanner.cpp:224-226,243-256
 *   mscan_params            <- cfg struct include/motion_trim/motion_scanner.hpp:86-93 filled at
 *                              src/motion_scanner.cpp:184-196, + Config::max_gap_sec/padding_sec/
 *                              min_savings_pct read at src/pipeline.cpp:333,337-338,358
 *   mscan_segment           <- TimeSegment include/motion_trim/types.hpp:56-59
 *   mscan_video_open        <- grid/margin derivation src/motion_scanner.cpp:189-199
 *   mscan_submit            <- the per-frame call `check_frame(frame)` src/motion_scanner.cpp:376
 *                              (decl include/motion_trim/motion_scanner.hpp:106), batched
 *   mscan_submit_device     <- same call site, records already in device memory (no host staging)
 *   mscan_submit_elided / mscan_elide_records
 *                           <- same call site; the caller's own static-elided encoding of the records
 *   mscan_submit_packed / mscan_pack_records / mscan_mv8
 *                           <- same call site; the record is the byte range [6,14) of AVMotionVector,
 *                              i.e. exactly the fields read at src/motion_scanner.cpp:243-256
 *   mscan_collect           <- `if (has_motion) ts.push_back(pts)` src/motion_scanner.cpp:382-383
 *   mscan_segments[_batch]  <- merge/segment/decision block src/pipeline.cpp:297-404
 *                              (sort+unique :302-304, no-motion :308-319, builder :325-344,
 *                              clamp+savings :349-356, decision :358-404)
 *   mscan_video_append_from <- the union of chunk results `results.extract()` src/pipeline.cpp:268 (+ src/task_queue.cpp:43-57)
 *                              when the chunks of one video were scanned by different GPUs
 *   mscan_scan_device / mscan_segments_device
 *                           <- same two islands on caller-owned device memory (decode-free
 *                              stream benchmark, BASELINE.json configs[4])
 *
 * Rules of the ABI: plain pointers and sizes, int status returns (0 = ok), no exceptions
 * cross it, no CPU fallback exists behind it (a missing GPU is MSCAN_ERR_CUDA).
 * Caller owns every host buffer; the library owns device memory, streams and pinned staging.
 */
#ifndef MOTIONSCAN_H
#define MOTIONSCAN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSCAN_ABI_VERSION 3 /* 3: mscan_stats grew (records_elided, elided_bytes); submit_device / submit_elided / elide_records / compact_records / reserve_staging / device_pci_bus_id; MSCAN_STAGING_ELIDE / _COMPACT */

/* ---- status codes ------------------------------------------------------ */
enum {
  MSCAN_OK = 0,
  MSCAN_ERR_INVALID = 1,     /* bad argument / unknown video id / bad state       */
  MSCAN_ERR_CUDA = 2,        /* CUDA runtime error or no usable device             */
  MSCAN_ERR_NOMEM = 3,       /* host or device allocation failed                   */
  MSCAN_ERR_CAPACITY = 4,    /* frame log / output buffer capacity exceeded        */
  MSCAN_ERR_UNSUPPORTED = 5  /* geometry does not fit the kernel's shared memory   */
};

/* ---- decisions (src/pipeline.cpp:308-319,358-404) ---------------------- */
enum {
  MSCAN_NO_MOTION = 0, /* empty timestamp list: reference pushes no job, writes no file */
  MSCAN_CUT = 1,       /* saved_pct >  MIN_SAVINGS_PCT: job carries the clamped segments */
  MSCAN_FULL_COPY = 2  /* saved_pct <= MIN_SAVINGS_PCT: job carries [{0, duration}]      */
};

/* ---- record type: FFmpeg's native 40-byte AVMotionVector --------------- */
typedef struct mscan_mv {
  int32_t source;        /* @0  */
  uint8_t w, h;          /* @4,@5 */
  int16_t src_x, src_y;  /* @6,@8   read by the path */
  int16_t dst_x, dst_y;  /* @10,@12 read by the path */
  /* 2 bytes padding @14 */
  uint64_t flags;        /* @16 */
  int32_t motion_x;      /* @24 */
  int32_t motion_y;      /* @28 */
  uint16_t motion_scale; /* @32 */
  /* 6 bytes padding → sizeof == 40 */
} mscan_mv;

/* ---- projected record: the 8 bytes of AVMotionVector the path reads ----- */
/* Bytes 6..13 of the native record, in place order (src/motion_scanner.cpp:243-256 reads nothing else).
 * A pure byte projection — no arithmetic of the path moves to the host. Host-fed callers that can afford
 * one pass over the side data while it is still cache-hot (right after avcodec_receive_frame) send
 * 8 B/record over PCIe instead of 40. */
typedef struct mscan_mv8 {
  int16_t src_x, src_y;
  int16_t dst_x, dst_y;
} mscan_mv8;

/* ---- how mscan_submit moves native records to the GPU -------------------- */
enum {
  MSCAN_STAGING_AUTO = 0,  /* pinned source: DMA the native records in place (no host pass);
                              pageable source: the staging pass keeps the 8 bytes the path reads — as mscan_mv8,
                              projected by the worker pool, for submits above 256 Ki records; below that (a decode
                              thread handing over its own frames) the calling thread leaves static macroblocks out:
                              MSCAN_STAGING_COMPACT while MV_THRESHOLD_SQ > 0, MSCAN_STAGING_ELIDE otherwise (default) */
  MSCAN_STAGING_PACK = 1,  /* always project on the host, even from pinned memory                */
  MSCAN_STAGING_NATIVE = 2, /* never project: pageable sources are memcpy'd as 40-byte records   */
  MSCAN_STAGING_ELIDE = 3,  /* project, and send static macroblocks (src == dst, ~90 % of a CCTV stream) as their 4 dst
                               bytes + a mask bit: a lossless transport form of the mscan_mv8 sequence that the kernel
                               expands again (≈ 4.6 B/record over PCIe at 10 % moving records). The calling thread does
                               the encoding — made for decode threads submitting their own frames. Grids beyond one
                               CTA's shared memory (8K and larger) are sent as mscan_mv8 instead. */
  MSCAN_STAGING_COMPACT = 4 /* send the mscan_mv8 projections of the MOVING records only (src != dst), nothing for a static
                               macroblock: with MV_THRESHOLD_SQ > 0 a record whose src equals its dst has mag_sq == 0 and
                               leaves check_frame at src/motion_scanner.cpp:251 before it can vote, so flags and cluster
                               counts do not depend on it (≈ 0.2–0.8 B/record over PCIe at 2–10 % moving records). The
                               calling thread compacts; the only operation on the data is that byte equality — every
                               record that could pass :251 reaches the kernel unchanged. Contexts whose threshold is
                               zero, negative or NaN (static records vote there) use MSCAN_STAGING_ELIDE instead. */
};

/* ---- TimeSegment (include/motion_trim/types.hpp:56-59) ----------------- */
typedef struct mscan_segment {
  double start;
  double end;
} mscan_segment;

/* ---- knobs (names = env vars of include/motion_trim/config.hpp:56-125) -- */
typedef struct mscan_params {
  double mv_threshold_sq;  /* MV_THRESHOLD_SQ  (code default 16.0)                     */
  int32_t block_size;      /* BLOCK_SIZE       (16) — grid rounding only               */
  int32_t block_shift;     /* BLOCK_SHIFT      (4)                                     */
  int32_t vectors_needed;  /* VECTORS_NEEDED   (2) — wrapped to uint8 like config.hpp:75 */
  int32_t clusters_needed; /* CLUSTERS_NEEDED  (2)                                     */
  float vertical_mask;     /* VERTICAL_MASK    (0.05f), float32 like config.hpp:87     */
  int32_t adjacency;       /* CLUSTER_ADJACENCY: 4 = the reference (motion_scanner.cpp:284-286); 8 adds the
                              diagonals — an extension outside the parity contract. 0 is read as 4. */
  double max_gap_sec;      /* MAX_GAP_SEC      (5.0)                                   */
  double padding_sec;      /* PADDING_SEC      (0.5)                                   */
  double min_savings_pct;  /* MIN_SAVINGS_PCT  (5.0)                                   */
} mscan_params;

/* Per-video block-grid geometry (src/motion_scanner.cpp:189-196). */
typedef struct mscan_geometry {
  int32_t grid_w;          /* int16((width  + BLOCK_SIZE-1) >> BLOCK_SHIFT)            */
  int32_t grid_h;          /* int16((height + BLOCK_SIZE-1) >> BLOCK_SHIFT)            */
  int32_t vertical_margin; /* (int)(grid_h * VERTICAL_MASK), float32 multiply          */
  int32_t reserved;
} mscan_geometry;

/* Result of the merge/segment/decision block for one video. */
typedef struct mscan_video_result {
  int32_t decision;         /* MSCAN_NO_MOTION / MSCAN_CUT / MSCAN_FULL_COPY           */
  uint32_t n_motion_frames; /* timestamps after sort+unique (pipeline.cpp:302-304)      */
  uint32_t n_segments;      /* motion segments built at pipeline.cpp:325-344            */
  uint32_t reserved;
  double out_dur;           /* Σ(end-start) after clamp, left-to-right (:350-354)       */
  double time_removed;      /* duration - out_dur (:355)                                */
  double saved_pct;         /* time_removed / duration * 100.0, or 0.0 (:356)           */
} mscan_video_result;

/* Per-context counters (bench.py reads these; times are CUDA-event milliseconds). */
typedef struct mscan_stats {
  uint64_t scan_launches;    /* K-A launches                                            */
  uint64_t segment_launches; /* K-C launches                                            */
  uint64_t aux_launches;     /* offset-scan / synth launches                            */
  uint64_t frames_scanned;
  uint64_t records_scanned;
  uint64_t h2d_bytes;
  uint64_t d2h_bytes;
  double scan_ms;            /* Σ K-A durations, only while profiling is enabled         */
  double segment_ms;         /* Σ K-C durations, only while profiling is enabled         */
  uint64_t records_projected; /* native records the staging pass projected to mscan_mv8  */
  double project_ms;          /* host wall time spent in that projection                 */
  uint64_t peer_bytes;        /* bytes mscan_video_append_from copied in from other contexts */
  uint64_t records_elided;    /* native records that went through MSCAN_STAGING_ELIDE / _COMPACT */
  uint64_t elided_bytes;      /* … and the record bytes that were sent for them                  */
} mscan_stats;

typedef struct mscan_ctx mscan_ctx;

/* ---- library-level ----------------------------------------------------- */
int mscan_abi_version(void);
int mscan_device_count(int* n_out);            /* MSCAN_ERR_CUDA when no driver/GPU */
/* "0000:3b:00.0"-style id of a device (len >= 13): lets a host look up the CPUs local to the GPU
 * (/sys/bus/pci/devices/<id>/local_cpulist) when it pins its decode threads */
int mscan_device_pci_bus_id(int device, char* buf, int len);
const char* mscan_status_string(int status);

/* config.hpp:56-125 code defaults */
int mscan_params_default(mscan_params* p);
/* defaults overridden by the reference's env vars; MSCAN_ERR_INVALID on unparsable values
 * (the reference's std::stod/stoi would throw → terminate). */
int mscan_params_from_env(mscan_params* p);
/* motion_scanner.cpp:189-196 — pure host arithmetic, usable without a GPU */
int mscan_geometry_from_dims(const mscan_params* p, int width, int height, mscan_geometry* g);

/* ---- context (one per GPU; `submit` may be called from many threads) ---- */
/* max_log_frames: capacity of the on-device frame log (0 → 16 Mi frames) = the most frames that videos which are
 * open at the same time can hold; frames of closed videos are reused.
 * slab_bytes: size of each of the 3 device record slabs + pinned staging (0 → 64 MiB: host-fed throughput is flat
 * from 32 to 256 MiB, profiles/r02_slab_sweep.log, while pinned staging costs ≈ 0.4 ms per MiB to allocate). */
int mscan_create(int device, const mscan_params* p, uint64_t max_log_frames, uint64_t slab_bytes,
                 mscan_ctx** ctx_out);
int mscan_destroy(mscan_ctx* ctx);
const char* mscan_last_error(mscan_ctx* ctx);
int mscan_get_params(mscan_ctx* ctx, mscan_params* p_out);
int mscan_sync(mscan_ctx* ctx);
int mscan_get_stats(mscan_ctx* ctx, mscan_stats* s_out);
int mscan_reset_stats(mscan_ctx* ctx);
int mscan_set_profiling(mscan_ctx* ctx, int enabled); /* CUDA-event pairs around K-A / K-C */

/* ---- host-fed path (pinned, double-buffered staging; memory_io.cpp role) */
int mscan_video_open(mscan_ctx* ctx, uint32_t video_id, int width, int height);
int mscan_video_open_geometry(mscan_ctx* ctx, uint32_t video_id, const mscan_geometry* g);
/* Append n_frames frames of one video. recs holds the frames' records back to back in FFmpeg's
 * native layout; rec_count[i] == 0 means "no MV side data" (motion_scanner.cpp:219-221).
 * Asynchronous, and callable concurrently from many decode threads (one scanner per thread in the reference,
 * include/motion_trim/motion_scanner.hpp:8-13): a call reserves its place under the context's mutex for a few
 * hundred ns and moves its records outside it, so the threads project their own cache-hot side data in parallel. If recs lies in pinned memory (mscan_host_alloc / mscan_host_register /
 * cudaHostRegister) it is DMA'd in place and must stay valid until the copy has completed, i.e. until
 * mscan_host_fence, mscan_sync, mscan_collect* or mscan_segments* returns (mscan_flush only enqueues);
 * pageable memory is consumed before the call returns: its records are projected to mscan_mv8 (the 8
 * bytes the path reads) straight into the library's pinned ring — see mscan_set_staging_mode.
 * first_frame_out (may be NULL): index, in the video's submission order, of this call's first
 * frame — what a chunk worker needs to read back its own frames with mscan_collect_range. */
int mscan_submit(mscan_ctx* ctx, uint32_t video_id, uint32_t n_frames, const double* pts,
                 const uint32_t* rec_count, const mscan_mv* recs, uint64_t* first_frame_out);
/* Same, for records the caller has already projected (mscan_pack_records, or a decoder front-end that
 * writes mscan_mv8 directly). recs must be 8-byte aligned. Pinned memory is DMA'd in place (same
 * lifetime rule as above), pageable memory is copied into the pinned ring. Native and packed submits
 * may be mixed freely, also within one video: a slab (= one K-A launch) holds one format. */
int mscan_submit_packed(mscan_ctx* ctx, uint32_t video_id, uint32_t n_frames, const double* pts,
                        const uint32_t* rec_count, const mscan_mv8* recs, uint64_t* first_frame_out);
/* Same call site, for producers whose records already lie in THIS GPU's memory (another GPU stage, a GPUDirect-storage
 * read of an MV stream file): the frames join the video's frame log like a host submit — so mscan_collect* /
 * mscan_segments* / mscan_video_append_from work on them unchanged — and K-A scans the caller's buffer in place on the
 * context's own stream; only 16 bytes of metadata per frame cross PCIe. d_recs: 16-byte aligned device pointer to the
 * frames' records back to back, native (packed == 0) or mscan_mv8 (packed != 0; readable for 16 bytes past the last
 * record). pts / rec_count: host arrays. ready_stream: CUDA stream on which the records become ready (the scan is
 * ordered after the work already enqueued on it), NULL if they are ready now. The buffer must stay valid until a
 * mscan_collect* / mscan_segments* / mscan_video_close / mscan_sync that covers these frames has returned. */
int mscan_submit_device(mscan_ctx* ctx, uint32_t video_id, uint32_t n_frames, const double* pts,
                        const uint32_t* rec_count, const void* d_recs, int packed, void* ready_stream,
                        uint64_t* first_frame_out);
/* The projection itself, usable from any thread without a context or a GPU: out[i] = bytes 6..13 of
 * recs[i]. out must be 8-byte aligned; written with streaming stores (it is read next by the DMA engine). */
int mscan_pack_records(const mscan_mv* recs, uint64_t n, mscan_mv8* out);
/* The compaction of MSCAN_STAGING_COMPACT, usable from any thread without a context or a GPU: out receives the
 * projections of the records with src != dst, in order, *n_out how many. out: 8-byte aligned, room for n records. A
 * caller that compacts into pinned memory hands the result to mscan_submit_packed with the per-frame MOVING counts as
 * rec_count — only valid for contexts with MV_THRESHOLD_SQ > 0 (see MSCAN_STAGING_COMPACT). */
int mscan_compact_records(const mscan_mv* recs, uint64_t n, mscan_mv8* out, uint64_t* n_out);
/* The static-elided transport form itself (what MSCAN_STAGING_ELIDE puts on the wire), usable without a context or a
 * GPU — ONE frame of n native records. The frame is cut into tiles of 1024 records (the last one shorter), tiles follow
 * each other 16-byte aligned; a tile is
 *     hdr : {uint32 mask, uint32 base} per block of 32 records — bit i of mask: record i of the block is moving
 *           (src != dst); base: index of the block's first entry in `src` — padded to 16 bytes
 *     dst : uint32 per record (dst_x | dst_y << 16), padded to 16 bytes
 *     src : uint32 per MOVING record (src_x | src_y << 16), in record order, padded to 16 bytes
 * i.e. exactly bytes 6..13 of every record (src/motion_scanner.cpp:243-256 reads nothing else), with the src half
 * dropped where it repeats the dst half; no arithmetic of the path. out: 16-byte aligned, cap >= mscan_elide_bound(n);
 * tile_end16[t]: end of tile t in 16-byte units from out (tile_cap >= ceil(n / 1024)); bytes_out: bytes written. */
int mscan_elide_records(const mscan_mv* recs, uint32_t n, void* out, size_t cap, uint32_t* tile_end16, uint32_t tile_cap,
                        size_t* bytes_out);
size_t mscan_elide_bound(uint32_t n);
/* Append frames the caller has already put into the static-elided form — a decode thread that runs mscan_elide_records
 * on its cache-hot side data and keeps the result in pinned memory. enc: the frames' encodings, each 16-byte aligned;
 * enc_off[n_frames + 1]: byte offset of every frame's encoding in enc (frames without records have none);
 * tile_end16: for all tiles of all frames in order, what mscan_elide_records returned (ends in 16-byte units from the
 * start of the frame's own encoding). Pinned memory is DMA'd in place (lifetime rule of mscan_submit), pageable memory is
 * copied into the pinned ring. MSCAN_ERR_UNSUPPORTED on contexts whose largest grid needs the cluster kernel. */
int mscan_submit_elided(mscan_ctx* ctx, uint32_t video_id, uint32_t n_frames, const double* pts,
                        const uint32_t* rec_count, const void* enc, const uint64_t* enc_off,
                        const uint32_t* tile_end16, uint64_t* first_frame_out);
/* MSCAN_STAGING_*; default AUTO. */
int mscan_set_staging_mode(mscan_ctx* ctx, int mode);
/* Threads (including the caller) that project one large submit; 0 → as many as the process may run on
 * (sched_getaffinity), or $MSCAN_PACK_THREADS. Submits below 256 Ki records never leave the calling thread. */
int mscan_set_pack_threads(mscan_ctx* ctx, int n_threads);
/* Allocates the pinned staging ring now (3 x slab_bytes) instead of at the first submits of pageable records: a host
 * whose decode / file threads are about to start calls it right after mscan_create, so the page pinning does not
 * stall them later. Optional. */
int mscan_reserve_staging(mscan_ctx* ctx);
int mscan_flush(mscan_ctx* ctx); /* launch whatever is staged; does not wait */
/* Per-frame results in submission order. cap = capacity of flags/full_counts (either may be NULL). */
int mscan_collect(mscan_ctx* ctx, uint32_t video_id, uint8_t* flags, uint32_t* full_counts,
                  uint32_t cap, uint32_t* n_frames_out);
/* Frames [first, first+n) of the video's submission order (one chunk's scan_range result). */
int mscan_collect_range(mscan_ctx* ctx, uint32_t video_id, uint64_t first, uint32_t n, uint8_t* flags,
                        uint32_t* full_counts);
/* Segments for the FFmpegJob of one video: on MSCAN_CUT the clamped motion segments, on
 * MSCAN_FULL_COPY the single {0,duration}, on MSCAN_NO_MOTION none. MSCAN_ERR_CAPACITY if
 * cap is too small (n_out still reports the needed count). */
int mscan_segments(mscan_ctx* ctx, uint32_t video_id, double duration, mscan_segment* out,
                   uint32_t cap, uint32_t* n_out, mscan_video_result* res_out);
/* Same, but always the clamped motion segments of pipeline.cpp:325-354 whatever the decision. */
int mscan_motion_segments(mscan_ctx* ctx, uint32_t video_id, double duration, mscan_segment* out,
                          uint32_t cap, uint32_t* n_out, mscan_video_result* res_out);
/* Many videos, one K-C launch. seg_off_out[n_videos+1] indexes out[]; job segments as above. */
int mscan_segments_batch(mscan_ctx* ctx, uint32_t n_videos, const uint32_t* video_ids,
                         const double* durations, mscan_segment* out, uint64_t cap,
                         uint64_t* seg_off_out, mscan_video_result* res_out);
int mscan_video_close(mscan_ctx* ctx, uint32_t video_id);
/* One long video split over several GPUs (each context scanned some of its chunks under its own video id):
 * appends the per-frame results (pts, flag, full count) of (src, src_video) to (dst, dst_video) by a
 * device-to-device copy — over NVLink when the two GPUs are peers — after which mscan_segments(dst, …) sees
 * the whole video; the order of chunks does not matter (K-C sorts and de-duplicates like pipeline.cpp:302-304).
 * src and dst may be the same context. The source video stays open and unchanged. */
int mscan_video_append_from(mscan_ctx* dst, uint32_t dst_video, mscan_ctx* src, uint32_t src_video);

/* pinned host memory for zero-copy submit */
int mscan_host_alloc(mscan_ctx* ctx, size_t bytes, void** p_out);
int mscan_host_free(mscan_ctx* ctx, void* p);
/* Pin memory the caller already owns (e.g. the mmap of an MV stream file; read_only for PROT_READ
 * mappings) so mscan_submit DMAs it in place. MSCAN_ERR_CUDA if the platform refuses; the caller then
 * simply submits the pageable pointer. */
int mscan_host_register(mscan_ctx* ctx, void* p, size_t bytes, int read_only);
int mscan_host_unregister(mscan_ctx* ctx, void* p);
/* Launches what is staged and blocks until every record submitted so far has been copied off the host
 * (kernels may still be running): after it returns, pinned source buffers can be reused. */
int mscan_host_fence(mscan_ctx* ctx);

/* ---- device-resident path (caller-owned device buffers) ---------------- */
int mscan_dev_alloc(mscan_ctx* ctx, size_t bytes, void** d_out);
int mscan_dev_free(mscan_ctx* ctx, void* d);
int mscan_memcpy_h2d(mscan_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);
int mscan_memcpy_d2h(mscan_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);
/* exclusive prefix sum: d_rec_off[0..n_frames] from d_rec_count[0..n_frames) */
int mscan_offsets_from_counts(mscan_ctx* ctx, const uint32_t* d_rec_count, uint32_t n_frames,
                              uint64_t* d_rec_off, void* stream);
/* K-A on device memory. d_recs: 16-byte aligned; d_rec_off[n_frames+1] record indices into d_recs;
 * d_frame_geom: per-frame index into geoms[] or NULL (all frames use geoms[0]).
 * Asynchronous on `stream` (NULL → the context's scan stream), with two exceptions that synchronise the device and are
 * therefore not legal under stream capture: a call whose geoms[] differ from the previous call's (the table is
 * re-uploaded), and grids so large that the vote counters live in global memory (beyond 16 CTAs of shared memory). */
int mscan_scan_device(mscan_ctx* ctx, const mscan_mv* d_recs, const uint64_t* d_rec_off,
                      const uint32_t* d_frame_geom, const mscan_geometry* geoms, uint32_t n_geoms,
                      uint32_t n_frames, uint8_t* d_flags, uint32_t* d_full_counts, void* stream);
/* Same on projected records (16-byte aligned base; frames may start at any record; the buffer must be
 * readable for 16 bytes past its last record). */
int mscan_scan_device_packed(mscan_ctx* ctx, const mscan_mv8* d_recs, const uint64_t* d_rec_off,
                             const uint32_t* d_frame_geom, const mscan_geometry* geoms, uint32_t n_geoms,
                             uint32_t n_frames, uint8_t* d_flags, uint32_t* d_full_counts, void* stream);
/* mscan_pack_records for records already in device memory (8-byte aligned buffers). */
int mscan_pack_records_device(mscan_ctx* ctx, const mscan_mv* d_recs, uint64_t n, mscan_mv8* d_out, void* stream);
/* K-C on device memory: video v owns frames [h_video_off[v], h_video_off[v+1]) of d_pts/d_flags.
 * d_segments: capacity = h_video_off[n_videos] entries; video v's segments start at
 * d_segments[h_video_off[v]]. d_results[n_videos]. Asynchronous on `stream`. */
int mscan_segments_device(mscan_ctx* ctx, uint32_t n_videos, const uint64_t* h_video_off,
                          const double* h_durations, const double* d_pts, const uint8_t* d_flags,
                          mscan_segment* d_segments, mscan_video_result* d_results, void* stream);

/* ---- measurement harness: deterministic synthetic MV streams ----------- */
/* Spec and generator semantics: include/mvgen_core.h (same code on host and device). */
struct mvgen_spec;
/* presets for BASELINE.json configs[0..4] */
int mscan_synth_preset(struct mvgen_spec* spec, int config, uint64_t seed);
/* host generator (no GPU needed): counts, then records written at rec_off[] (record indices) */
int mscan_synth_host_counts(const struct mvgen_spec* spec, uint64_t frame0, uint32_t n_frames,
                            uint32_t* rec_count, int n_threads);
int mscan_synth_host_fill(const struct mvgen_spec* spec, uint64_t frame0, uint32_t n_frames,
                          const uint64_t* rec_off, mscan_mv* recs, double* pts, int n_threads);
/* device generator, same bytes */
int mscan_synth_counts(mscan_ctx* ctx, const struct mvgen_spec* spec, uint64_t frame0,
                       uint32_t n_frames, uint32_t* d_rec_count, void* stream);
int mscan_synth_fill(mscan_ctx* ctx, const struct mvgen_spec* spec, uint64_t frame0,
                     uint32_t n_frames, const uint64_t* d_rec_off, mscan_mv* d_recs, double* d_pts,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MOTIONSCAN_H */
