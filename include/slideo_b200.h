/*
 * slideo_b200.h -- C ABI of libslideo_b200.so: the B200-native (sm_100a) replacement for the per-frame hot
 * path of hediet/slideo's `crates/matching-opencv`.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types, status codes instead of panics,
 * no exception ever crosses it.  A Rust `matching-b200` crate binds these symbols 1:1 and
 * implements the reference's traits `ImageVideoMatcher / VideoMatcher / VideoMatcherTask`
 * (crates/matching/src/lib.rs:7-29) on top of them.  Every entry point cites the reference interface it replaces.
 *
 * Ownership  : the caller owns every buffer it passes; the library copies what it keeps.  The only interior
 *              pointer ever returned is the `const char*` of slideo_b200_last_error (valid until the next call
 *              on that ctx) and the device pointers of slideo_b200_pool_device_view (valid until the pool changes).
 * Threading  : a ctx is NOT re-entrant: one in-flight call per ctx.  Different ctxs (one per GPU) may be driven
 *              from different threads / processes.  (The reference keeps one ORB + one FLANN object per rayon
 *              thread for the same reason: crates/matching-opencv/src/lib.rs:88-90,136.)
 * Errors     : every function returns slideo_b200_status (0 = OK, negative = error).  The reference unwraps /
 *              panics everywhere (lib.rs:95-101, feature_extractor.rs:24,40, flann.rs:16-46); a Rust binding maps
 *              non-zero to panic!() to keep the trait signatures unchanged.
 * No fallback: there is no CPU path.  Without a CUDA device slideo_b200_create fails with SLIDEO_B200_E_CUDA.
 */
#ifndef SLIDEO_B200_H
#define SLIDEO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLIDEO_B200_ABI_VERSION 1

typedef enum slideo_b200_status {
    SLIDEO_B200_OK = 0,
    SLIDEO_B200_E_INVALID_ARG = -1,
    SLIDEO_B200_E_CUDA = -2,
    SLIDEO_B200_E_OOM = -3,
    SLIDEO_B200_E_NOTIMPL = -4,
    SLIDEO_B200_E_STATE = -5,    /* call order violated (e.g. match before finalize_pool) */
    SLIDEO_B200_E_CAPACITY = -6, /* an internal fixed capacity was exceeded; nothing was silently truncated */
    SLIDEO_B200_E_INTERNAL = -7
} slideo_b200_status;

typedef enum slideo_b200_descriptor_kind {
    SLIDEO_B200_DESC_ORB256 = 0, /* 32-byte binary descriptors, Hamming distance (the reference's path) */
    SLIDEO_B200_DESC_SIFT128 = 1 /* 128-d float descriptors, L2 distance (north_star variant; not in the reference) */
} slideo_b200_descriptor_kind;

/* All algorithm parameters of the reference are hard-coded literals; they are collected here with those values
 * as defaults (slideo_b200_default_config).
 *   ORB       : crates/matching-opencv/src/feature_extractor.rs:13-23
 *   knn_k     : crates/matching-opencv/src/lib.rs:266
 *   vote_ratio: crates/matching-opencv/src/lib.rs:275 */
typedef struct slideo_b200_config {
    int32_t abi_version;     /* SLIDEO_B200_ABI_VERSION */
    int32_t device;          /* CUDA device ordinal */
    int32_t nfeatures;       /* 2000 */
    float scale_factor;      /* 1.2f */
    int32_t nlevels;         /* 8 */
    int32_t edge_threshold;  /* 62 */
    int32_t patch_size;      /* 62 */
    int32_t fast_threshold;  /* 20 */
    int32_t knn_k;           /* 30 (1..32) */
    float vote_ratio;        /* 1.05f */
    int32_t descriptor_kind; /* slideo_b200_descriptor_kind */
    int32_t max_batch;       /* frames processed per internal batch (default 32) */
    int32_t keep_matches;    /* !=0: keep the k-NN rows of the last match call for slideo_b200_get_matches */
    int32_t geometric_verification; /* 1: match_frames_* also runs the RANSAC gate (lib.rs:284-333), see slideo_b200_get_verification;
                                       2: plus the warp + similarity gate (lib.rs:335-389), see slideo_b200_get_decisions */
    int32_t knn_impl;        /* 0 (= 5): bit-sliced K8 v5 everywhere; 4: the XOR/POPC K8 v4 in the stage-level k-NN entry points (kept as an
                                independent implementation for cross-checks; the frame path always runs v5) */
    int32_t reserved[1];
} slideo_b200_config;

/* Hot-path output per frame: the head of the reference's ranking, lib.rs:268-295.
 * best_slide = argmax over pages of the 1.05-ratio votes (ties -> lowest page index; -1 when nobody voted),
 * votes = votes[best_slide], n_keypoints = ORB keypoints found in the frame (-1: capacity error on this frame). */
typedef struct slideo_b200_frame_result {
    int32_t best_slide;
    int32_t votes;
    int32_t n_keypoints;
} slideo_b200_frame_result;

/* One k-NN entry, field for field the reference's KeyedDMatch (crates/matching-opencv/src/flann.rs:51-59). */
typedef struct slideo_b200_match {
    int32_t query_idx; /* descriptor index within the frame (canonical order: octave, y, x) */
    int32_t train_idx; /* descriptor index within the page */
    int32_t source;    /* page index (position in add_page order == `images` order, lib.rs:259) */
    float distance;    /* Hamming distance as float, like cv::DMatch */
} slideo_b200_match;

/* Device-time accounting of the last call(s) on a ctx, CUDA events on the ctx's own stream. */
typedef struct slideo_b200_timings {
    float ms_detect;         /* K1-K7 ORB extraction kernels */
    float ms_knn;            /* K8 brute-force k-NN kernel(s) */
    float ms_vote;           /* K9 vote + argmax (0 when fused into K8) */
    float ms_h2d;            /* host->device copies issued by the library */
    int64_t knn_pairs;       /* descriptor pairs evaluated by K8 */
    int64_t knn_launches;    /* K8 launches */
    int64_t kernel_launches; /* all kernel launches of the library */
    int64_t frames;          /* frames processed */
    float ms_total;          /* whole match_* calls, first enqueue to last result on the ctx stream */
    float ms_verify;         /* K12 geometric verification */
} slideo_b200_timings;

/* Geometric verification of one frame (lib.rs:284-333): the (<= 40) slides with most votes in
 * ranking order (ties -> lower page index) with their RANSAC rating = inlier count of
 * estimateAffinePartial2D(slide keypoints -> frame keypoints, RANSAC, 3.0, 2000, 0.99, 10) (image_utils.rs:45-60), and the
 * (<= 10) survivors of `sort by rating; truncate(10); retain(rating > 50 && rating / best > 0.2)` in that order. */
#define SLIDEO_B200_TOP_SLIDES 40
#define SLIDEO_B200_TOP_RATED 10
typedef struct slideo_b200_verify_result {
    int32_t n_candidates;
    int32_t n_survivors;
    int32_t cand_page[SLIDEO_B200_TOP_SLIDES];
    int32_t cand_votes[SLIDEO_B200_TOP_SLIDES];
    int32_t cand_rating[SLIDEO_B200_TOP_SLIDES];
    int32_t survivor_page[SLIDEO_B200_TOP_RATED];
    int32_t survivor_rating[SLIDEO_B200_TOP_RATED];
} slideo_b200_verify_result;

/* Final decision for one frame (cfg.geometric_verification == 2): the survivors of the RANSAC gate re-ranked by the similarity
 * of the warped frame to the slide (lib.rs:335-389: warp_affine(WARP_INVERSE_MAP) with the LM-refined matrix -> to_small_image ->
 * compute_similarity; sort by similarity; retain > 0.5).  image = rated_page[0] or -1 (Matching.image = None). */
typedef struct slideo_b200_decision {
    int32_t image;
    int32_t n_rated;
    int32_t rated_page[SLIDEO_B200_TOP_RATED];
    float rated_similarity[SLIDEO_B200_TOP_RATED];
    double refined_matrix[SLIDEO_B200_TOP_RATED][4]; /* (a, b, tx, ty) of [a -b tx; b a ty] for the RANSAC-gate survivors, in survivor order */
} slideo_b200_decision;

typedef struct slideo_b200_ctx slideo_b200_ctx;

/* ---- lifecycle ------------------------------------------------------------------------------------------ */
/* Fills *cfg with the reference's literals (see slideo_b200_config). */
int32_t slideo_b200_default_config(slideo_b200_config* cfg);
/* Replaces OpenCVImageVideoMatcher::default() + the thread-local FeatureExtractor (lib.rs:33-35, 88-90). */
int32_t slideo_b200_create(const slideo_b200_config* cfg, slideo_b200_ctx** out_ctx);
int32_t slideo_b200_destroy(slideo_b200_ctx* ctx);
/* Human-readable text of the last error on this ctx ("" if none).  ctx may be NULL (global create errors). */
const char* slideo_b200_last_error(const slideo_b200_ctx* ctx);
const char* slideo_b200_version(void);

/* ---- page pool  (replaces ProcessedImage::compute lib.rs:93-131 + FlannMatcher::new flann.rs:64-71) ------- */
/* Extracts (ORB256: K1-K7, SIFT128: K11) one page given as the 8-bit gray that `imread(path, IMREAD_GRAYSCALE)` returns
 * (lib.rs:98) and appends its descriptors to the pool.  Pages get consecutive indices in call order. */
int32_t slideo_b200_add_page_gray8(slideo_b200_ctx* ctx, const uint8_t* px, int32_t w, int32_t h, int32_t stride,
                                   int32_t* out_n_keypoints);
/* Appends a page whose descriptors were computed elsewhere (n x 32 bytes for ORB256, n x 128 floats for SIFT128). */
int32_t slideo_b200_add_page_descriptors(slideo_b200_ctx* ctx, const void* desc, int32_t n);
/* Like add_page_descriptors, plus the KeyPoint.pt of every descriptor (n x 2 floats, level-0 coordinates): needed by the
 * geometric verification when the page features were extracted elsewhere.  ORB256 only. */
int32_t slideo_b200_add_page_features(slideo_b200_ctx* ctx, const void* desc, const float* pt_xy, int32_t n);
/* Concatenates the pages into the device-resident pool (flann.rs add + train; brute force has no index to build). */
int32_t slideo_b200_finalize_pool(slideo_b200_ctx* ctx);
int32_t slideo_b200_pool_info(const slideo_b200_ctx* ctx, int32_t* out_n_descriptors, int32_t* out_n_pages);
/* Pool replication across GPUs (one ctx per GPU): export on the rank that built it, import on the others.
 * Host variant: desc = n_desc*desc_bytes, page_offsets = n_pages+1 int32.  Either pointer may be NULL to skip. */
int32_t slideo_b200_pool_export(const slideo_b200_ctx* ctx, void* desc, int32_t* page_offsets);
int32_t slideo_b200_pool_import(slideo_b200_ctx* ctx, const void* desc, int32_t n_desc, const int32_t* page_offsets,
                                int32_t n_pages);
/* Device variant for an NCCL broadcast done by the host process (torch.distributed / ncclBroadcast): reserve a
 * pool of the given geometry, expose its device buffers, then call slideo_b200_pool_commit once they are filled.
 * The descriptor buffer holds n_desc x 32 bytes (ORB256) or n_desc x 128 floats (SIFT128; the bf16 tensor-core
 * operands are derived from it at commit). */
int32_t slideo_b200_pool_reserve(slideo_b200_ctx* ctx, int32_t n_desc, int32_t n_pages);
int32_t slideo_b200_pool_device_view(slideo_b200_ctx* ctx, void** d_desc, size_t* desc_bytes, void** d_page_offsets,
                                     size_t* offsets_bytes);
int32_t slideo_b200_pool_commit(slideo_b200_ctx* ctx);
/* Device buffer of the pooled keypoint coordinates (n_desc x 2 floats; geometric verification).  On the rank that built the
 * pool: *has_points tells whether every page came with coordinates.  On a reserved ctx: the buffer to receive them into;
 * call with received != 0 after filling it (before pool_commit) to declare the coordinates valid. */
int32_t slideo_b200_pool_points_device_view(slideo_b200_ctx* ctx, void** d_pt, size_t* bytes, int32_t* has_points, int32_t received);

/* Small gray images of the pages (slide.small_img, lib.rs:105; the warp + similarity gate of cfg.geometric_verification == 2).  On
 * the rank that built the pool: *page_w / *page_h = the uniform page size (0 when the pages differ in size or came without images:
 * the gates themselves handle mixed page sizes, as the reference does per slide -- only this replication view carries one size)
 * and *d_small = n_pages small images.  On a reserved ctx: pass the sender's page size as set_w / set_h to size the buffer and
 * get the pointer to receive into (before pool_commit). */
int32_t slideo_b200_pool_pages_device_view(slideo_b200_ctx* ctx, void** d_small, size_t* bytes, int32_t* page_w, int32_t* page_h,
                                           int32_t set_w, int32_t set_h);

/* ---- the per-frame hot path  (replaces match_images_with_frame lib.rs:249-295, head of the ranking) ------- */
/* n BGR 8UC3 frames (what VideoCapture::retrieve yields, video_capture.rs:45-53), HOST memory, frame i at
 * frames + i*frame_stride; rows `stride` bytes apart.  Copies to the device, extracts ORB (or SIFT for a SIFT128 ctx),
 * k-NN against the pool, votes, writes n results.  Pinned host memory (slideo_b200_host_alloc) makes the copies asynchronous. */
int32_t slideo_b200_match_frames_bgr8(slideo_b200_ctx* ctx, const uint8_t* frames, int32_t n, int32_t w, int32_t h,
                                      int32_t stride, size_t frame_stride, slideo_b200_frame_result* out);
/* Same with frames already resident in device memory (HBM-resident throughput measurement). */
int32_t slideo_b200_match_frames_bgr8_device(slideo_b200_ctx* ctx, const void* d_frames, int32_t n, int32_t w,
                                             int32_t h, int32_t stride, size_t frame_stride,
                                             slideo_b200_frame_result* out);
/* The same path, asynchronous (the reference's streaming loop spawns one job per changed frame while decoding goes on,
 * lib.rs:196-220).  submit enqueues the uploads, the extraction and the k-NN launches of n frames and returns at once with a
 * ticket; nothing on the host waits for the GPU.  collect blocks until the results of that ticket are complete and writes
 * n results.  Several tickets may be in flight (at most 65536 uncollected frames); the query stream is continuous across
 * submits, so the upload + extraction of one call overlap the k-NN of the previous one.  The host frame buffer of a submit
 * must stay valid (and should be pinned) until its ticket is collected.  ORB256 ctxs without geometric_verification /
 * keep_matches (those need match_frames_*, which is submit + collect + the verification tail).
 * match_frames_bgr8 == submit_frames_bgr8 + collect, result for result. */
int32_t slideo_b200_submit_frames_bgr8(slideo_b200_ctx* ctx, const uint8_t* frames, int32_t n, int32_t w, int32_t h, int32_t stride,
                                       size_t frame_stride, int64_t* out_ticket);
int32_t slideo_b200_submit_frames_bgr8_device(slideo_b200_ctx* ctx, const void* d_frames, int32_t n, int32_t w, int32_t h,
                                              int32_t stride, size_t frame_stride, int64_t* out_ticket);
/* out may be NULL (drop the results); cap = capacity of out in frames; *out_n = frames of the ticket. */
int32_t slideo_b200_collect(slideo_b200_ctx* ctx, int64_t ticket, slideo_b200_frame_result* out, int32_t cap, int32_t* out_n);
/* Match pre-extracted descriptors of n frames (frame i owns rows [frame_offsets[i], frame_offsets[i+1])). */
int32_t slideo_b200_match_descriptors(slideo_b200_ctx* ctx, const void* desc, const int32_t* frame_offsets, int32_t n,
                                      slideo_b200_frame_result* out);
/* k-NN rows of frame `frame_i` of the last match call (needs cfg.keep_matches): n_query*k entries, row-major,
 * each row ascending by (distance, pooled index) -- what FlannMatcher::knn_match returns (flann.rs:73-89). */
int32_t slideo_b200_get_matches(slideo_b200_ctx* ctx, int32_t frame_i, slideo_b200_match* out, int32_t cap_rows,
                                int32_t* out_rows);

/* Verification records of frames [frame0, frame0 + n) of the last match_frames_* call (needs cfg.geometric_verification and
 * pages added through add_page_gray8 / add_page_features). */
int32_t slideo_b200_get_verification(slideo_b200_ctx* ctx, int32_t frame0, int32_t n, slideo_b200_verify_result* out);

/* Decisions of frames [frame0, frame0 + n) of the last match_frames_* call (needs cfg.geometric_verification == 2 and pages
 * added through add_page_gray8, all of one size). */
int32_t slideo_b200_get_decisions(slideo_b200_ctx* ctx, int32_t frame0, int32_t n, slideo_b200_decision* out);

/* ---- changed-frame prefilter  (replaces MarkSimilarIter video_capture.rs:60-103 + image_utils.rs:8-27) -------------- */
/* For n consecutive SAMPLED frames (HOST, BGR 8UC3): similarity of each frame's small image (resize INTER_AREA to the
 * ~300x400-area size) to the previous sampled frame's, and changed = similarity < 0.98.  The first frame after a reset
 * (reset != 0, or the first call on a ctx, or a geometry change) has similarity 0 and is changed; otherwise the chain
 * continues from the last frame of the previous call.  out_changed: n bytes (0/1); out_similarity: n floats (may be NULL). */
int32_t slideo_b200_mark_changed_bgr8(slideo_b200_ctx* ctx, const uint8_t* frames, int32_t n, int32_t w, int32_t h, int32_t stride,
                                      size_t frame_stride, int32_t reset, uint8_t* out_changed, float* out_similarity);
/* Same with frames resident in device memory. */
int32_t slideo_b200_mark_changed_bgr8_device(slideo_b200_ctx* ctx, const void* d_frames, int32_t n, int32_t w, int32_t h,
                                             int32_t stride, size_t frame_stride, int32_t reset, uint8_t* out_changed,
                                             float* out_similarity);

/* ---- stage-level entry points (parity tests, matcher sweeps) --------------------------------------------- */
/* ORB::detectAndCompute (feature_extractor.rs:29-46) on one HOST image, channels = 1 (gray) or 3 (BGR).
 * Canonical order (octave, y, x).  kp_i: n x 4 {x_level, y_level, octave, score}; kp_f: n x 4 {pt.x, pt.y, size,
 * angle_deg}; desc: n x 32.  Any output pointer may be NULL. */
int32_t slideo_b200_extract_orb(slideo_b200_ctx* ctx, const uint8_t* img, int32_t w, int32_t h, int32_t stride,
                                int32_t channels, int32_t* kp_i, float* kp_f, uint8_t* desc, int32_t cap,
                                int32_t* out_n);
/* SIFT::detectAndCompute with cv::SIFT::create()'s defaults (the north_star's SIFT-128 variant; would sit where
 * feature_extractor.rs:29-46 calls ORB) on one HOST image, channels = 1 (gray) or 3 (BGR).  OpenCV's output order
 * (KeyPoint_LessThan after removeDuplicated).  kp_f: n x 5 {pt.x, pt.y, size, angle_deg, response}; kp_octave: n packed
 * cv::KeyPoint::octave; desc: n x 128 floats (integer-valued 0..255).  Any output pointer may be NULL. */
int32_t slideo_b200_extract_sift(slideo_b200_ctx* ctx, const uint8_t* img, int32_t w, int32_t h, int32_t stride,
                                 int32_t channels, float* kp_f, int32_t* kp_octave, float* desc, int32_t cap,
                                 int32_t* out_n);
/* Gaussian layer (0..5) of octave `octave` of the scale space of the last SIFT extract/match call (image 0), w x h floats. */
int32_t slideo_b200_debug_fetch_sift(slideo_b200_ctx* ctx, int32_t octave, int32_t layer, float* out, size_t cap_bytes,
                                     int32_t* out_w, int32_t* out_h, int32_t* out_n_octaves);
/* Intermediate images of the last extract/match call, image 0 of the batch: what = 0 pyramid level, 1 blurred
 * level, 2 FAST candidates (packed score<<24|y<<12|x, unordered).  Tightly packed into `out`. */
int32_t slideo_b200_debug_fetch(slideo_b200_ctx* ctx, int32_t what, int32_t level, void* out, size_t cap_bytes,
                                int32_t* out_w, int32_t* out_h);
/* BFMatcher(NORM_HAMMING).knnMatch semantics: k smallest by (distance, index), HOST buffers.
 * idx/dist: nq x k int32, rows padded with -1 when nt < k. */
int32_t slideo_b200_bf_knn_hamming(slideo_b200_ctx* ctx, const uint8_t* q, int32_t nq, const uint8_t* t, int32_t nt,
                                   int32_t k, int32_t* idx, int32_t* dist);
/* Same on DEVICE buffers (q, t 16-byte aligned); keys_out: nq x k uint32 = dist<<23 | idx (0xFFFFFFFF = empty). */
int32_t slideo_b200_bf_knn_hamming_device(slideo_b200_ctx* ctx, const void* d_q, int32_t nq, const void* d_t,
                                          int32_t nt, int32_t k, void* d_keys_out);
/* BFMatcher(NORM_L2).knnMatch: dist = sqrtf(sum (a-b)^2); bf16 tcgen05 cross term, exact for integer-valued
 * descriptors (cv2 SIFT).  HOST buffers. */
int32_t slideo_b200_bf_knn_l2(slideo_b200_ctx* ctx, const float* q, int32_t nq, const float* t, int32_t nt,
                              int32_t dim, int32_t k, int32_t* idx, float* dist);

/* Same on DEVICE buffers (fp32 q/t, 16-byte aligned; idx int32 nq x k, dist float nq x k). */
int32_t slideo_b200_bf_knn_l2_device(slideo_b200_ctx* ctx, const void* d_q, int32_t nq, const void* d_t, int32_t nt,
                                     int32_t dim, int32_t k, void* d_idx, void* d_dist);

/* ---- utilities ------------------------------------------------------------------------------------------ */
int32_t slideo_b200_host_alloc(void** out, size_t bytes); /* pinned host memory */
int32_t slideo_b200_host_free(void* p);
int32_t slideo_b200_get_timings(slideo_b200_ctx* ctx, slideo_b200_timings* out, int32_t reset);
/* Optional progress callback, the C form of ProgressReporter (crates/matching/src/progress.rs:3-17: Fn(u64, u64, &str)): while a
 * match_frames_* / collect call waits for the GPU it reports (frames of this call finished, frames of this call, message) from
 * the calling thread, once per finished k-NN launch group and once when the call's frames are complete.  fn = NULL switches it off.  The message pointer is only valid during
 * the callback. */
typedef void (*slideo_b200_progress_fn)(uint64_t processed, uint64_t total, const char* message, void* user);
int32_t slideo_b200_set_progress_callback(slideo_b200_ctx* ctx, slideo_b200_progress_fn fn, void* user);
/* Integer-pipe micro-benchmarks used as the K8 roofline denominators: which = 0 LOP3, 1 POPC (thread-level ops per second over
 * the whole GPU), 2 the inner-loop instruction mix of K8 v4 (XOR/POPC), 3 the inner loop of K8 v5 (bit-sliced: list walk +
 * carry-save tree + compare on synthetic shared-memory contents) -- both in descriptor pairs per second. */
int32_t slideo_b200_microbench(slideo_b200_ctx* ctx, int32_t which, double* out_per_second);
int32_t slideo_b200_synchronize(slideo_b200_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SLIDEO_B200_H */
