// common.cuh -- shared declarations of libslideo_b200 (sm_100a only; no other backend, no CPU fallback).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

namespace slideo {

struct CudaError : std::runtime_error {
    cudaError_t code;
    CudaError(cudaError_t c, const std::string& what) : std::runtime_error(what), code(c) {}
};
struct ArgError : std::runtime_error { using std::runtime_error::runtime_error; };
struct StateError : std::runtime_error { using std::runtime_error::runtime_error; };
struct CapacityError : std::runtime_error { using std::runtime_error::runtime_error; };
struct NotImplError : std::runtime_error { using std::runtime_error::runtime_error; };

#define SLIDEO_CUDA(expr)                                                                                   \
    do {                                                                                                    \
        cudaError_t _e = (expr);                                                                            \
        if (_e != cudaSuccess)                                                                              \
            throw ::slideo::CudaError(_e, std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" +  \
                                              __FILE__ + ":" + std::to_string(__LINE__) + ")");             \
    } while (0)

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- K8/K9: brute-force Hamming k-NN + vote (knn_hamming.cu) -------------------------------------------------
// A k-NN entry is one uint32 key = distance << 23 | pooled index, so that integer order == the oracle's
// (distance, pooled index) order.  0xFFFFFFFF = empty slot.
constexpr int KEY_IDX_BITS = 23;
constexpr uint32_t KEY_IDX_MASK = (1u << KEY_IDX_BITS) - 1;
constexpr uint32_t KEY_EMPTY = 0xFFFFFFFFu;
constexpr int KNN_MAX_K = 32;
constexpr int KNN_MAX_POOL = 1 << KEY_IDX_BITS;  // 8,388,608 pooled descriptors (README.md:41 expects < 1000 slides)

struct KnnPlan {
    int nq = 0, nt = 0, k = 0;
    int qr = 8, tile = 1024;   // queries per thread (4 or 8) and per tile (128 threads x qr)
    int n_tiles = 0;           // query tiles
    int n_chunks = 0;          // pool chunks per tile
    int max_seg = 1;           // max CTAs sharing one tile (stream-K); > 1 -> partial rows + merge kernel
    long long total_units = 0; // n_tiles * n_chunks, divided evenly over the grid
    int grid = 0;              // persistent CTAs
    size_t scratch_bytes = 0;  // candidate buffers
    size_t partial_bytes = 0;  // per-segment partial rows (0 when every tile has one segment)
    size_t pool_bytes = 0;     // expanded pool (48 B rows)
};

struct VoteArgs {            // fused K9: vote straight out of K8's final sort (n_splits == 1) or from vote kernel
    const int32_t* q_frame;  // [nq] frame of each query row
    const uint16_t* page_of; // [nt] page of each pooled descriptor
    int32_t* votes;          // [n_frames][n_pages]
    int n_pages;
    float ratio;
};

// ctas_per_sm: resident K8 CTAs per SM (1..4, 0 = 4).  3 leaves a quarter of the register file to kernels of another
// stream (the frame path overlaps ORB extraction with K8 that way) at a ~4 % cost in K8 throughput.
KnnPlan knn_hamming_plan(int nq, int nt, int k, int num_sms, int ctas_per_sm = 0);
constexpr int KNN_TILE_QUERIES = 512;   // queries per K8 tile at 4 queries per thread
// 32 B descriptors -> the 48 B rows K8 streams ({w0..w7, w0^w1^w2, w3^w4^w5, w0^..^w6, 0}); once per pool
size_t knn_pool_expanded_bytes(int nt);
void knn_pool_expand_launch(const void* d_pool32, int nt, void* d_pool48, cudaStream_t stream);
// keys_out: [nq][k] uint32 (may be nullptr when vote != nullptr).  scratch/partial sized per plan.
void knn_hamming_launch(const KnnPlan& plan, const void* d_q, const void* d_pool48, uint32_t* d_keys_out,
                        uint32_t* d_scratch, uint32_t* d_partial, const VoteArgs* vote, cudaStream_t stream,
                        int* launches);
// frame results from the vote table
void vote_argmax_launch(const int32_t* d_votes, int n_frames, int n_pages, const int32_t* d_frame_nkp,
                        int32_t* d_results /* n_frames x 3 */, cudaStream_t stream);
void keys_to_idx_dist_launch(const uint32_t* d_keys, size_t n, int32_t* d_idx, int32_t* d_dist, cudaStream_t stream);
double microbench_run(int which, int num_sms, cudaStream_t stream);

// ---- K8 v5: bit-sliced Hamming k-NN (knn_hamming5.cu) ---------------------------------------------------------------
// The pool lives as "slabs" of 4096 rows: 256 bit rows + 9 planes of (256 - popcount) + a valid mask, 512 B each.
constexpr int KNN5_TILE = 128;            // queries per work item (one CTA of 16 warps, 8 queries per warp)
constexpr int KNN5_SLAB_ROWS = 4096;
constexpr int KNN5_SLAB_BYTES = 266 * 512;
inline int knn5_slabs(int nt) { return (nt + KNN5_SLAB_ROWS - 1) / KNN5_SLAB_ROWS; }

// Device-side state of a query stream (the frame path): detection appends, plan kernels hand ranges to K8, finalize kernels
// turn finished frames into results.  All counters are per epoch (see api.cu).
struct KnnStream {
    int q_write;               // queries appended by detection so far
    int q_matched;             // queries handed to K8 so far
    int f_write;               // frames appended so far
    int f_done;                // frames whose results have been written
    int flags;                 // sticky capacity bits: 1 FAST candidates, 2 selected keypoints, 4 query stream, 8 frame table
    int pad;
    unsigned long long pairs;  // descriptor pairs handed to K8 (accounting)
};
// One K8 launch over a range of the stream, written by knn5_plan_kernel, read by K8 / merge / finalize.
struct KnnDyn {
    int q0, nq, n_tiles, splits;
    int f_limit;               // frames appended when the range was cut
    int pad[3];
};

struct Knn5Plan {
    int nq = 0, nt = 0, k = 0, n_tiles = 0, n_slabs = 0, splits = 1, grid = 0;
    size_t partial_bytes = 0;  // per-split partial rows (0 when no tile is split)
};
size_t knn5_pool_bytes(int nt);
void knn5_pool_prepare_launch(const void* d_pool32, int nt, void* d_slabs, cudaStream_t stream);
Knn5Plan knn5_plan(int nq, int nt, int k, int num_sms);
void knn5_launch(const Knn5Plan& plan, const void* d_q, const void* d_slabs, uint32_t* d_keys_out, uint32_t* d_partial,
                 const VoteArgs* vote, cudaStream_t stream, int* launches);
// Device-driven variant: the range comes from *d_dyn (filled by knn5_plan_launch earlier in stream order); the bases are the
// bases of the whole stream.  flush == 0: whole waves of tiles only, the remainder rides with the next launch.
size_t knn5_dyn_partial_bytes(int num_sms, int k);
void knn5_plan_launch(KnnStream* d_state, KnnDyn* d_dyn, int nt, int num_sms, int flush, cudaStream_t stream);
void knn5_launch_dyn(const KnnDyn* d_dyn, int nq_max, int nt, int k, int num_sms, const void* d_q_base, const void* d_slabs,
                     uint32_t* d_keys_base, uint32_t* d_partial, const VoteArgs* vote_base, cudaStream_t stream, int* launches);
double knn5_microbench_run(int num_sms, cudaStream_t stream);   // pairs/s of the v5 inner loop alone
// frames of the stream whose queries are all matched -> (best_slide, votes, n_keypoints) rows in a host-mapped ring + progress word
void stream_finalize_launch(KnnStream* d_state, const KnnDyn* d_dyn, const int32_t* d_frame_q0, const int32_t* d_votes, int n_pages,
                            const int32_t* d_frame_nkp, int32_t* h_ring, int ring_mask, long long seq_base, int32_t* d_results,
                            volatile long long* h_progress, volatile int* h_flags, cudaStream_t stream);

}  // namespace slideo
