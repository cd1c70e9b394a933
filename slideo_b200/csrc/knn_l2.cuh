// knn_l2.cuh -- K10: brute-force L2 k-NN of 128-d float descriptors (SIFT variant of the path):
// cv2.BFMatcher(NORM_L2).knnMatch semantics; the whole squared distance comes out of one bf16 tcgen05 GEMM (knn_l2.cu).
#pragma once
#include "common.cuh"

namespace slideo {

struct L2Workspace {
    void* d_q_main = nullptr;   // queries as bf16 (-2 q), padded rows
    void* d_q_tail = nullptr;   // query norm tails
    void* d_scratch = nullptr;  // candidate buffers [grid][2][128][128] u64
    void* d_part = nullptr;     // per-split partial rows
    size_t q_main_cap = 0, q_tail_cap = 0, scratch_cap = 0, part_cap = 0;
    L2Workspace() = default;
    L2Workspace(const L2Workspace&) = delete;
    L2Workspace& operator=(const L2Workspace&) = delete;
    ~L2Workspace();
};

// rows are padded to a multiple of the pool tile (256) so TMA boxes never leave the allocation
int l2_rows_padded(int n);
size_t l2_main_bytes(int n);   // bf16 [rows_padded][128]
size_t l2_tail_bytes(int n);   // bf16 [rows_padded][16]
// fp32 [n][128] -> GEMM operands (main + norm tail).  is_query: rows scaled by -2 and the tail laid out for the A side.
void l2_prepare_launch(const float* d_src, int n, bool is_query, void* d_main, void* d_tail, cudaStream_t stream);
// idx: [nq][k] int32 pooled index, dist: [nq][k] float = sqrtf(d2); rows ascending by (distance, index); -1 padding
void l2_knn_launch(L2Workspace& ws, const float* d_q, int nq, const void* d_pool_main, const void* d_pool_tail, int nt, int k,
                   int32_t* d_idx, float* d_dist, int num_sms, cudaStream_t stream, int* launches);
// the reference vote (lib.rs:270-282) on float rows
void l2_vote_launch(const int32_t* d_idx, const float* d_dist, int nq, int k, const int32_t* d_q_frame, const uint16_t* d_page_of,
                    int32_t* d_votes, int n_pages, float ratio, cudaStream_t stream);

}  // namespace slideo
