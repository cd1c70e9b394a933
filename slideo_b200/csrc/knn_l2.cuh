// knn_l2.cuh -- K10: brute-force L2 k-NN of 128-d float descriptors (SIFT variant of the path):
// cv2.BFMatcher(NORM_L2).knnMatch semantics, cross term -2*D*Q^T as a bf16 tcgen05 GEMM with TMEM accumulators.
#pragma once
#include "common.cuh"

namespace slideo {

struct L2Workspace {
    void* d_qb = nullptr;      // queries as bf16, padded to the tile height
    float* d_qn = nullptr;     // squared norms of the queries
    void* d_part = nullptr;    // per-split partial top-k rows
    size_t qb_cap = 0, part_cap = 0;
    ~L2Workspace();
};

// rows are padded to a multiple of the pool tile so TMA boxes never run past the allocation
int l2_rows_padded(int n);
// float [n][128] -> bf16 [n_padded][128] (K-major) + squared norms (fp32, exact for integer-valued descriptors)
void l2_prepare_launch(const float* d_src, int n, int dim, uint16_t* d_bf16, float* d_norm, cudaStream_t stream);
// idx: [nq][k] int32 pooled index, dist: [nq][k] float = sqrtf(d2); rows ascending by (distance, index); -1 padding
void l2_knn_launch(L2Workspace& ws, const float* d_q, int nq, const uint16_t* d_pool_bf16, const float* d_pool_norm, int nt, int k,
                   int32_t* d_idx, float* d_dist, int num_sms, cudaStream_t stream, int* launches);
// the reference vote (lib.rs:270-282) on float rows
void l2_vote_launch(const int32_t* d_idx, const float* d_dist, int nq, int k, const int32_t* d_q_frame, const uint16_t* d_page_of,
                    int32_t* d_votes, int n_pages, float ratio, cudaStream_t stream);

}  // namespace slideo
