// verify.cuh -- K12: geometric verification (top-40 slides by votes -> RANSAC similarity rating -> gates), lib.rs:284-333.
#pragma once
#include "common.cuh"

namespace slideo {

constexpr int VERIFY_TOP_SLIDES = 40;    // lib.rs:295  take(40)
constexpr int VERIFY_TOP_RATED = 10;     // lib.rs:330  truncate(10)
constexpr int VERIFY_MAX_ITERS = 2000;   // image_utils.rs:52

// one per frame; layout == slideo_b200_verify_result (include/slideo_b200.h)
struct VerifyRecord {
    int32_t n_candidates, n_survivors;
    int32_t cand_page[VERIFY_TOP_SLIDES], cand_votes[VERIFY_TOP_SLIDES], cand_rating[VERIFY_TOP_SLIDES];
    int32_t survivor_page[VERIFY_TOP_RATED], survivor_rating[VERIFY_TOP_RATED];
};

struct VerifyArgs {
    int n_frames, n_pages, k;
    float ratio;
    const int32_t* d_votes;       // [n_frames][n_pages]
    const uint32_t* d_keys;       // k-NN rows of the frames' queries: [total_q][k]
    const int32_t* d_frame_q0;    // [n_frames + 1] first query of each frame
    const uint16_t* d_page_of;    // [Nt]
    const float2* d_frame_pt;     // [total_q] KeyPoint.pt of the frame keypoints
    const float2* d_pool_pt;      // [Nt] KeyPoint.pt of the pooled (slide) keypoints
    int32_t *d_cand_page, *d_cand_votes, *d_n_cand, *d_rating;   // workspaces: [n_frames][40] x3, [n_frames]
    void* d_corr;                 // verify_corr_bytes(total_q * k)
    VerifyRecord* d_out;          // [n_frames]
    int32_t* d_best_it;           // [n_frames][40] RANSAC iteration whose model won (photometric stage), may be null
    int32_t* d_survivor_cand;     // [n_frames][10] candidate slot of each survivor (photometric stage), may be null
    const uint2* d_pairs = nullptr;   // verify_pairs_bytes(): RANSAC sample pairs of every correspondence count n <= VERIFY_PAIRS_N (built once)
};
constexpr int VERIFY_PAIRS_N = 256;
size_t verify_pairs_bytes();
void verify_pairs_build(void* d_pairs, cudaStream_t stream);

// K14: photometric verification of the survivors (lib.rs:335-389): LM-refined matrix -> warpAffine(WARP_INVERSE_MAP, nearest)
// of the frame into the slide's geometry -> INTER_AREA small image -> sum of squared differences to the slide's small image
// geometry of one page size: the slide image size the frame is warped to (lib.rs:339-348, slide_info.img.size()) and the area tables
// of to_small_image for it (image_utils.rs:8-19)
struct PageGeom {
    int page_w, page_h, small_w, small_h;
    const int32_t *d_xoff, *d_xsi, *d_yoff, *d_ysi;
    const float *d_xa, *d_ya;
};

struct PhotoArgs {
    int n_frames, k;
    const VerifyRecord* d_records;     // [n_frames]
    const int32_t* d_survivor_cand;    // [n_frames][10]
    const int32_t* d_best_it;          // [n_frames][40]
    const int32_t* d_cand_votes;       // [n_frames][40]
    const void* d_corr;
    const int32_t* d_frame_q0;
    const float2* d_frame_pt;
    const float2* d_pool_pt;
    const uint8_t* d_frames;           // BGR8 frames of the group
    int frame_w, frame_h, frame_stride_row;
    size_t frame_stride;
    const PageGeom* d_geom;            // one entry per distinct page size of the deck
    const int32_t* d_page_class;       // [n_pages] index into d_geom
    const unsigned long long* d_page_small_off;   // [n_pages] byte offset of the page's small image in d_page_small
    int max_small_w, max_small_h;      // over the deck's page sizes (launch geometry)
    const uint8_t* d_page_small;       // gray small images, page p at d_page_small_off[p], small_h * small_w bytes of its size class
    double* d_refined;                 // [n_frames][10][4]  (a, b, tx, ty)
    unsigned long long* d_sumsq;       // [n_frames][10]
};
void photometric_launch(const PhotoArgs& a, cudaStream_t stream, int* launches);

size_t verify_corr_bytes(long long total_entries);
void verify_launch(const VerifyArgs& a, cudaStream_t stream, int* launches);

}  // namespace slideo
