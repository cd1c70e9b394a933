// prefilter.cuh -- K13: changed-frame prefilter: to_small_image (INTER_AREA) + compute_similarity,
// crates/matching-opencv/src/image_utils.rs:8-27, video_capture.rs:86-102.
#pragma once
#include <vector>

#include "common.cuh"

namespace slideo {

// per-axis area tables in CSR form (OpenCV computeResizeAreaTab): destination d owns entries [off[d], off[d+1])
struct AreaTables {
    int sw = 0, sh = 0, dw = 0, dh = 0;
    int32_t *d_xoff = nullptr, *d_yoff = nullptr, *d_xsi = nullptr, *d_ysi = nullptr;
    float *d_xa = nullptr, *d_ya = nullptr;
    void build(int sw, int sh);   // host tables for (sw, sh) -> small size, uploaded
    void release();
    ~AreaTables() { release(); }
};

void small_size(int w, int h, int* sw, int* sh);   // image_utils.rs:10-16 (f32, truncating)
// 8-bit images (device, 3 interleaved channels or 1) -> small images, image i to d_small + i * (dw*dh*channels)
void area_small_launch(const AreaTables& t, const uint8_t* d_frames, int n, int stride, size_t frame_stride, uint8_t* d_small, cudaStream_t stream,
                       int channels = 3);
// sumsq[i] = sum over all bytes of (small[i] - small[i+1])^2 for i in [0, n)   (n+1 consecutive small images)
void small_sumsq_launch(const uint8_t* d_small, int n, size_t small_bytes, unsigned long long* d_sumsq, cudaStream_t stream);

}  // namespace slideo
