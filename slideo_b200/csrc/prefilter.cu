// prefilter.cu -- K13: the changed-frame prefilter of the reference on the GPU.
//   to_small_image  (image_utils.rs:8-19)  : cv::resize(frame, ~300x400 area, INTER_AREA), 8UC3, non-integer shrink factor
//   compute_similarity (image_utils.rs:21-27): 1 - ||a - b||_2 / sqrt(255^2 * 3 * pixels)
//   MarkSimilarIter (video_capture.rs:86-102): changed iff similarity to the previous sampled frame < 0.98
// INTER_AREA is restated from OpenCV's computeResizeAreaTab + ResizeArea_Invoker<uchar, float> exactly as pinned in
// oracle/area_oracle.c: per destination pixel, `buf += src * alpha` along x in table order (fp32, no contraction), then
// `sum = beta * buf` / `sum += beta * buf` over the source rows, then round-half-even + saturate.  Every destination pixel is
// independent, so one thread computes one (x, y) for the three channels with the identical operation order.
// Compiled with -fmad=false.
#include "prefilter.cuh"

#include <math.h>

namespace slideo {

namespace {

template <int CN>
__global__ void __launch_bounds__(128) area_small_kernel(const uint8_t* __restrict__ frames, int stride, size_t frame_stride, int dw, int dh,
                                                         const int32_t* __restrict__ xoff, const int32_t* __restrict__ xsi,
                                                         const float* __restrict__ xa, const int32_t* __restrict__ yoff,
                                                         const int32_t* __restrict__ ysi, const float* __restrict__ ya,
                                                         uint8_t* __restrict__ small) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y, img = blockIdx.z;
    if (dx >= dw) return;
    const uint8_t* src = frames + (size_t)img * frame_stride;
    const int x0 = xoff[dx], x1 = xoff[dx + 1], y0 = yoff[dy], y1 = yoff[dy + 1];
    float s[CN];
#pragma unroll
    for (int c = 0; c < CN; ++c) s[c] = 0.f;
    for (int j = y0; j < y1; ++j) {
        const uint8_t* row = src + (size_t)ysi[j] * stride;
        const float beta = ya[j];
        float b[CN];
#pragma unroll
        for (int c = 0; c < CN; ++c) b[c] = 0.f;
        for (int k = x0; k < x1; ++k) {
            const uint8_t* p = row + CN * xsi[k];
            const float a = xa[k];
#pragma unroll
            for (int c = 0; c < CN; ++c) b[c] = __fadd_rn(b[c], __fmul_rn((float)p[c], a));
        }
#pragma unroll
        for (int c = 0; c < CN; ++c) s[c] = j == y0 ? __fmul_rn(beta, b[c]) : __fadd_rn(s[c], __fmul_rn(beta, b[c]));
    }
    uint8_t* out = small + ((size_t)img * dh + dy) * dw * CN + (size_t)dx * CN;
#pragma unroll
    for (int c = 0; c < CN; ++c) out[c] = (uint8_t)min(max(__float2int_rn(s[c]), 0), 255);
}

__global__ void __launch_bounds__(256) small_sumsq_kernel(const uint8_t* __restrict__ small, size_t small_bytes, unsigned long long* __restrict__ sumsq) {
    const int pair = blockIdx.y;
    const uint8_t* a = small + (size_t)pair * small_bytes;
    const uint8_t* b = a + small_bytes;
    unsigned long long acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < small_bytes; i += (size_t)gridDim.x * blockDim.x) {
        const int d = (int)a[i] - (int)b[i];
        acc += (unsigned long long)(d * d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(&sumsq[pair], acc);
}

struct Ent { int si; float a; };

// computeResizeAreaTab (imgproc/resize.cpp)
void build_axis(int ssize, int dsize, std::vector<int32_t>& off, std::vector<int32_t>& si, std::vector<float>& al) {
    const double scale = (double)ssize / dsize;
    off.assign(1, 0);
    si.clear();
    al.clear();
    for (int dx = 0; dx < dsize; ++dx) {
        const double fsx1 = dx * scale, fsx2 = fsx1 + scale;
        const double cell = scale < ssize - fsx1 ? scale : ssize - fsx1;
        int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
        sx2 = sx2 < ssize - 1 ? sx2 : ssize - 1;
        sx1 = sx1 < sx2 ? sx1 : sx2;
        if (sx1 - fsx1 > 1e-3) { si.push_back(sx1 - 1); al.push_back((float)((sx1 - fsx1) / cell)); }
        for (int sx = sx1; sx < sx2; ++sx) { si.push_back(sx); al.push_back((float)(1.0 / cell)); }
        if (fsx2 - sx2 > 1e-3) {
            double a = fsx2 - sx2;
            a = a < 1.0 ? a : 1.0;
            a = a < cell ? a : cell;
            si.push_back(sx2);
            al.push_back((float)(a / cell));
        }
        off.push_back((int32_t)si.size());
    }
}

template <typename T>
T* upload(const std::vector<T>& v) {
    T* d = nullptr;
    SLIDEO_CUDA(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(T)));
    if (!v.empty()) SLIDEO_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}

}  // namespace

void small_size(int w, int h, int* sw, int* sh) {
    const float factor = sqrtf((float)(300 * 400) / (float)(w * h));
    *sw = (int)((float)w * factor);
    *sh = (int)((float)h * factor);
}

void AreaTables::release() {
    cudaFree(d_xoff); cudaFree(d_yoff); cudaFree(d_xsi); cudaFree(d_ysi); cudaFree(d_xa); cudaFree(d_ya);
    d_xoff = d_yoff = d_xsi = d_ysi = nullptr;
    d_xa = d_ya = nullptr;
}

void AreaTables::build(int w, int h) {
    release();
    sw = w; sh = h;
    small_size(w, h, &dw, &dh);
    if (dw < 1 || dh < 1 || dw > w || dh > h) throw ArgError("frame size unsuitable for the small-image prefilter (must shrink)");
    std::vector<int32_t> off, si;
    std::vector<float> al;
    build_axis(w, dw, off, si, al);
    d_xoff = upload(off); d_xsi = upload(si); d_xa = upload(al);
    build_axis(h, dh, off, si, al);
    d_yoff = upload(off); d_ysi = upload(si); d_ya = upload(al);
}

void area_small_launch(const AreaTables& t, const uint8_t* d_frames, int n, int stride, size_t frame_stride, uint8_t* d_small, cudaStream_t stream,
                       int channels) {
    if (n <= 0) return;
    const dim3 grid(cdiv(t.dw, 128), t.dh, n);
    if (channels == 3)
        area_small_kernel<3><<<grid, 128, 0, stream>>>(d_frames, stride, frame_stride, t.dw, t.dh, t.d_xoff, t.d_xsi, t.d_xa, t.d_yoff, t.d_ysi,
                                                      t.d_ya, d_small);
    else
        area_small_kernel<1><<<grid, 128, 0, stream>>>(d_frames, stride, frame_stride, t.dw, t.dh, t.d_xoff, t.d_xsi, t.d_xa, t.d_yoff, t.d_ysi,
                                                      t.d_ya, d_small);
    SLIDEO_CUDA(cudaGetLastError());
}

void small_sumsq_launch(const uint8_t* d_small, int n, size_t small_bytes, unsigned long long* d_sumsq, cudaStream_t stream) {
    if (n <= 0) return;
    SLIDEO_CUDA(cudaMemsetAsync(d_sumsq, 0, (size_t)n * sizeof(unsigned long long), stream));
    small_sumsq_kernel<<<dim3(64, n), 256, 0, stream>>>(d_small, small_bytes, d_sumsq);
    SLIDEO_CUDA(cudaGetLastError());
}

}  // namespace slideo
