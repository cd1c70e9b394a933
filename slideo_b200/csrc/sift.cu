// sift.cu -- K11: batched SIFT (cv::SIFT::create() defaults: nOctaveLayers 3, contrastThreshold 0.04, edgeThreshold 10,
// sigma 1.6, CV_32F descriptors) for the north_star's SIFT-128 / L2 variant of the path; it sits where
// crates/matching-opencv/src/feature_extractor.rs:29-46 calls ORB::detectAndCompute.  Every kernel follows the stage of
// oracle/sift_oracle.c named beside it (S.1 .. S.11), operation for operation: the file is compiled -fmad=false and every
// fused multiply-add OpenCV's AVX2 build performs is an explicit __fmaf_rn.
//
//   sift_gray_kernel      BGR -> gray (cvtColor fixed point)
//   sift_blur_kernel      S.2/S.3 separable Gaussian on float images, 64x64 tiles staged in shared memory with their halo;
//                         (interior tiles by 16-byte cp.async); sift_up2_kernel writes the 2x INTER_LINEAR initial image
//   sift_half_kernel      S.5 INTER_NEAREST half of layer 3 -> layer 0 of the next octave
//   sift_extrema_kernel   S.8 26-neighbour extrema; the DoG images are never materialised (differences of the Gaussian tiles)
//   sift_refine_kernel    S.6 adjustLocalExtrema, one thread per candidate, survivors compacted
//   sift_orient_kernel    S.7 orientation histogram (raster-order sums) + peaks, one thread per refined candidate
//   sift_rank_kernel /    S.9 KeyPoint_LessThan order by rank counting, removeDuplicatedSorted, firstOctave = -1 rescale
//   sift_unique_kernel
//   sift_bucket_*         descriptor threads are scheduled by window radius (counting sort), so warps run equal-length loops
//   sift_descriptor_kernel S.10 4x4x8 histogram, one thread per keypoint with its 360 bins in shared memory
//
// The oracle calls libm for three values (exp2f for the keypoint size, cosf / sinf for the descriptor rotation); OpenCV calls the
// same libm.  glibc evaluates them as short double-precision polynomials with one final rounding (flt-32/e_exp2f.c,
// s_sincosf.h -- the ARM optimized-routines algorithms); glibc_exp2f / glibc_sinf / glibc_cosf below restate exactly those
// algorithms, so the device returns the same bits (checked against glibc 2.39 over every 7th float of [0, 1.6] and every 5th of
// [0, 6.4], with and without FMA contraction: 0 differences; tests/test_gpu_sift.py checks the device against the host libm).
#include "sift.cuh"

#include <math.h>
#include <string.h>

namespace slideo {

namespace {

// ---- glibc's float exp2 / sin / cos, restated (see the header comment) -----------------------------------------------------------
// 2^(i/32) as doubles with the exponent contribution of i removed (glibc's __exp2f_data.tab)
__device__ const unsigned long long GLIBC_EXP2F_TAB[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, 0x3fef72b83c7d517bull, 0x3fef54873168b9aaull,
    0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, 0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull, 0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull,
    0x3feea11473eb0187ull, 0x3feea589994cce13ull, 0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull, 0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full,
    0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};

// exp2f for |x| < 128 (the keypoint size uses x in (0, 1.2))
__device__ __forceinline__ float glibc_exp2f(float x) {
    const double SHIFT = 211106232532992.0;   // 0x1.8p+52 / 32
    const double C0 = 0x1.c6af84b912394p-5, C1 = 0x1.ebfce50fac4f3p-3, C2 = 0x1.62e42ff0c52d6p-1;
    const double xd = (double)x;
    double kd = __dadd_rn(xd, SHIFT);
    const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
    kd = __dsub_rn(kd, SHIFT);
    const double r = __dsub_rn(xd, kd);
    const unsigned long long t = GLIBC_EXP2F_TAB[ki & 31] + (ki << 47);
    const double sc = __longlong_as_double((long long)t);
    const double z = __dadd_rn(__dmul_rn(C0, r), C1);
    const double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(C2, r), 1.0);
    y = __dadd_rn(__dmul_rn(z, r2), y);
    return (float)__dmul_rn(y, sc);
}

// sinf / cosf for 0 <= |y| < 120: reduce_fast + sinf_poly of glibc's s_sincosf.h.  want_cos selects the polynomial like cosf does.
__device__ __forceinline__ float glibc_sincosf(float yf, bool want_cos) {
    const double HPI_INV = 0x1.45F306DC9C883p+23, HPI = 0x1.921FB54442D18p0;
    const double c0 = 1.0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10, c4 = 0x1.99343027bf8c3p-16;
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    double x = (double)yf;
    const unsigned top = (__float_as_uint(yf) >> 20) & 0x7ffu;
    int n = 0;
    double sgn = 1.0, flip = 1.0;
    if (top < ((0x3f490fdbu >> 20) & 0x7ffu)) {           // |y| < pi/4 (abstop12 comparison)
        if (top < ((0x39800000u >> 20) & 0x7ffu)) return want_cos ? 1.0f : yf;   // |y| < 2^-12
    } else {
        const double r = __dmul_rn(x, HPI_INV);
        n = (__double2int_rz(r) + 0x800000) >> 24;
        x = __dsub_rn(x, __dmul_rn((double)n, HPI));
        sgn = (n & 3) == 1 || (n & 3) == 2 ? -1.0 : 1.0;   // sign[] = {1, -1, -1, 1}
        if (n & 2) flip = -1.0;                             // __sincosf_table[1]: the cosine coefficients negated
    }
    const int sel = want_cos ? (n ^ 1) : n;
    const double x2 = __dmul_rn(x, x);
    const double xs = __dmul_rn(x, sgn);
    if ((sel & 1) == 0) {
        const double x3 = __dmul_rn(xs, x2);
        const double t1 = __dadd_rn(s2, __dmul_rn(x2, s3));
        const double x7 = __dmul_rn(x3, x2);
        const double s = __dadd_rn(xs, __dmul_rn(x3, s1));
        return (float)__dadd_rn(s, __dmul_rn(x7, t1));
    }
    const double x4 = __dmul_rn(x2, x2);
    const double q2 = __dadd_rn(flip * c3, __dmul_rn(x2, flip * c4));
    const double q1 = __dadd_rn(flip * c1, __dmul_rn(x2, flip * c2));
    const double x6 = __dmul_rn(x4, x2);
    const double c = __dadd_rn(flip * c0, __dmul_rn(x2, q1));
    return (float)__dadd_rn(c, __dmul_rn(x6, q2));
}
__device__ __forceinline__ float glibc_sinf(float y) { return glibc_sincosf(y, false); }
__device__ __forceinline__ float glibc_cosf(float y) { return glibc_sincosf(y, true); }

constexpr int TW = 64, TH = 64;      // blur tile
constexpr int EX_TX = 32, EX_TY = 16;  // extrema tile
constexpr int ORI_BINS = 36;
constexpr int DESC_THREADS = 64;
constexpr int DESC_HIST = 6 * 6 * 10;

__constant__ float c_taps[SIFT_GAUSS][32];
__constant__ float c_exptab[64];

__device__ __forceinline__ int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
    return p;
}

// ---- cvtColor BGR2GRAY 8u ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sift_gray_kernel(const uint8_t* __restrict__ src, int stride, size_t frame_stride,
                                                        uint8_t* __restrict__ dst, int w, int h) {
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y, img = blockIdx.z;
    if (x >= w) return;
    const uint8_t* p = src + (size_t)img * frame_stride + (size_t)y * stride + 3 * x;
    dst[((size_t)img * h + y) * w + x] = (uint8_t)((p[0] * 3735u + p[1] * 19235u + p[2] * 9798u + 16384u) >> 15);
}

// ---- S.3: one sample of resize(gray -> 2x, INTER_LINEAR); weights 0.25 / 0.75, every product and sum exact in fp32 -----
__device__ __forceinline__ float up2_sample(const uint8_t* __restrict__ g, int pitch, int w, int h, int dx, int dy) {
    int sx = (dx >> 1) - ((dx & 1) ^ 1), sy = (dy >> 1) - ((dy & 1) ^ 1);
    float fx = (dx & 1) ? 0.25f : 0.75f, fy = (dy & 1) ? 0.25f : 0.75f;
    if (sx < 0) { fx = 0.f; sx = 0; }
    if (sx >= w - 1) { fx = 0.f; sx = w - 1; }
    if (sy < 0) { fy = 0.f; sy = 0; }
    if (sy >= h - 1) { fy = 0.f; sy = h - 1; }
    const int sx1 = sx + 1 < w ? sx + 1 : w - 1, sy1 = sy + 1 < h ? sy + 1 : h - 1;
    const uint8_t* s0 = g + (size_t)sy * pitch;
    const uint8_t* s1 = g + (size_t)sy1 * pitch;
    const float a1 = fx, a0 = 1.f - fx, b1 = fy, b0 = 1.f - fy;
    const float h0 = (float)s0[sx] * a0 + (float)s0[sx1] * a1;
    const float h1 = (float)s1[sx] * a0 + (float)s1[sx1] * a1;
    return h0 * b0 + h1 * b1;
}

// the whole 2x image (S.3), one thread per 2x2 output block; written once, read once by the first blur
__global__ void __launch_bounds__(256) sift_up2_kernel(const uint8_t* __restrict__ gray, size_t gray_img_stride, int gpitch, int w, int h,
                                                       float* __restrict__ dst, size_t dst_img_stride, int dpitch) {
    const int sx = blockIdx.x * 256 + threadIdx.x, sy = blockIdx.y, img = blockIdx.z;
    if (sx >= w) return;
    const uint8_t* g = gray + (size_t)img * gray_img_stride;
    float* d = dst + (size_t)img * dst_img_stride;
    float2 r0, r1;
    r0.x = up2_sample(g, gpitch, w, h, 2 * sx, 2 * sy);
    r0.y = up2_sample(g, gpitch, w, h, 2 * sx + 1, 2 * sy);
    r1.x = up2_sample(g, gpitch, w, h, 2 * sx, 2 * sy + 1);
    r1.y = up2_sample(g, gpitch, w, h, 2 * sx + 1, 2 * sy + 1);
    *reinterpret_cast<float2*>(d + (size_t)(2 * sy) * dpitch + 2 * sx) = r0;
    *reinterpret_cast<float2*>(d + (size_t)(2 * sy + 1) * dpitch + 2 * sx) = r1;
}

// ---- S.2: GaussianBlur on a float image = sepFilter2D, BORDER_REFLECT_101 ------------------------------------------
//   rows    : s = x0*k0; s = fma(x_i, k_i, s), i ascending            (columns >= W - W%4: mul, then add)
//   columns : s = c*k_R; s = fma(r[+j] + r[-j], k_{R+j}, s), j ascending (columns >= W - W%8: mul, then add)
// Tile = 64 x 64 outputs.  The source tile is staged with a 16-column halo on both sides (so that its rows start 16-byte
// aligned and interior tiles are moved with 128-bit loads and stores) and R rows above and below.
constexpr int BLUR_SW = TW + 32;
template <int R>
constexpr int blur_smem_bytes() { return ((TH + 2 * R) * BLUR_SW + (TH + 2 * R) * TW) * 4; }

template <int R>
__global__ void __launch_bounds__(256) sift_blur_kernel(const float* __restrict__ src, size_t src_img_stride, int src_pitch,
                                                        float* __restrict__ dst, size_t dst_img_stride, int dst_pitch, int W, int H,
                                                        int tap_set) {
    constexpr int N = 2 * R + 1, SH = TH + 2 * R, SW = BLUR_SW;
    constexpr int BASE = (16 - R) & ~3, OFS = (16 - R) & 3, NV = (OFS + N + 3 + 3) / 4;   // row-pass window inside the staged row
    static_assert(R <= 16, "halo");
    extern __shared__ __align__(16) float s_blur[];
    float* s_src = s_blur;
    float* s_row = s_blur + SH * SW;
    const int img = blockIdx.z, tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    const bool interior = tx0 >= 16 && tx0 + TW + 16 <= W && ty0 >= R && ty0 + TH + R <= H;
    if (interior) {
        const float* base = src + (size_t)img * src_img_stride + (size_t)(ty0 - R) * src_pitch + (tx0 - 16);
        // 16-byte asynchronous copies straight into shared memory: every copy of the tile is in flight at once
        for (int e = threadIdx.x; e < SH * (SW / 4); e += 256) {
            const int ry = e / (SW / 4), q = e - ry * (SW / 4);
            const uint32_t sdst = (uint32_t)__cvta_generic_to_shared(s_src + ry * SW + 4 * q);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst), "l"(base + (size_t)ry * src_pitch + 4 * q) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
        for (int e = threadIdx.x; e < SH * SW; e += 256) {
            const int ry = e / SW, rx = e - ry * SW;
            const int gy = reflect101(ty0 - R + ry, H), gx = reflect101(tx0 - 16 + rx, W);
            s_src[e] = src[(size_t)img * src_img_stride + (size_t)gy * src_pitch + gx];
        }
    }
    float k[N];
#pragma unroll
    for (int t = 0; t < N; ++t) k[t] = c_taps[tap_set][t];
    __syncthreads();
    const int row_tail = W - (W & 3), col_tail = W - (W & 7);
    for (int it = threadIdx.x; it < SH * (TW / 4); it += 256) {
        const int ry = it / (TW / 4), x0 = (it - ry * (TW / 4)) * 4;
        const float4* p = reinterpret_cast<const float4*>(s_src + ry * SW + x0 + BASE);
        float v[NV * 4];
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const float4 f = p[q];
            v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
        }
        float a[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) a[m] = v[OFS + m] * k[0];
        if (tx0 + x0 < row_tail) {
#pragma unroll
            for (int t = 1; t < N; ++t)
#pragma unroll
                for (int m = 0; m < 4; ++m) a[m] = __fmaf_rn(v[OFS + t + m], k[t], a[m]);
        } else {
#pragma unroll
            for (int t = 1; t < N; ++t)
#pragma unroll
                for (int m = 0; m < 4; ++m) a[m] = a[m] + v[OFS + t + m] * k[t];
        }
        *reinterpret_cast<float4*>(s_row + ry * TW + x0) = make_float4(a[0], a[1], a[2], a[3]);
    }
    __syncthreads();
    for (int it = threadIdx.x; it < TW * (TH / 4); it += 256) {
        const int x = it % TW, y0 = (it / TW) * 4, gx = tx0 + x;
        if (gx >= W) continue;
        float v[2 * R + 4];
#pragma unroll
        for (int j = 0; j < 2 * R + 4; ++j) v[j] = s_row[(y0 + j) * TW + x];
        const bool fused = gx < col_tail;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            float d = v[m + R] * k[R];
            if (fused) {
#pragma unroll
                for (int j = 1; j <= R; ++j) d = __fmaf_rn(v[m + R + j] + v[m + R - j], k[R + j], d);
            } else {
#pragma unroll
                for (int j = 1; j <= R; ++j) d = d + (v[m + R + j] + v[m + R - j]) * k[R + j];
            }
            const int gy = ty0 + y0 + m;
            if (gy < H) dst[(size_t)img * dst_img_stride + (size_t)gy * dst_pitch + gx] = d;
        }
    }
}

// ---- S.5: resize(INTER_NEAREST) of layer nOctaveLayers to (cols/2, rows/2): sx = min(floor(x * ifx), cols - 1) -------------
__global__ void __launch_bounds__(256) sift_half_kernel(const float* __restrict__ src, size_t img_stride, int spitch, int sw, int sh,
                                                        float* __restrict__ dst, int dpitch, int W, int H, double ifx, double ify) {
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y, img = blockIdx.z;
    if (x >= W) return;
    int sx = (int)floor((double)x * ifx), sy = (int)floor((double)y * ify);
    sx = sx > sw - 1 ? sw - 1 : sx;
    sy = sy > sh - 1 ? sh - 1 : sy;
    dst[(size_t)img * img_stride + (size_t)y * dpitch + x] = src[(size_t)img * img_stride + (size_t)sy * spitch + sx];
}

// ---- S.8 (first half): 26-neighbour extrema of the DoG stack, |val| > threshold, ties allowed ----------------------------------
__device__ __forceinline__ int octave_of_tile(const SiftGeo& g, int t) {
    int o = 0;
    while (o + 1 < g.n_oct && t >= g.oc[o + 1].tile_base) ++o;
    return o;
}

__global__ void __launch_bounds__(256) sift_extrema_kernel(const float* __restrict__ pyr, const __grid_constant__ SiftGeo g,
                                                           uint32_t* __restrict__ cand, int32_t* __restrict__ cand_cnt, float threshold) {
    constexpr int PW = EX_TX + 2, PH = EX_TY + 2;
    __shared__ float s_dog[SIFT_LAYERS + 2][PH * PW];
    const int img = blockIdx.y;
    const int o = octave_of_tile(g, blockIdx.x);
    const SiftOctave& oc = g.oc[o];
    const int t = blockIdx.x - oc.tile_base;
    const int x0 = (t % oc.tiles_x) * EX_TX, y0 = (t / oc.tiles_x) * EX_TY;
    const int W = oc.w, H = oc.h;
    const float* base = pyr + (size_t)img * g.img_floats + oc.off;
    // staging is bound by load latency (a thread stages two or three positions): the loop is unrolled completely and all loads of
    // a thread are issued before the first difference is formed
    constexpr int NE = (PH * PW + 255) / 256;
    float gv[NE][SIFT_GAUSS];
#pragma unroll
    for (int k = 0; k < NE; ++k) {
        const int e = min((int)threadIdx.x + k * 256, PH * PW - 1);
        const int ly = e / PW, lx = e - ly * PW;
        int gy = y0 - 1 + ly, gx = x0 - 1 + lx;
        gy = gy < 0 ? 0 : gy > H - 1 ? H - 1 : gy;
        gx = gx < 0 ? 0 : gx > W - 1 ? W - 1 : gx;
        const float* p = base + (size_t)gy * oc.pitch + gx;
#pragma unroll
        for (int l = 0; l < SIFT_GAUSS; ++l) gv[k][l] = __ldg(p + (size_t)l * oc.layer_stride);
    }
#pragma unroll
    for (int k = 0; k < NE; ++k) {
        const int e = (int)threadIdx.x + k * 256;
        if (e < PH * PW) {
#pragma unroll
            for (int l = 1; l < SIFT_GAUSS; ++l) s_dog[l - 1][e] = gv[k][l] - gv[k][l - 1];
        }
    }
    __syncthreads();
    const int lx = threadIdx.x % EX_TX;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int ly = threadIdx.x / EX_TX + half * (EX_TY / 2);
        const int r = y0 + ly, c = x0 + lx;
        if (r < 5 || r >= H - 5 || c < 5 || c >= W - 5) continue;   // SIFT_IMG_BORDER
        const int ctr = (ly + 1) * PW + lx + 1;
#pragma unroll
        for (int i = 1; i <= SIFT_LAYERS; ++i) {
            const float val = s_dog[i][ctr];
            if (!(fabsf(val) > threshold)) continue;
            bool ext = true;
            if (val > 0) {
#pragma unroll
                for (int dl = -1; dl <= 1; ++dl)
#pragma unroll
                    for (int dr = -1; dr <= 1; ++dr)
#pragma unroll
                        for (int dc = -1; dc <= 1; ++dc) ext = ext && (val >= s_dog[i + dl][ctr + dr * PW + dc]);
            } else {
#pragma unroll
                for (int dl = -1; dl <= 1; ++dl)
#pragma unroll
                    for (int dr = -1; dr <= 1; ++dr)
#pragma unroll
                        for (int dc = -1; dc <= 1; ++dc) ext = ext && (val <= s_dog[i + dl][ctr + dr * PW + dc]);
            }
            if (!ext) continue;
            const int slot = atomicAdd(cand_cnt + img, 1);
            if (slot < g.cand_cap) cand[(size_t)img * g.cand_cap + slot] = ((uint32_t)o << 28) | ((uint32_t)i << 26) | ((uint32_t)r << 13) | (uint32_t)c;
        }
    }
}

// ---- S.4: hal leaf functions as OpenCV's SIMD build evaluates them ---------------------------------------------------------
__device__ __forceinline__ float sift_exp32f(float x) {   // cv::hal::exp32f: 64-entry table * cubic in the remainder
    const float prescale = (float)(1.4426950408889634073599246810019 * 64);
    const float postscale = (float)(1. / 64);
    const float A0 = (float)(1.000000000000002438532970795181890933776 / .9670371139572337719125840413672004409288e-2);   // A4 of the source
    const float A3 = (float)(.6931471805521448196800669615864773144641 / .9670371139572337719125840413672004409288e-2);
    const float A2 = (float)(.2402265109513301490103372422686535526573 / .9670371139572337719125840413672004409288e-2);
    const float A1 = (float)(.5550339366753125211915322047004666939128e-1 / .9670371139572337719125840413672004409288e-2);
    const float maxval = (float)(3000. * 64 / (1.4426950408889634073599246810019 * 64)), minval = -maxval;
    float x0 = fminf(fmaxf(x, minval), maxval);
    x0 = x0 * prescale;
    const int xi = __float2int_rn(x0);
    x0 = (x0 - (float)xi) * postscale;
    int t = (xi >> 6) + 127;
    t = t < 0 ? 0 : t > 255 ? 255 : t;
    const float y = c_exptab[xi & 63] * __int_as_float(t << 23);
    float z = x0 + A1;
    z = __fmaf_rn(z, x0, A2);
    z = __fmaf_rn(z, x0, A3);
    z = __fmaf_rn(z, x0, A0);
    return z * y;
}

// cv::hal::fastAtan2 (v_atan_f32, degrees) and cv::hal::magnitude32f = sqrtf(fma(x, x, y * y)) are evaluated in stages, see below
// fastAtan2 in two halves around its division, so that the IEEE divisions (and square roots) of several window samples can be
// issued back to back: each carries a rarely-taken slow-path branch that stops the compiler from interleaving whole samples
__device__ __forceinline__ float sift_atan2_post(float c, float y, float x) {
    const float RAD = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * RAD, p3 = -0.3258083974640975f * RAD, p5 = 0.1555786518463281f * RAD,
                p7 = -0.04432655554792128f * RAD;
    const float ax = fabsf(x), ay = fabsf(y);
    const float cc = c * c;
    float a = __fmaf_rn(__fmaf_rn(__fmaf_rn(cc, p7, p5), cc, p3), cc, p1) * c;
    if (!(ax >= ay)) a = 90.f - a;
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

struct OctView {
    const float* base;   // Gaussian layer 0
    size_t ls;
    int pitch, W, H;
    __device__ __forceinline__ float G(int l, int r, int c) const { return base[(size_t)l * ls + (size_t)r * pitch + c]; }
    __device__ __forceinline__ float D(int l, int r, int c) const { return G(l + 1, r, c) - G(l, r, c); }   // DoG layer l
};

#define SIFT_FMS(p, q, r, s) __fmaf_rn((p), (q), -((r) * (s)))
#define SIFT_OUT3(p, m1, q, m2, r, m3) __fmaf_rn((r), (m3), __fmaf_rn((p), (m1), -((q) * (m2))))

// ---- S.6: adjustLocalExtrema, one thread per candidate; survivors go to a compact list so that the (long) orientation loops
//      of the next kernel run on full warps -----------------------------------------------------------------------------------
struct RefinedCand {
    float kx, ky, ksize, kresp;
    int koct, packed;   // packed = octave << 28 | layer << 26 | r << 13 | c  (refined integer position)
};

__global__ void __launch_bounds__(128) sift_refine_kernel(const float* __restrict__ pyr, const __grid_constant__ SiftGeo g,
                                                          const uint32_t* __restrict__ cand, const int32_t* __restrict__ cand_cnt,
                                                          RefinedCand* __restrict__ refined, int32_t* __restrict__ refined_cnt) {
    const int img = blockIdx.y, tid = threadIdx.x;
    const int n = min(cand_cnt[img], g.cand_cap);
    const int ci = blockIdx.x * 128 + tid;
    if (ci >= n) return;
    const uint32_t packed = cand[(size_t)img * g.cand_cap + ci];
    const int octv = packed >> 28;
    int l = (packed >> 26) & 3, r = (packed >> 13) & 8191, c = packed & 8191;
    const SiftOctave& oc = g.oc[octv];
    OctView V{pyr + (size_t)img * g.img_floats + oc.off, oc.layer_stride, oc.pitch, oc.w, oc.h};
    const int W = oc.w, H = oc.h;

    // <= 5 Newton steps on the 3-D quadratic, then the contrast and edge tests
    const float img_scale = 1.f / 255, deriv_scale = img_scale * 0.5f, second_deriv_scale = img_scale, cross_deriv_scale = img_scale * 0.25f;
    float xi = 0, xr = 0, xc = 0;
    int it = 0;
    for (; it < 5; ++it) {
        const float v = V.D(l, r, c);
        const float cxp = V.D(l, r, c + 1), cxm = V.D(l, r, c - 1), cyp = V.D(l, r + 1, c), cym = V.D(l, r - 1, c);
        const float nx = V.D(l + 1, r, c), pv = V.D(l - 1, r, c);
        const float dD0 = (cxp - cxm) * deriv_scale, dD1 = (cyp - cym) * deriv_scale, dD2 = (nx - pv) * deriv_scale;
        const float v2 = v * 2;
        const float dxx = (cxp + cxm - v2) * second_deriv_scale, dyy = (cyp + cym - v2) * second_deriv_scale, dss = (nx + pv - v2) * second_deriv_scale;
        const float dxy = (V.D(l, r + 1, c + 1) - V.D(l, r + 1, c - 1) - V.D(l, r - 1, c + 1) + V.D(l, r - 1, c - 1)) * cross_deriv_scale;
        const float dxs = (V.D(l + 1, r, c + 1) - V.D(l + 1, r, c - 1) - V.D(l - 1, r, c + 1) + V.D(l - 1, r, c - 1)) * cross_deriv_scale;
        const float dys = (V.D(l + 1, r + 1, c) - V.D(l + 1, r - 1, c) - V.D(l - 1, r + 1, c) + V.D(l - 1, r - 1, c)) * cross_deriv_scale;
        // Matx33f H(dxx, dxy, dxs, dxy, dyy, dys, dxs, dys, dss); X = H.solve(dD, DECOMP_LU) = Cramer's rule in float, contracted
        const float a00 = dxx, a01 = dxy, a02 = dxs, a10 = dxy, a11 = dyy, a12 = dys, a20 = dxs, a21 = dys, a22 = dss;
        const float det = SIFT_OUT3(a00, SIFT_FMS(a11, a22, a21, a12), a01, SIFT_FMS(a10, a22, a20, a12), a02, SIFT_FMS(a10, a21, a20, a11));
        float X0 = 0, X1 = 0, X2 = 0;
        if (det != 0) {
            const float d = 1 / det;
            X0 = d * SIFT_OUT3(dD0, SIFT_FMS(a11, a22, a12, a21), a01, SIFT_FMS(dD1, a22, a12, dD2), a02, SIFT_FMS(dD1, a21, a11, dD2));
            X1 = d * SIFT_OUT3(a00, SIFT_FMS(dD1, a22, a12, dD2), dD0, SIFT_FMS(a10, a22, a12, a20), a02, SIFT_FMS(a10, dD2, dD1, a20));
            X2 = d * SIFT_OUT3(a00, SIFT_FMS(a11, dD2, dD1, a21), a01, SIFT_FMS(a10, dD2, dD1, a20), dD0, SIFT_FMS(a10, a21, a11, a20));
        }
        xi = -X2;
        xr = -X1;
        xc = -X0;
        if (fabsf(xi) < 0.5f && fabsf(xr) < 0.5f && fabsf(xc) < 0.5f) break;
        const float big = (float)(2147483647 / 3);
        if (fabsf(xi) > big || fabsf(xr) > big || fabsf(xc) > big) return;
        c += __float2int_rn(xc);
        r += __float2int_rn(xr);
        l += __float2int_rn(xi);
        if (l < 1 || l > SIFT_LAYERS || c < 5 || c >= W - 5 || r < 5 || r >= H - 5) return;
    }
    if (it >= 5) return;
    float contr;
    {
        const float v = V.D(l, r, c);
        const float cxp = V.D(l, r, c + 1), cxm = V.D(l, r, c - 1), cyp = V.D(l, r + 1, c), cym = V.D(l, r - 1, c);
        const float dD0 = (cxp - cxm) * deriv_scale, dD1 = (cyp - cym) * deriv_scale, dD2 = (V.D(l + 1, r, c) - V.D(l - 1, r, c)) * deriv_scale;
        const float t = __fmaf_rn(dD2, xi, __fmaf_rn(dD1, xr, dD0 * xc));
        contr = __fmaf_rn(v, img_scale, t * 0.5f);
        if (fabsf(contr) * SIFT_LAYERS < (float)0.04) return;
        const float v2 = v * 2.f;
        const float dxx = (cxp + cxm - v2) * second_deriv_scale, dyy = (cyp + cym - v2) * second_deriv_scale;
        const float dxy = (V.D(l, r + 1, c + 1) - V.D(l, r + 1, c - 1) - V.D(l, r - 1, c + 1) + V.D(l, r - 1, c - 1)) * cross_deriv_scale;
        const float tr = dxx + dyy;
        const float det = SIFT_FMS(dxx, dyy, dxy, dxy);
        const float e = 10.f;
        if (det <= 0 || tr * tr * e >= (e + 1) * (e + 1) * det) return;
    }
    RefinedCand rc;
    rc.kx = ((float)c + xc) * (float)(1 << octv);
    rc.ky = ((float)r + xr) * (float)(1 << octv);
    rc.koct = octv + (l << 8) + (__double2int_rn(((double)xi + 0.5) * 255) << 16);
    rc.ksize = (float)1.6 * glibc_exp2f(((float)l + xi) / SIFT_LAYERS) * (float)(1 << octv) * 2;
    rc.kresp = fabsf(contr);
    rc.packed = (int)(((uint32_t)octv << 28) | ((uint32_t)l << 26) | ((uint32_t)r << 13) | (uint32_t)c);
    const int slot = atomicAdd(refined_cnt + img, 1);
    if (slot < g.kp_cap) refined[(size_t)img * g.kp_cap + slot] = rc;
}

// ---- S.7 + S.8 (second half): calcOrientationHist on Gaussian layer l (sums in raster order) and the orientation peaks;
//      one thread per refined candidate, its 36 bins in shared memory; two window samples are evaluated per iteration (ILP) ----
__global__ void __launch_bounds__(128) sift_orient_kernel(const float* __restrict__ pyr, const __grid_constant__ SiftGeo g,
                                                          const RefinedCand* __restrict__ refined, const int32_t* __restrict__ refined_cnt,
                                                          float* __restrict__ raw, int32_t* __restrict__ raw_cnt) {
    __shared__ float s_tmp[ORI_BINS][128];
    __shared__ float s_hist[ORI_BINS][128];
    const int img = blockIdx.y, tid = threadIdx.x;
    const int n = min(refined_cnt[img], g.kp_cap);
    const int ci = blockIdx.x * 128 + tid;
    if (ci >= n) return;
    const RefinedCand rc = refined[(size_t)img * g.kp_cap + ci];
    const uint32_t packed = (uint32_t)rc.packed;
    const int octv = packed >> 28, l = (packed >> 26) & 3, r = (packed >> 13) & 8191, c = packed & 8191;
    const SiftOctave& oc = g.oc[octv];
    const int W = oc.w, H = oc.h, pitch = oc.pitch;
    const float* im = pyr + (size_t)img * g.img_floats + oc.off + (size_t)l * oc.layer_stride;
    const float scl_octv = rc.ksize * 0.5f / (float)(1 << octv);
    const int radius = __float2int_rn(4.5f * scl_octv);
    const float sigma = 1.5f * scl_octv;
    const float expf_scale = -1.f / (2.f * sigma * sigma);
#pragma unroll
    for (int b = 0; b < ORI_BINS; ++b) s_tmp[b][tid] = 0.f;
    const int j_lo = max(-radius, 1 - c), j_hi = min(radius, W - 2 - c);   // 0 < x < W - 1
    for (int i = -radius; i <= radius; ++i) {
        const int y = r + i;
        if (y <= 0 || y >= H - 1) continue;
        // four window samples per iteration from 16-byte aligned loads of the three rows (a quarter of the L1 lookups that
        // per-sample scalar loads cost, and 4-way ILP); samples outside [j_lo, j_hi] are computed and dropped
        // The loads of the next four columns are issued before the current four are worked on (one thread per keypoint: nothing
        // else covers the load latency).
        const float* rowm = im + (size_t)y * pitch;
        const int c4_end = c + j_hi;
        int c4 = (c + j_lo) & ~3;
        float4 up_n = make_float4(0.f, 0.f, 0.f, 0.f), mid_n = up_n, dn_n = up_n;
        float left_n = 0.f, right_n = 0.f;
        auto fetch = [&](int cc) {
            up_n = __ldg(reinterpret_cast<const float4*>(rowm - pitch + cc));
            mid_n = __ldg(reinterpret_cast<const float4*>(rowm + cc));
            dn_n = __ldg(reinterpret_cast<const float4*>(rowm + pitch + cc));
            left_n = cc > 0 ? __ldg(rowm + cc - 1) : 0.f;
            right_n = cc + 4 < pitch ? __ldg(rowm + cc + 4) : 0.f;
        };
        if (c4 <= c4_end) fetch(c4);
        for (; c4 <= c4_end; c4 += 4) {
            const float4 up = up_n, mid = mid_n, dn = dn_n;
            const float left = left_n, right = right_n;
            if (c4 + 4 <= c4_end) fetch(c4 + 4);
            const float dxs[4] = {mid.y - left, mid.z - mid.x, mid.w - mid.y, right - mid.z};
            const float dys[4] = {up.x - dn.x, up.y - dn.y, up.z - dn.z, up.w - dn.w};
            float wm[4], num[4], den[4], q[4], sq[4];
            int bn[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) {   // stage 1: weights and the operands of the division / square root
                const int j = c4 + m - c;
                wm[m] = sift_exp32f((float)(i * i + j * j) * expf_scale);
                const float ax = fabsf(dxs[m]), ay = fabsf(dys[m]);
                num[m] = fminf(ax, ay);
                den[m] = fmaxf(ax, ay) + (float)2.2204460492503131e-16;
                sq[m] = __fmaf_rn(dxs[m], dxs[m], dys[m] * dys[m]);
            }
#pragma unroll
            for (int m = 0; m < 4; ++m) q[m] = num[m] / den[m];          // stage 2: the four divisions
#pragma unroll
            for (int m = 0; m < 4; ++m) sq[m] = sqrtf(sq[m]);            // stage 3: the four square roots (magnitude32f)
#pragma unroll
            for (int m = 0; m < 4; ++m) {   // stage 4: angle polynomial, bin, weight
                const int j = c4 + m - c;
                const float o = sift_atan2_post(q[m], dys[m], dxs[m]);
                int bb = __float2int_rn((ORI_BINS / 360.f) * o);
                if (bb >= ORI_BINS) bb -= ORI_BINS;
                if (bb < 0) bb += ORI_BINS;
                const bool ok = j >= j_lo && j <= j_hi;
                bn[m] = ok ? bb : -1;
                wm[m] = wm[m] * sq[m];
            }
#pragma unroll
            for (int m = 0; m < 4; ++m)
                if (bn[m] >= 0) s_tmp[bn[m]][tid] += wm[m];
        }
    }
    float maxval = 0.f;
    for (int b = 0; b < ORI_BINS; ++b) {   // smoothing [1 4 6 4 1]/16: bins 0..31 in the SIMD loop (nested fma), 32..35 in its scalar tail
        const int m2 = b - 2 < 0 ? b - 2 + ORI_BINS : b - 2, m1 = b - 1 < 0 ? b - 1 + ORI_BINS : b - 1;
        const int p1 = b + 1 >= ORI_BINS ? b + 1 - ORI_BINS : b + 1, p2 = b + 2 >= ORI_BINS ? b + 2 - ORI_BINS : b + 2;
        const float sa = s_tmp[m2][tid] + s_tmp[p2][tid], sb = s_tmp[m1][tid] + s_tmp[p1][tid], sc = s_tmp[b][tid];
        float hv;
        if (b < 32) hv = __fmaf_rn(sa, 1.f / 16.f, __fmaf_rn(sb, 4.f / 16.f, sc * (6.f / 16.f)));
        else hv = __fmaf_rn(sc, 6.f / 16.f, sa * (1.f / 16.f) + sb * (4.f / 16.f));
        s_hist[b][tid] = hv;
        maxval = b == 0 ? hv : (maxval > hv ? maxval : hv);
    }
    const float mag_thr = maxval * 0.8f;
    for (int j = 0; j < ORI_BINS; ++j) {
        const int lft = j > 0 ? j - 1 : ORI_BINS - 1, rgt = j < ORI_BINS - 1 ? j + 1 : 0;
        const float hj = s_hist[j][tid], hl = s_hist[lft][tid], hr = s_hist[rgt][tid];
        if (hj > hl && hj > hr && hj >= mag_thr) {
            float bin = (float)j + 0.5f * (hl - hr) / (hl - 2 * hj + hr);
            bin = bin < 0 ? ORI_BINS + bin : bin >= ORI_BINS ? bin - ORI_BINS : bin;
            float angle = __fmaf_rn(-(360.f / ORI_BINS), bin, 360.f);
            if (fabsf(angle - 360.f) < 1.1920928955078125e-7f) angle = 0.f;
            const int slot = atomicAdd(raw_cnt + img, 1);
            if (slot < g.kp_cap) {
                float* o = raw + ((size_t)img * g.kp_cap + slot) * 6;
                o[0] = rc.kx; o[1] = rc.ky; o[2] = rc.ksize; o[3] = angle; o[4] = rc.kresp; o[5] = __int_as_float(rc.koct);
            }
        }
    }
}

// ---- S.9: KeyPoint_LessThan rank of every keypoint (rank counting: one thread per keypoint, the list streamed through smem) ----
struct KpRec { float x, y, size, angle, resp; int oct; };
__device__ __forceinline__ bool kp_less(const KpRec& a, const KpRec& b) {
    if (a.x != b.x) return a.x < b.x;
    if (a.y != b.y) return a.y < b.y;
    if (a.size != b.size) return a.size > b.size;
    if (a.angle != b.angle) return a.angle < b.angle;
    if (a.resp != b.resp) return a.resp > b.resp;
    if (a.oct != b.oct) return a.oct > b.oct;
    return false;
}

__global__ void __launch_bounds__(256) sift_rank_kernel(const float* __restrict__ raw, const int32_t* __restrict__ raw_cnt, int kp_cap,
                                                        int32_t* __restrict__ order) {
    constexpr int TILE = 1024;
    __shared__ __align__(16) float s_x[TILE];   // pt.x of the tile: the comparison that decides almost every pair
    const int img = blockIdx.y;
    const int n = min(raw_cnt[img], kp_cap);
    if (blockIdx.x * 256 >= n) return;
    const float* list = raw + (size_t)img * kp_cap * 6;
    const int i = blockIdx.x * 256 + threadIdx.x;
    KpRec me{};
    if (i < n) me = KpRec{list[i * 6], list[i * 6 + 1], list[i * 6 + 2], list[i * 6 + 3], list[i * 6 + 4], __float_as_int(list[i * 6 + 5])};
    int rank = 0;
    for (int t0 = 0; t0 < n; t0 += TILE) {
        const int tn = min(TILE, n - t0);
        __syncthreads();
        for (int e = threadIdx.x; e < TILE; e += 256) s_x[e] = e < tn ? list[(size_t)(t0 + e) * 6] : __int_as_float(0x7f800000);   // +inf: never less, never equal
        __syncthreads();
        if (i < n) {
            const float4* x4 = reinterpret_cast<const float4*>(s_x);
            for (int j4 = 0; j4 < (tn + 3) / 4; ++j4) {
                const float4 o = x4[j4];
                rank += (o.x < me.x) + (o.y < me.x) + (o.z < me.x) + (o.w < me.x);
                if (o.x == me.x || o.y == me.x || o.z == me.x || o.w == me.x) {   // rare: same pt.x -> the full KeyPoint_LessThan order
                    const float ox[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int j = t0 + j4 * 4 + q;
                        if (ox[q] != me.x || j >= n) continue;
                        const float* r = list + (size_t)j * 6;
                        const KpRec other{r[0], r[1], r[2], r[3], r[4], __float_as_int(r[5])};
                        if (kp_less(other, me) || (!kp_less(me, other) && j < i)) ++rank;
                    }
                }
            }
        }
    }
    if (i < n) order[(size_t)img * kp_cap + rank] = i;
}

// descriptor window radius of a keypoint (S.10 / S.11), shared by the bucket kernels and the descriptor kernel
__device__ __forceinline__ int desc_radius(float ksize, int koct, const SiftGeo& g, int* o_out, int* layer_out, float* scale_out) {
    int octave = koct & 255;
    *layer_out = (koct >> 8) & 255;
    octave = octave < 128 ? octave : (-128 | octave);
    const float scale = octave >= 0 ? 1.f / (float)(1 << octave) : (float)(1 << -octave);
    const float size = ksize * scale;
    const float scl = size * 0.5f;
    const float hist_width = 3.f * scl;
    int radius = __float2int_rn(hist_width * 1.4142135623730951f * 5 * 0.5f);
    const int o = octave + 1;
    radius = radius > g.oc[o].diag ? g.oc[o].diag : radius;
    *o_out = o;
    *scale_out = scale;
    return radius;
}
__device__ __forceinline__ int radius_bucket(int radius) { return 63 - (radius > 63 ? 63 : radius); }

__global__ void __launch_bounds__(1024) sift_unique_kernel(const float* __restrict__ raw, const int32_t* __restrict__ raw_cnt, int kp_cap,
                                                           const int32_t* __restrict__ order, float* __restrict__ uniq, int32_t* __restrict__ uniq_cnt) {
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = min(raw_cnt[img], kp_cap);
    const float* list = raw + (size_t)img * kp_cap * 6;
    const int32_t* ord = order + (size_t)img * kp_cap;
    float* out = uniq + (size_t)img * kp_cap * 6;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int p0 = 0; p0 < n; p0 += 1024) {
        const int p = p0 + tid;
        int keep = 0;
        const float* a = nullptr;
        if (p < n) {
            a = list + (size_t)ord[p] * 6;
            keep = 1;
            if (p > 0) {
                const float* b = list + (size_t)ord[p - 1] * 6;
                keep = (a[0] != b[0] || a[1] != b[1] || a[2] != b[2] || a[3] != b[3]) ? 1 : 0;
            }
        }
        int incl = keep;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xFFFFFFFFu, w, d);
                if (lane >= d) w += v;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int base = s_base + (wid > 0 ? s_warp[wid - 1] : 0);
        if (keep) {
            float* o = out + (size_t)(base + incl - 1) * 6;
            const int oct = __float_as_int(a[5]);
            o[0] = a[0] * 0.5f; o[1] = a[1] * 0.5f; o[2] = a[2] * 0.5f; o[3] = a[3]; o[4] = a[4];
            o[5] = __int_as_float((oct & ~255) | ((oct - 1) & 255));
        }
        __syncthreads();
        if (tid == 1023) s_base = base + incl;
        __syncthreads();
    }
    if (tid == 0) uniq_cnt[img] = s_base;
}

// ---- scheduling order of the descriptor threads: keypoints bucketed by window radius, largest first, so that the threads of
//      a warp run loops of (nearly) equal length.  perm entry = image << 16 | index within the image. ----------------------------
__global__ void __launch_bounds__(256) sift_bucket_count_kernel(const float* __restrict__ uniq, const int32_t* __restrict__ uniq_cnt,
                                                                const __grid_constant__ SiftGeo g, int32_t* __restrict__ bucket_cnt) {
    __shared__ int s_cnt[64];
    const int img = blockIdx.y, j = blockIdx.x * 256 + threadIdx.x;
    if (blockIdx.x * 256 >= uniq_cnt[img]) return;
    if (threadIdx.x < 64) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    if (j < uniq_cnt[img]) {
        const float* kp = uniq + ((size_t)img * g.kp_cap + j) * 6;
        int o, layer;
        float scale;
        atomicAdd(&s_cnt[radius_bucket(desc_radius(kp[2], __float_as_int(kp[5]), g, &o, &layer, &scale))], 1);
    }
    __syncthreads();
    if (threadIdx.x < 64 && s_cnt[threadIdx.x]) atomicAdd(bucket_cnt + threadIdx.x, s_cnt[threadIdx.x]);
}

__global__ void sift_bucket_scan_kernel(const int32_t* __restrict__ bucket_cnt, int32_t* __restrict__ bucket_cursor) {
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int b = 0; b < 64; ++b) {
            bucket_cursor[b] = acc;
            acc += bucket_cnt[b];
        }
    }
}

__global__ void __launch_bounds__(256) sift_bucket_scatter_kernel(const float* __restrict__ uniq, const int32_t* __restrict__ uniq_cnt,
                                                                  const __grid_constant__ SiftGeo g, int32_t* __restrict__ bucket_cursor,
                                                                  uint32_t* __restrict__ perm) {
    const int img = blockIdx.y, j = blockIdx.x * 256 + threadIdx.x;
    if (j >= uniq_cnt[img]) return;
    const float* kp = uniq + ((size_t)img * g.kp_cap + j) * 6;
    int o, layer;
    float scale;
    const int b = radius_bucket(desc_radius(kp[2], __float_as_int(kp[5]), g, &o, &layer, &scale));
    perm[atomicAdd(bucket_cursor + b, 1)] = ((uint32_t)img << 16) | (uint32_t)j;
}

// ---- S.10 + S.11: calcSIFTDescriptor, one thread per keypoint; the 6x6x10 histogram of thread t lives in s_h[bin * 64 + t].
//      Per window row only the columns that can pass the rotated-window test are visited (a conservative interval, the exact
//      float test decides inside it), two samples per iteration for ILP; histogram updates stay in raster order. ----------------
struct DescSample {
    bool ok;
    int idx;
    float mag, rb, cb, ob;
};

__global__ void __launch_bounds__(DESC_THREADS) sift_descriptor_kernel(const float* __restrict__ pyr, const __grid_constant__ SiftGeo g,
                                                                       const float* __restrict__ uniq, const int32_t* __restrict__ frame_off,
                                                                       const uint32_t* __restrict__ perm, int total, float* __restrict__ kp_f,
                                                                       int32_t* __restrict__ kp_oct, int32_t* __restrict__ q_frame,
                                                                       float* __restrict__ desc) {
    extern __shared__ float s_h[];
    __shared__ int s_row[DESC_THREADS];
    const int tid = threadIdx.x;
    const int slot = blockIdx.x * DESC_THREADS + tid;
    const bool active = slot < total;
    s_row[tid] = -1;
    if (active) {
        const uint32_t pe = perm[slot];
        const int img = pe >> 16, jj = pe & 0xFFFF;
        const int gi = frame_off[img] + jj;
        s_row[tid] = gi;
        const float* kp = uniq + ((size_t)img * g.kp_cap + jj) * 6;
        const float kx = kp[0], ky = kp[1], ksize = kp[2], kangle = kp[3], kresp = kp[4];
        const int koct = __float_as_int(kp[5]);
        kp_f[(size_t)gi * 5] = kx; kp_f[(size_t)gi * 5 + 1] = ky; kp_f[(size_t)gi * 5 + 2] = ksize; kp_f[(size_t)gi * 5 + 3] = kangle;
        kp_f[(size_t)gi * 5 + 4] = kresp;
        kp_oct[gi] = koct;
        q_frame[gi] = img;
        // unpackOctave; the image of (octave - firstOctave, layer)
        int o, layer;
        float scale;
        const int radius = desc_radius(ksize, koct, g, &o, &layer, &scale);
        const float size = ksize * scale;
        float ori = 360.f - kangle;
        if (fabsf(ori - 360.f) < 1.1920928955078125e-7f) ori = 0.f;
        const SiftOctave& oc = g.oc[o];
        const int W = oc.w, H = oc.h;
        const float* im = pyr + (size_t)img * g.img_floats + oc.off + (size_t)layer * oc.layer_stride;
        const int pitch = oc.pitch;
        const float ptx = kx * scale, pty = ky * scale, scl = size * 0.5f;
        const int d = 4, n = 8;
        const int px = __float2int_rn(ptx), py = __float2int_rn(pty);
        const float ang = ori * (float)(3.14159265358979323846 / 180);
        float cos_t = glibc_cosf(ang), sin_t = glibc_sinf(ang);
        const float bins_per_rad = n / 360.f;
        const float exp_scale = -1.f / (d * d * 0.5f);
        const float hist_width = 3.f * scl;
        cos_t = cos_t / hist_width;
        sin_t = sin_t / hist_width;
        for (int b = 0; b < DESC_HIST; ++b) s_h[b * DESC_THREADS + tid] = 0.f;
        // conservative column interval of a window row: |j*sin + i*cos| and |j*cos - i*sin| must stay below ~2.5; a constraint
        // whose slope is tiny is skipped (the exact test then decides), otherwise one column of slack covers the float rounding
        const bool use_s = fabsf(sin_t) >= 1e-3f, use_c = fabsf(cos_t) >= 1e-3f;
        const float inv_s = use_s ? 1.f / sin_t : 0.f, inv_c = use_c ? 1.f / cos_t : 0.f;
        const int c_lo = max(-radius, 1 - px), c_hi = min(radius, W - 2 - px);   // 0 < c < W - 1
        for (int i = -radius; i <= radius; ++i) {
            const int r = py + i;
            if (r <= 0 || r >= H - 1) continue;
            const float is = (float)i * sin_t, ic = (float)i * cos_t;
            int jlo = c_lo, jhi = c_hi;
            if (use_s) {   // -2.6 < j*sin_t + ic < 2.6
                const float a = (-2.6f - ic) * inv_s, b = (2.6f - ic) * inv_s;
                jlo = max(jlo, __float2int_rd(fminf(a, b)) - 1);
                jhi = min(jhi, __float2int_ru(fmaxf(a, b)) + 1);
            }
            if (use_c) {   // -2.6 < j*cos_t - is < 2.6
                const float a = (-2.6f + is) * inv_c, b = (2.6f + is) * inv_c;
                jlo = max(jlo, __float2int_rd(fminf(a, b)) - 1);
                jhi = min(jhi, __float2int_ru(fmaxf(a, b)) + 1);
            }
            const float* rowm = im + (size_t)r * pitch;
            auto apply = [&](const DescSample& sm) {
                if (!sm.ok) return;
                const float v_r1 = sm.mag * sm.rb, v_r0 = sm.mag - v_r1;
                const float v_rc11 = v_r1 * sm.cb, v_rc10 = v_r1 - v_rc11;
                const float v_rc01 = v_r0 * sm.cb, v_rc00 = v_r0 - v_rc01;
                const float v_rco111 = v_rc11 * sm.ob, v_rco110 = v_rc11 - v_rco111;
                const float v_rco101 = v_rc10 * sm.ob, v_rco100 = v_rc10 - v_rco101;
                const float v_rco011 = v_rc01 * sm.ob, v_rco010 = v_rc01 - v_rco011;
                const float v_rco001 = v_rc00 * sm.ob, v_rco000 = v_rc00 - v_rco001;
                float* hp = s_h + sm.idx * DESC_THREADS + tid;
                hp[0] += v_rco000;
                hp[1 * DESC_THREADS] += v_rco001;
                hp[(n + 2) * DESC_THREADS] += v_rco010;
                hp[(n + 3) * DESC_THREADS] += v_rco011;
                hp[(d + 2) * (n + 2) * DESC_THREADS] += v_rco100;
                hp[((d + 2) * (n + 2) + 1) * DESC_THREADS] += v_rco101;
                hp[(d + 3) * (n + 2) * DESC_THREADS] += v_rco110;
                hp[((d + 3) * (n + 2) + 1) * DESC_THREADS] += v_rco111;
            };
            // four samples per iteration from 16-byte aligned loads of the three image rows (see sift_orient_kernel)
            // (the loads of the next four columns are issued before these four are worked on: one thread per keypoint and 4 warps
            //  per SM leave nothing else to cover the load latency -- 40 % of the kernel's stall samples sat on the first use)
            const int c4_end = px + jhi;
            int c4 = (px + jlo) & ~3;
            float4 up_n = make_float4(0.f, 0.f, 0.f, 0.f), mid_n = up_n, dn_n = up_n;
            float left_n = 0.f, right_n = 0.f;
            auto fetch = [&](int c) {
                up_n = __ldg(reinterpret_cast<const float4*>(rowm - pitch + c));
                mid_n = __ldg(reinterpret_cast<const float4*>(rowm + c));
                dn_n = __ldg(reinterpret_cast<const float4*>(rowm + pitch + c));
                left_n = c > 0 ? __ldg(rowm + c - 1) : 0.f;
                right_n = c + 4 < pitch ? __ldg(rowm + c + 4) : 0.f;
            };
            if (c4 <= c4_end) fetch(c4);
            for (; c4 <= c4_end; c4 += 4) {
                const float4 up = up_n, mid = mid_n, dn = dn_n;
                const float left = left_n, right = right_n;
                if (c4 + 4 <= c4_end) fetch(c4 + 4);
                const int j0 = c4 - px;
                const float dxs[4] = {mid.y - left, mid.z - mid.x, mid.w - mid.y, right - mid.z};
                const float dys[4] = {up.x - dn.x, up.y - dn.y, up.z - dn.z, up.w - dn.w};
                DescSample sm[4];
                float wgt[4], num[4], den[4], q[4], sq[4], rbin[4], cbin[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) {   // stage 1: rotated coordinates, window test, weight, operands of the division / square root
                    const int j = j0 + m;
                    const float c_rot = (float)j * cos_t - is;
                    const float r_rot = (float)j * sin_t + ic;
                    rbin[m] = r_rot + (float)(d / 2) - 0.5f;
                    cbin[m] = c_rot + (float)(d / 2) - 0.5f;
                    sm[m].ok = j >= jlo && j <= jhi && rbin[m] > -1 && rbin[m] < d && cbin[m] > -1 && cbin[m] < d;
                    wgt[m] = sift_exp32f((c_rot * c_rot + r_rot * r_rot) * exp_scale);
                    const float ax = fabsf(dxs[m]), ay = fabsf(dys[m]);
                    num[m] = fminf(ax, ay);
                    den[m] = fmaxf(ax, ay) + (float)2.2204460492503131e-16;
                    sq[m] = __fmaf_rn(dxs[m], dxs[m], dys[m] * dys[m]);
                }
#pragma unroll
                for (int m = 0; m < 4; ++m) q[m] = num[m] / den[m];      // stage 2: the four divisions of fastAtan2
#pragma unroll
                for (int m = 0; m < 4; ++m) sq[m] = sqrtf(sq[m]);        // stage 3: the four square roots of magnitude32f
#pragma unroll
                for (int m = 0; m < 4; ++m) {   // stage 4: angle, bins, interpolation weights
                    const float og = sift_atan2_post(q[m], dys[m], dxs[m]);
                    sm[m].mag = sq[m] * wgt[m];
                    float obin = (og - ori) * bins_per_rad;
                    const int r0 = __float2int_rd(rbin[m]), c0 = __float2int_rd(cbin[m]);
                    int o0 = __float2int_rd(obin);
                    sm[m].rb = rbin[m] - (float)r0;
                    sm[m].cb = cbin[m] - (float)c0;
                    sm[m].ob = obin - (float)o0;
                    if (o0 < 0) o0 += n;
                    if (o0 >= n) o0 -= n;
                    sm[m].idx = ((r0 + 1) * (d + 2) + c0 + 1) * (n + 2) + o0;
                }
#pragma unroll
                for (int m = 0; m < 4; ++m) apply(sm[m]);                // stage 5: histogram updates, in raster order
            }
        }
        // circular orientation bins, then the descriptor is normalised in place (bins k < 8 of the 16 inner cells)
        float nrm2 = 0;
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) {
                float* hp = s_h + (((i + 1) * (d + 2) + (j + 1)) * (n + 2)) * DESC_THREADS + tid;
                hp[0] += hp[n * DESC_THREADS];
                hp[DESC_THREADS] += hp[(n + 1) * DESC_THREADS];
                for (int k = 0; k < n; ++k) nrm2 = __fmaf_rn(hp[k * DESC_THREADS], hp[k * DESC_THREADS], nrm2);
            }
        const float thr = sqrtf(nrm2) * 0.2f;
        nrm2 = 0;
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) {
                float* hp = s_h + (((i + 1) * (d + 2) + (j + 1)) * (n + 2)) * DESC_THREADS + tid;
                for (int k = 0; k < n; ++k) {
                    const float raw = hp[k * DESC_THREADS];
                    const float val = raw < thr ? raw : thr;
                    hp[k * DESC_THREADS] = val;
                    nrm2 = __fmaf_rn(val, val, nrm2);
                }
            }
        const float s = sqrtf(nrm2);
        const float f = 512.f / (s > 1.1920928955078125e-7f ? s : 1.1920928955078125e-7f);
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) {
                float* hp = s_h + (((i + 1) * (d + 2) + (j + 1)) * (n + 2)) * DESC_THREADS + tid;
                for (int k = 0; k < n; ++k) {
                    int v = __float2int_rn(hp[k * DESC_THREADS] * f);
                    v = v < 0 ? 0 : v > 255 ? 255 : v;
                    hp[k * DESC_THREADS] = (float)v;
                }
            }
    }
    __syncthreads();
    // coalesced write: thread-row t of the CTA = the 128 floats of keypoint s_row[t]
    for (int e = tid; e < DESC_THREADS * 128; e += DESC_THREADS) {
        const int t = e >> 7, k = e & 127;
        const int gi = s_row[t];
        if (gi < 0) continue;
        const int cell = k >> 3, kk = k & 7, i = cell >> 2, j = cell & 3;
        desc[(size_t)gi * 128 + k] = s_h[((((i + 1) * 6 + (j + 1)) * 10) + kk) * DESC_THREADS + t];
    }
}

int cv_round_d(double v) { return (int)lrint(v); }

// S.1 getGaussianKernel(n, sigma, CV_32F) -> getGaussianKernelBitExact (IEEE double arithmetic; exp() is the only libm call)
void gauss_taps(double sigma, int n, float* taps) {
    std::vector<double> v((size_t)n), r((size_t)n);
    const double scale2x = -0.125 / (sigma * sigma);
    const int n2 = (n - 1) / 2;
    double sum = 0;
    for (int i = 0, x = 1 - n; i < n2; ++i, x += 2) {
        v[i] = exp((double)(x * x) * scale2x);
        sum += v[i];
    }
    sum *= 2;
    sum += 1;
    const double mul1 = 1.0 / sum;
    double sum2 = 0;
    for (int i = 0; i < n2; ++i) {
        const double t = v[i] * mul1;
        r[i] = t;
        r[n - 1 - i] = t;
        sum2 += t;
    }
    sum2 *= 2;
    r[n2] = 1.0 * mul1;
    sum2 += r[n2];
    r[n2] += 1.0 - sum2;
    for (int i = 0; i < n; ++i) taps[i] = (float)r[i];
}

constexpr int kRadius[SIFT_GAUSS] = {5, 5, 6, 8, 10, 13};   // ksize = cvRound(sigma * 8 + 1) | 1 for the six layer sigmas at sigma 1.6

template <int R>
void configure_blur() {   // > 48 KB of dynamic shared memory needs the opt-in (per device: called from every extractor's constructor)
    SLIDEO_CUDA(cudaFuncSetAttribute(sift_blur_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, blur_smem_bytes<R>()));
}

template <int R>
void launch_blur_r(int layer, dim3 grid, cudaStream_t st, const float* src, size_t sis, int sp, float* dst, size_t dis, int dp, int W, int H) {
    sift_blur_kernel<R><<<grid, 256, blur_smem_bytes<R>(), st>>>(src, sis, sp, dst, dis, dp, W, H, layer);
}

void launch_blur(int layer, dim3 grid, cudaStream_t st, const float* src, size_t sis, int sp, float* dst, size_t dis, int dp, int W, int H) {
    switch (kRadius[layer]) {
        case 5: launch_blur_r<5>(layer, grid, st, src, sis, sp, dst, dis, dp, W, H); break;
        case 6: launch_blur_r<6>(layer, grid, st, src, sis, sp, dst, dis, dp, W, H); break;
        case 8: launch_blur_r<8>(layer, grid, st, src, sis, sp, dst, dis, dp, W, H); break;
        case 10: launch_blur_r<10>(layer, grid, st, src, sis, sp, dst, dis, dp, W, H); break;
        default: launch_blur_r<13>(layer, grid, st, src, sis, sp, dst, dis, dp, W, H); break;
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
SiftExtractor::SiftExtractor(int w, int h, int batch_cap, int kp_cap_per_image) : w_(w), h_(h), batch_cap_(batch_cap) {
    if (w < 8 || h < 8 || 2 * w > SIFT_MAX_DIM || 2 * h > SIFT_MAX_DIM) throw ArgError("image size out of range for SIFT (8..4095)");
    if (batch_cap < 1 || batch_cap > 256) throw ArgError("batch out of range (1..256)");
    // layer sigmas (S.5) and their taps (S.1)
    {
        double sig[SIFT_GAUSS];
        const float s = (float)1.6;
        const float dd = s * s - 0.5f * 0.5f * 4;
        sig[0] = (double)sqrtf(dd > 0.01f ? dd : 0.01f);
        const double k = pow(2., 1. / SIFT_LAYERS);
        for (int i = 1; i < SIFT_GAUSS; ++i) {
            const double sig_prev = pow(k, (double)(i - 1)) * 1.6, sig_total = sig_prev * k;
            sig[i] = sqrt(sig_total * sig_total - sig_prev * sig_prev);
        }
        float taps[SIFT_GAUSS][32];
        memset(taps, 0, sizeof taps);
        for (int i = 0; i < SIFT_GAUSS; ++i) {
            const int n = cv_round_d(sig[i] * 4 * 2 + 1) | 1;
            if (n != 2 * kRadius[i] + 1) throw NotImplError("unexpected Gaussian kernel size");
            gauss_taps(sig[i], n, taps[i]);
        }
        SLIDEO_CUDA(cudaMemcpyToSymbol(c_taps, taps, sizeof taps));
        float tab[64];
        for (int j = 0; j < 64; ++j) tab[j] = (float)(exp2((double)j / 64) * .9670371139572337719125840413672004409288e-2);
        SLIDEO_CUDA(cudaMemcpyToSymbol(c_exptab, tab, sizeof tab));
    }
    SiftGeo& g = geo_;
    memset(&g, 0, sizeof g);
    const int m = 2 * (w < h ? w : h);
    g.n_oct = cv_round_d(log((double)m) / log(2.) - 2) + 1;
    if (g.n_oct < 1 || g.n_oct > SIFT_MAX_OCT) throw ArgError("unsupported octave count");
    int W = 2 * w, H = 2 * h, tile_base = 0;
    size_t off = 0;
    for (int o = 0; o < g.n_oct; ++o) {
        SiftOctave& oc = g.oc[o];
        if (W < 1 || H < 1) { g.n_oct = o; break; }
        oc.w = W; oc.h = H;
        oc.pitch = (int)align_up((size_t)W, 4);
        oc.diag = (int)sqrt((double)W * W + (double)H * H);
        oc.tiles_x = cdiv(W, EX_TX); oc.tiles_y = cdiv(H, EX_TY);
        oc.tile_base = tile_base;
        tile_base += oc.tiles_x * oc.tiles_y;
        oc.layer_stride = align_up((size_t)oc.pitch * H, 64);
        oc.off = off;
        off += oc.layer_stride * SIFT_GAUSS;
        if (o > 0) {
            oc.ifx = 1. / ((double)W / g.oc[o - 1].w);
            oc.ify = 1. / ((double)H / g.oc[o - 1].h);
        }
        W /= 2; H /= 2;
    }
    g.total_tiles = tile_base;
    g.img_floats = off;
    const long long px = (long long)w * h;
    g.kp_cap = kp_cap_per_image > 0 ? kp_cap_per_image : (px >= 1000000 ? 65536 : px >= 200000 ? 32768 : 16384);
    g.cand_cap = g.kp_cap * 8;
    total_cap_ = (size_t)batch_cap * g.kp_cap;

    const size_t B = (size_t)batch_cap;
    SLIDEO_CUDA(cudaMalloc(&d_pyr_, B * g.img_floats * 4));
    SLIDEO_CUDA(cudaMalloc(&d_gray_, B * (size_t)w * h));
    SLIDEO_CUDA(cudaMalloc(&d_cand_, B * g.cand_cap * 4));
    if (batch_cap * (long long)g.kp_cap >= (1ll << 31) || g.kp_cap > 65536) throw ArgError("SIFT keypoint capacity out of range");
    SLIDEO_CUDA(cudaMalloc(&d_cnt_, (4 * B + 128) * 4));
    SLIDEO_CUDA(cudaMalloc(&d_refined_, B * g.kp_cap * 6 * 4));
    SLIDEO_CUDA(cudaMalloc(&d_perm_, total_cap_ * 4));
    SLIDEO_CUDA(cudaMalloc(&d_raw_, B * g.kp_cap * 6 * 4));
    SLIDEO_CUDA(cudaMalloc(&d_order_, B * g.kp_cap * 4));
    SLIDEO_CUDA(cudaMalloc(&d_uniq_, B * g.kp_cap * 6 * 4));
    SLIDEO_CUDA(cudaMalloc(&d_kp_f_, total_cap_ * 5 * 4));
    SLIDEO_CUDA(cudaMalloc(&d_kp_oct_, total_cap_ * 4));
    SLIDEO_CUDA(cudaMalloc(&d_q_frame_, total_cap_ * 4));
    SLIDEO_CUDA(cudaMalloc(&d_desc_, total_cap_ * 128 * 4));
    SLIDEO_CUDA(cudaMalloc(&d_frame_off_, (B + 1) * 4));
    SLIDEO_CUDA(cudaMalloc(&d_frame_nkp_, B * 4));
    SLIDEO_CUDA(cudaMallocHost(&h_pinned_, (4 * B + 4) * 4));
    SLIDEO_CUDA(cudaFuncSetAttribute(sift_descriptor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DESC_HIST * DESC_THREADS * 4));
    configure_blur<5>(); configure_blur<6>(); configure_blur<8>(); configure_blur<10>(); configure_blur<13>();
}

SiftExtractor::~SiftExtractor() {
    cudaFree(d_pyr_); cudaFree(d_gray_); cudaFree(d_cand_); cudaFree(d_cnt_); cudaFree(d_raw_); cudaFree(d_order_); cudaFree(d_uniq_);
    cudaFree(d_refined_); cudaFree(d_perm_);
    cudaFree(d_kp_f_); cudaFree(d_kp_oct_); cudaFree(d_q_frame_); cudaFree(d_desc_); cudaFree(d_frame_off_); cudaFree(d_frame_nkp_);
    cudaFreeHost(h_pinned_);
}

int SiftExtractor::run(const uint8_t* d_src, int n, int stride, size_t frame_stride, int channels, cudaStream_t stream, int* launches) {
    if (n < 1 || n > batch_cap_) throw ArgError("batch size out of range");
    if (channels != 1 && channels != 3) throw ArgError("channels must be 1 or 3");
    const SiftGeo& g = geo_;
    const size_t B = (size_t)batch_cap_;
    int nl = 0;
    int32_t *d_cand_cnt = d_cnt_, *d_raw_cnt = d_cnt_ + B, *d_uniq_cnt = d_cnt_ + 2 * B, *d_ref_cnt = d_cnt_ + 3 * B;
    int32_t *d_bucket_cnt = d_cnt_ + 4 * B, *d_bucket_cursor = d_cnt_ + 4 * B + 64;
    SLIDEO_CUDA(cudaMemsetAsync(d_cnt_, 0, (4 * B + 128) * 4, stream));

    const uint8_t* gray = d_src;
    int gpitch = stride;
    size_t gstride = frame_stride;
    if (channels == 3) {
        sift_gray_kernel<<<dim3(cdiv(w_, 256), h_, n), 256, 0, stream>>>(d_src, stride, frame_stride, d_gray_, w_, h_);
        ++nl;
        gray = d_gray_;
        gpitch = w_;
        gstride = (size_t)w_ * h_;
    }
    // S.5 buildGaussianPyramid
    for (int o = 0; o < g.n_oct; ++o) {
        const SiftOctave& oc = g.oc[o];
        float* L0 = d_pyr_ + oc.off;
        const dim3 grid(cdiv(oc.w, TW), cdiv(oc.h, TH), n);
        if (o == 0) {   // the 2x image is parked in the slot of layer 5 (written for real only after layers 0..4 exist)
            float* up = L0 + (size_t)(SIFT_GAUSS - 1) * oc.layer_stride;
            sift_up2_kernel<<<dim3(cdiv(w_, 256), h_, n), 256, 0, stream>>>(gray, gstride, gpitch, w_, h_, up, g.img_floats, oc.pitch);
            ++nl;
            launch_blur(0, grid, stream, up, g.img_floats, oc.pitch, L0, g.img_floats, oc.pitch, oc.w, oc.h);
        } else {
            const SiftOctave& po = g.oc[o - 1];
            sift_half_kernel<<<dim3(cdiv(oc.w, 256), oc.h, n), 256, 0, stream>>>(d_pyr_ + po.off + SIFT_LAYERS * po.layer_stride, g.img_floats, po.pitch,
                                                                                 po.w, po.h, L0, oc.pitch, oc.w, oc.h, oc.ifx, oc.ify);
        }
        ++nl;
        for (int i = 1; i < SIFT_GAUSS; ++i) {
            launch_blur(i, grid, stream, L0 + (size_t)(i - 1) * oc.layer_stride, g.img_floats, oc.pitch, L0 + (size_t)i * oc.layer_stride, g.img_floats,
                        oc.pitch, oc.w, oc.h);
            ++nl;
        }
    }
    const float threshold = (float)(int)floor(0.5 * 0.04 / SIFT_LAYERS * 255);
    sift_extrema_kernel<<<dim3(g.total_tiles, n), 256, 0, stream>>>(d_pyr_, g, d_cand_, d_cand_cnt, threshold);
    ++nl;
    sift_refine_kernel<<<dim3(cdiv(g.cand_cap, 128), n), 128, 0, stream>>>(d_pyr_, g, d_cand_, d_cand_cnt, static_cast<RefinedCand*>(d_refined_), d_ref_cnt);
    ++nl;
    sift_orient_kernel<<<dim3(cdiv(g.kp_cap, 128), n), 128, 0, stream>>>(d_pyr_, g, static_cast<const RefinedCand*>(d_refined_), d_ref_cnt, d_raw_, d_raw_cnt);
    ++nl;
    sift_rank_kernel<<<dim3(cdiv(g.kp_cap, 256), n), 256, 0, stream>>>(d_raw_, d_raw_cnt, g.kp_cap, d_order_);
    ++nl;
    sift_unique_kernel<<<n, 1024, 0, stream>>>(d_raw_, d_raw_cnt, g.kp_cap, d_order_, d_uniq_, d_uniq_cnt);
    ++nl;
    sift_bucket_count_kernel<<<dim3(cdiv(g.kp_cap, 256), n), 256, 0, stream>>>(d_uniq_, d_uniq_cnt, g, d_bucket_cnt);
    sift_bucket_scan_kernel<<<1, 32, 0, stream>>>(d_bucket_cnt, d_bucket_cursor);
    sift_bucket_scatter_kernel<<<dim3(cdiv(g.kp_cap, 256), n), 256, 0, stream>>>(d_uniq_, d_uniq_cnt, g, d_bucket_cursor, d_perm_);
    nl += 3;
    SLIDEO_CUDA(cudaGetLastError());
    SLIDEO_CUDA(cudaMemcpyAsync(h_pinned_, d_cnt_, 4 * B * 4, cudaMemcpyDeviceToHost, stream));
    SLIDEO_CUDA(cudaStreamSynchronize(stream));   // keypoint counts size the descriptor launch
    h_frame_off_.assign(1, 0);
    last_cand_ = 0;
    for (int i = 0; i < n; ++i) {
        if (h_pinned_[i] > g.cand_cap) throw CapacityError("SIFT extrema candidate capacity exceeded on at least one image");
        if (h_pinned_[B + i] > g.kp_cap || h_pinned_[3 * B + i] > g.kp_cap) throw CapacityError("SIFT keypoint capacity exceeded on at least one image");
        last_cand_ += h_pinned_[i];
        h_frame_off_.push_back(h_frame_off_.back() + h_pinned_[2 * B + i]);
    }
    const int total = h_frame_off_.back();
    SLIDEO_CUDA(cudaMemcpyAsync(d_frame_off_, h_frame_off_.data(), ((size_t)n + 1) * 4, cudaMemcpyHostToDevice, stream));
    SLIDEO_CUDA(cudaMemcpyAsync(d_frame_nkp_, d_uniq_cnt, (size_t)n * 4, cudaMemcpyDeviceToDevice, stream));
    if (total > 0) {
        sift_descriptor_kernel<<<cdiv(total, DESC_THREADS), DESC_THREADS, DESC_HIST * DESC_THREADS * 4, stream>>>(
            d_pyr_, g, d_uniq_, d_frame_off_, d_perm_, total, d_kp_f_, d_kp_oct_, d_q_frame_, d_desc_);
        ++nl;
        SLIDEO_CUDA(cudaGetLastError());
    }
    if (launches) *launches += nl;
    return total;
}

}  // namespace slideo
