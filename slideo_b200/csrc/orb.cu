// orb.cu -- K1-K7: batched ORB feature extraction, bit-exact with OpenCV's ORB as the reference configures it
// (crates/matching-opencv/src/feature_extractor.rs:12-46: ORB::create(2000, 1.2, 8, 62, 0, 2, FAST_SCORE, 62, 20)
//  .detectAndCompute).  The algorithm is the one restated in oracle/orb_oracle.c; every
// kernel below names the stage it implements.  All images of a batch share one geometry.
//
//   K1 gray_kernel      BGR -> gray (fixed point), level 0 of the pyramid
//   K2 resize_kernel    INTER_LINEAR_EXACT 8.8 fixed-point chain, level l from level l-1
//   K3 fast_kernel      FAST-9/16 score + 3x3 NMS + border(edgeThreshold) filter, all levels in one launch
//   K4 select_kernel    retainBest(quota) with ties via a 256-bin score histogram, canonical (y, x) sort
//      scan_kernel      keypoint offsets per (image, level); scatter_kernel writes keypoint records
//   K6 blur_kernel      7x7 sigma-2 Gaussian, separable fp32 with OpenCV's exact fma order, all levels in one launch
//   K5+K7 describe_kernel  intensity-centroid angle (fastAtan2) + 256-bit steered descriptor, one warp per keypoint
#include "orb.cuh"
#include "tma.cuh"

#include <math.h>
#include <string.h>

#include <algorithm>

namespace slideo {

namespace {

constexpr int TILE_W = 64, TILE_H = 64;
constexpr unsigned FULL = 0xFFFFFFFFu;

struct Geo {
    int nlevels, total_tiles, edge, fast_thr, patch, half;
    size_t pyr_img_bytes, cand_img_words, sel_img_words;
    int umax[40];
    OrbLevelGeom lv[ORB_MAX_LEVELS];
};

__device__ __forceinline__ int reflect101(int p, int n) {
    // BORDER_REFLECT_101; n >= 2 in practice, loop handles far-out coordinates
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
    return p;
}

// ---- K1 -----------------------------------------------------------------------------------------------------
// cvtColor BGR2GRAY 8u: (B*3735 + G*19235 + R*9798 + 2^14) >> 15   
__device__ __forceinline__ uint32_t gray_of(uint32_t B, uint32_t G, uint32_t R) { return (B * 3735u + G * 19235u + R * 9798u + 16384u) >> 15; }

// A thread converts 16 pixels: three 16-byte loads in flight, one 16-byte store (the kernel is bound by memory latency, not by the
// arithmetic: with 4 pixels per thread 60 % of its stall samples sat on the first use of the loaded word).
__global__ void __launch_bounds__(128) gray_kernel(const uint8_t* __restrict__ src, int stride, size_t frame_stride,
                                                   int channels, uint8_t* __restrict__ pyr, size_t pyr_img_bytes, int w,
                                                   int h, int pitch) {
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    const int y = blockIdx.y, img = blockIdx.z;
    if (x >= w) return;
    const uint8_t* row = src + (size_t)img * frame_stride + (size_t)y * stride;
    uint8_t* dst = pyr + (size_t)img * pyr_img_bytes + (size_t)y * pitch + x;   // pitch and x are multiples of 16
    if (channels == 3) {
        const uint8_t* p = row + 3 * x;
        if (x + 16 <= w && (((uintptr_t)p) & 15) == 0) {
            const uint4* p128 = reinterpret_cast<const uint4*>(p);
            const uint4 v0 = __ldg(p128), v1 = __ldg(p128 + 1), v2 = __ldg(p128 + 2);
            const uint32_t wds[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
            uint32_t out[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {   // 4 pixels = 3 words: B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
                const uint32_t a = wds[3 * q], bb = wds[3 * q + 1], c = wds[3 * q + 2];
                const uint32_t g0 = gray_of(a & 255, (a >> 8) & 255, (a >> 16) & 255);
                const uint32_t g1 = gray_of(a >> 24, bb & 255, (bb >> 8) & 255);
                const uint32_t g2 = gray_of((bb >> 16) & 255, bb >> 24, c & 255);
                const uint32_t g3 = gray_of((c >> 8) & 255, (c >> 16) & 255, c >> 24);
                out[q] = g0 | g1 << 8 | g2 << 16 | g3 << 24;
            }
            *reinterpret_cast<uint4*>(dst) = make_uint4(out[0], out[1], out[2], out[3]);
        } else {
            for (int i = 0; i < 16 && x + i < w; ++i) dst[i] = (uint8_t)gray_of(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
        }
    } else if (x + 16 <= w && (((uintptr_t)(row + x)) & 15) == 0) {
        *reinterpret_cast<uint4*>(dst) = __ldg(reinterpret_cast<const uint4*>(row + x));
    } else {
        for (int i = 0; i < 16 && x + i < w; ++i) dst[i] = row[x + i];
    }
}

// ---- K2 -----------------------------------------------------------------------------------------------------
// resize INTER_LINEAR_EXACT, 8-bit: horizontal 8.8 fixed point then vertical, (v + 2^15) >> 16   
// tables (host-built in double precision, identical to the oracle): xofs[dw], xc1[dw], yofs[dh], yc1[dh]
// A thread owns 4 destination columns and walks down a strip of RS_H destination rows: its x tables stay in registers, and
// the horizontally interpolated source row shared by two consecutive destination rows (scale 1.2: the lower source row of row y
// is usually the upper source row of row y + 1) is computed once.
constexpr int RS_H = 30;

__global__ void __launch_bounds__(128) resize_kernel(uint8_t* __restrict__ pyr, size_t pyr_img_bytes, size_t src_off,
                                                     int sw, int sh, int spitch, size_t dst_off, int dw, int dh,
                                                     int dpitch, const int32_t* __restrict__ tab) {
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x0 >= dw) return;
    const int y0 = blockIdx.y * RS_H, y1 = min(y0 + RS_H, dh), img = blockIdx.z;
    const int32_t* xofs = tab;
    const int32_t* xc1 = tab + dw;
    const int32_t* yofs = tab + 2 * dw;
    const int32_t* yc1 = tab + 2 * dw + dh;
    int ox[4], ox1[4];
    uint32_t c0[4], c1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = min(x0 + i, dw - 1);
        ox[i] = __ldg(xofs + x);
        c1[i] = (uint32_t)__ldg(xc1 + x);
        c0[i] = 256u - c1[i];
        ox1[i] = min(ox[i] + 1, sw - 1);
    }
    const uint8_t* base = pyr + (size_t)img * pyr_img_bytes + src_off;
    uint8_t* dbase = pyr + (size_t)img * pyr_img_bytes + dst_off;
    auto hrow = [&](int r, uint32_t (&o)[4]) {   // horizontal 8.8 fixed-point pass of source row r at this thread's 4 columns
        const uint8_t* sr = base + (size_t)r * spitch;
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = (uint32_t)sr[ox[i]] * c0[i] + (uint32_t)sr[ox1[i]] * c1[i];
    };
    int prev_row = -1;
    uint32_t prev[4] = {0, 0, 0, 0};
    for (int y = y0; y < y1; ++y) {
        const int oy = __ldg(yofs + y), oy1 = min(oy + 1, sh - 1);
        const uint32_t fy = (uint32_t)__ldg(yc1 + y);
        uint32_t a[4], b[4];
        if (oy == prev_row) {
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = prev[i];
        } else {
            hrow(oy, a);
        }
        if (oy1 == oy) {
#pragma unroll
            for (int i = 0; i < 4; ++i) b[i] = a[i];
        } else {
            hrow(oy1, b);
        }
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            prev[i] = b[i];
            const uint32_t v = (a[i] * (256u - fy) + b[i] * fy + (1u << 15)) >> 16;
            if (x0 + i < dw) o |= min(v, 255u) << (8 * i);
        }
        prev_row = oy1;
        *reinterpret_cast<uint32_t*>(dbase + (size_t)y * dpitch + x0) = o;
    }
}

// ---- K3 -----------------------------------------------------------------------------------------------------
// FAST-9/16 (threshold t, nonmaxSuppression=true) restated: score = max over the 16 arcs of 9 contiguous circle
// pixels of max(min(c - p), min(p - c)); corner iff score > t; response = score - 1; NMS strict > over 8 neighbours
// (SURVEY A.4).  Then KeyPointsFilter::runByImageBorder(edgeThreshold) (A.5).
__device__ __forceinline__ int fast_score(const uint8_t* sp, int pitch) {
    // sp points at the centre pixel inside the shared tile.  Both polarities run in one pass on packed 16-bit lanes
    // (VIMNMX3.S16x2): lane 0 holds c - p ("darker" margin), lane 1 holds p - c ("brighter" margin); the min over a 9-arc is
    // min3 of three min3's over 3 consecutive circle pixels, the score is the max over the 16 arcs and both lanes, floored at 0.
    const int c = sp[0];
    const int off[16] = {3 * pitch,      3 * pitch + 1,  2 * pitch + 2,  pitch + 3,  3,  -pitch + 3,  -2 * pitch + 2, -3 * pitch + 1,
                         -3 * pitch,     -3 * pitch - 1, -2 * pitch - 2, -pitch - 3, -3, pitch - 3,   2 * pitch - 2,  3 * pitch - 1};
    uint32_t de[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int p = sp[off[k]];
        de[k] = __byte_perm((uint32_t)(c - p), (uint32_t)(p - c), 0x5410);
    }
    uint32_t a3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a3[k] = __vimin3_s16x2(de[k], de[(k + 1) & 15], de[(k + 2) & 15]);
    uint32_t best = 0;   // (0, 0): scores are floored at 0 like the scalar max(best = 0, ...)
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const uint32_t a9 = __vimin3_s16x2(a3[k], a3[(k + 3) & 15], a3[(k + 6) & 15]);
        const uint32_t b9 = __vimin3_s16x2(a3[k + 1], a3[(k + 4) & 15], a3[(k + 7) & 15]);
        best = __vimax3_s16x2(best, a9, b9);
    }
    const int lo = (int)(short)(best & 0xFFFFu), hi = (int)best >> 16;
    return max(lo, hi);
}

// byte flags (0x80 per byte) of the bytes of x that are > thr, for thr < 127:  k127 = (127 - thr) * 0x01010101
__device__ __forceinline__ uint32_t bytes_gt(uint32_t x, uint32_t k127) {
    return (((x & 0x7F7F7F7Fu) + k127) | x) & 0x80808080u;
}

struct FastMaps { CUtensorMap lv[ORB_MAX_LEVELS]; };   // one rank-3 u8 tensor {x, y, image} per pyramid level

__global__ void __launch_bounds__(256) fast_kernel(const FastMaps* __restrict__ maps, const Geo* __restrict__ gp,
                                                   const uint32_t* __restrict__ tiles, uint32_t* __restrict__ cand,
                                                   int32_t* __restrict__ cand_cnt) {
    // pixel tile: 72 rows x 24 words (columns tx0-16 .. tx0+79); score tile: 66 rows x 72 columns = tx0-4 .. tx0+67, i.e. the
    // 64 x 64 pixels of the tile plus the halo the 3x3 NMS needs, rounded to whole words so every thread handles 4 pixels
    // (64 x 32 tiles spent a quarter of the kernel's instructions on the per-CTA prologue: 64 x 64 is 1 % of the step faster)
    constexpr int PH = TILE_H + 8, PWW = (TILE_W + 32) / 4, PWB = PWW * 4, PX0 = 2;   // the staged rows start 16 B-aligned at tx0 - 16 (TMA): PX0 words before tx0 - 8
    constexpr int SH = TILE_H + 2, SWW = (TILE_W + 8) / 4, SWB = SWW * 4;
    __shared__ __align__(128) uint32_t s_px[PH][PWW];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ __align__(16) uint8_t s_sc[SH][SWB + 8];
    __shared__ uint16_t s_list[SH * SWB];              // (row << 7 | col) of the positions that pass the cheap necessary test
    __shared__ int s_n;

    const Geo& g = *gp;
    const uint32_t tref = __ldg(tiles + blockIdx.x);   // level | tx0 << 4 | ty0 << 16, precomputed on the host: no per-CTA search / division
    const int l = tref & 15, tx0 = (tref >> 4) & 0xFFF, ty0 = tref >> 16;
    const OrbLevelGeom& L = g.lv[l];
    const int img = blockIdx.y;
    // keypoints survive only edge pixels inside the level (runByImageBorder): tiles that lie entirely in that band do nothing,
    // and inside a tile scores are only needed for rows / columns [edge - 1, size - edge] (the 3x3 NMS looks one pixel out)
    const int eb = g.edge;
    if (tx0 + TILE_W <= eb || tx0 >= L.w - eb || ty0 + TILE_H <= eb || ty0 >= L.h - eb) return;

    // tile load by TMA: one 96 x 72 byte box of the level's {x, y, image} tensor (columns tx0-16 .., rows ty0-4 ..) lands in s_px
    // while the threads clear the score tile; elements outside the level (negative coordinates included) arrive as zero,
    // which is what FAST needs there (it is only evaluated 3 pixels inside the level)
    if (threadIdx.x == 0) {
        s_n = 0;
        tma_mbar_init(&s_bar, 1);
        tma_mbar_expect_tx(&s_bar, PH * PWB);
        tma_load_3d(&s_px[0][0], &maps->lv[l], tx0 - 16, ty0 - 4, img, &s_bar);
    }
    for (int i = threadIdx.x; i < SH * (SWB + 8) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(&s_sc[0][0])[i] = 0;
    __syncthreads();
    tma_mbar_wait(&s_bar, 0);

    // pass 1, 4 pixels per thread: a corner needs two ADJACENT compass points (N,E,S,W at distance 3) on the same side of the
    // centre by more than thr (any 9-arc contains two adjacent compass points); the test below drops the "same side"
    // part -- (|N-c| > thr or |S-c| > thr) and (|E-c| > thr or |W-c| > thr) -- so it is a superset, evaluated with
    // VABSDIFF4 on packed bytes.  Survivors are compacted into s_list so that the exact score runs on dense warps.
    const int thr = g.fast_thr;
    const uint32_t k127 = (uint32_t)(127 - thr) * 0x01010101u;
    const bool interior = tx0 - 4 >= 3 && tx0 + TILE_W + 4 <= L.w - 3 && ty0 - 1 >= 3 && ty0 + TILE_H + 1 <= L.h - 3;
    for (int i = threadIdx.x; i < SH * SWW; i += blockDim.x) {
        const int r = i / SWW, wq = i - r * SWW + 1;      // score row r <-> pixel row r + 3; score word wq - 1 <-> pixel word wq
        {
            const int gy = ty0 - 1 + r, gx0 = tx0 - 8 + wq * 4;
            if (gy < eb - 1 || gy > L.h - eb || gx0 + 3 < eb - 1 || gx0 > L.w - eb) continue;   // outside the band that can matter
        }
        const int pw = wq + PX0;
        const uint32_t C = s_px[r + 3][pw], Wm = s_px[r + 3][pw - 1], Wp = s_px[r + 3][pw + 1], U = s_px[r][pw], D = s_px[r + 6][pw];
        const uint32_t Lw = __byte_perm(Wm, C, 0x4321), Rw = __byte_perm(C, Wp, 0x6543);
        uint32_t f = (bytes_gt(__vabsdiffu4(U, C), k127) | bytes_gt(__vabsdiffu4(D, C), k127)) &
                     (bytes_gt(__vabsdiffu4(Lw, C), k127) | bytes_gt(__vabsdiffu4(Rw, C), k127));
        if (f != 0 && !interior) {   // FAST is only defined 3 pixels inside the level
            const int gy = ty0 - 1 + r;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int gx = tx0 - 8 + wq * 4 + b;
                if (!(gx >= 3 && gx < L.w - 3 && gy >= 3 && gy < L.h - 3)) f &= ~(0x80u << (8 * b));
            }
        }
        if (f != 0) {
            int pos = atomicAdd(&s_n, __popc(f));
            const int c0 = (wq - 1) * 4;
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (f & (0x80u << (8 * b))) s_list[pos++] = (uint16_t)((r << 7) | (c0 + b));
        }
    }
    __syncthreads();

    // pass 2: exact FAST score on the survivors only
    const int n_list = s_n;
    for (int j = threadIdx.x; j < n_list; j += blockDim.x) {
        const int code = s_list[j], r = code >> 7, c = code & 127;
        const uint8_t* sp = reinterpret_cast<const uint8_t*>(&s_px[r + 3][0]) + (c + 4 + 4 * PX0);
        const int sc = fast_score(sp, PWB);
        if (sc > thr) s_sc[r][c] = (uint8_t)(sc - 1);
    }
    __syncthreads();

    // pass 3: 3x3 non-maximum suppression + border filter, again only where a score can be non-zero
    const int e = g.edge;
    for (int j = threadIdx.x; j < n_list; j += blockDim.x) {
        const int code = s_list[j], r = code >> 7, c = code & 127;
        if (r < 1 || r > TILE_H || c < 4 || c >= TILE_W + 4) continue;   // halo positions belong to the neighbouring tiles
        const int sc = s_sc[r][c];
        if (sc == 0) continue;
        const int gx = tx0 - 4 + c, gy = ty0 - 1 + r;
        if (gx < e || gx >= L.w - e || gy < e || gy >= L.h - e) continue;
        if (sc > s_sc[r - 1][c - 1] && sc > s_sc[r - 1][c] && sc > s_sc[r - 1][c + 1] && sc > s_sc[r][c - 1] && sc > s_sc[r][c + 1] &&
            sc > s_sc[r + 1][c - 1] && sc > s_sc[r + 1][c] && sc > s_sc[r + 1][c + 1]) {
            const int pos = atomicAdd(&cand_cnt[img * g.nlevels + l], 1);
            if (pos < L.cand_cap)
                cand[(size_t)img * g.cand_img_words + L.cand_off + pos] = ((uint32_t)sc << 24) | ((uint32_t)gy << 12) | (uint32_t)gx;
        }
    }
}

// ---- K4 -----------------------------------------------------------------------------------------------------
// KeyPointsFilter::retainBest(quota): keep everything >= the quota-th largest response (ties kept),
// then canonical (y, x) order (A.9).  One CTA per (level, image).
__global__ void __launch_bounds__(256) select_kernel(const Geo* __restrict__ gp, const uint32_t* __restrict__ cand,
                                                     const int32_t* __restrict__ cand_cnt, uint32_t* __restrict__ sel,
                                                     int32_t* __restrict__ sel_cnt, int32_t* __restrict__ flags) {
    extern __shared__ uint32_t s_keys[];  // sel_cap entries
    __shared__ int s_hist[256];
    __shared__ int s_thr, s_n;
    const Geo& g = *gp;
    const int l = blockIdx.x, img = blockIdx.y;
    const OrbLevelGeom L = g.lv[l];
    int n = cand_cnt[img * g.nlevels + l];
    if (n > L.cand_cap) {
        if (threadIdx.x == 0) atomicOr(flags, 1);
        n = L.cand_cap;
    }
    const uint32_t* c = cand + (size_t)img * g.cand_img_words + L.cand_off;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&s_hist[c[i] >> 24], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int thr = 0;
        if (n > L.quota) {
            thr = 256;
            int acc = 0;
            for (int t = 255; t > 0 && L.quota > 0; --t) {
                acc += s_hist[t];
                if (acc >= L.quota) { thr = t; break; }
            }
        }
        s_thr = thr;
    }
    __syncthreads();
    const int thr = s_thr;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t v = c[i];
        if ((int)(v >> 24) >= thr) {
            const int pos = atomicAdd(&s_n, 1);
            if (pos < L.sel_cap) s_keys[pos] = v;
        }
    }
    __syncthreads();
    int m = s_n;
    if (m > L.sel_cap) {
        if (threadIdx.x == 0) atomicOr(flags, 2);
        m = L.sel_cap;
    }
    int p2 = 1;
    while (p2 < m) p2 <<= 1;
    for (int i = m + threadIdx.x; i < p2; i += blockDim.x) s_keys[i] = 0xFFFFFFFFu;
    __syncthreads();
    // bitonic sort on (y << 12 | x) = low 24 bits (padding sorts last)
    for (int size = 2; size <= p2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < p2; i += blockDim.x) {
                const int j = i ^ stride;
                if (j > i) {
                    const uint32_t a = s_keys[i], b = s_keys[j];
                    const bool pa = a == 0xFFFFFFFFu, pb = b == 0xFFFFFFFFu;
                    const uint32_t ka = pa ? 0xFFFFFFFFu : (a & 0xFFFFFFu), kb = pb ? 0xFFFFFFFFu : (b & 0xFFFFFFu);
                    const bool up = (i & size) == 0;
                    if ((ka > kb) == up) { s_keys[i] = b; s_keys[j] = a; }
                }
            }
            __syncthreads();
        }
    }
    uint32_t* out = sel + (size_t)img * g.sel_img_words + L.sel_off;
    for (int i = threadIdx.x; i < m; i += blockDim.x) out[i] = s_keys[i];
    if (threadIdx.x == 0) sel_cnt[img * g.nlevels + l] = m;
}

// exclusive scan of sel_cnt over (image, level) -> kp_off[n*nlevels + 1] (batch-local), frame_off[n + 1], frame_nkp[n], and the
// batch header info[] = {total keypoints, first query index of the batch in the stream, first frame index, flags}.
// Stream mode (st != nullptr): the batch is appended to a query stream whose counters live on the device (KnnStream), so the
// host never needs the counts: frame_q0 / frame_nkp of the stream are extended here and the counters advanced.
__global__ void __launch_bounds__(256) scan_kernel(const int32_t* __restrict__ sel_cnt, int n_img, int nlevels,
                                                   int32_t* __restrict__ kp_off, int32_t* __restrict__ frame_off,
                                                   int32_t* __restrict__ frame_nkp, const int32_t* __restrict__ flags,
                                                   int32_t* __restrict__ info, int32_t* __restrict__ h_out, KnnStream* st,
                                                   int32_t* __restrict__ s_frame_q0, int32_t* __restrict__ s_frame_nkp, int q_cap,
                                                   int f_cap, int kp_cap) {
    __shared__ int s_tot[1024];
    const int tid = threadIdx.x;
    for (int f = tid; f < n_img; f += blockDim.x) {
        int t = 0;
        for (int l = 0; l < nlevels; ++l) t += sel_cnt[f * nlevels + l];
        s_tot[f] = t;
        frame_nkp[f] = t;
    }
    __syncthreads();
    if (tid == 0) {
        const int base_q = st ? st->q_write : 0, base_f = st ? st->f_write : 0;
        int acc = 0;
        for (int f = 0; f < n_img; ++f) {
            frame_off[f] = acc;
            if (h_out) h_out[2 + f] = acc;
            int a2 = acc;
            for (int l = 0; l < nlevels; ++l) { kp_off[f * nlevels + l] = a2; a2 += sel_cnt[f * nlevels + l]; }
            acc += s_tot[f];
        }
        frame_off[n_img] = acc;
        kp_off[n_img * nlevels] = acc;
        int fl = flags[0];
        if (acc > kp_cap) fl |= 4;
        if (st) {
            if ((long long)base_q + acc > q_cap) fl |= 4;
            if (base_f + n_img > f_cap) fl |= 8;
        }
        if (fl & 12) acc = 0;                          // never read / write past a buffer: the batch is dropped and the flag fails the call
        if (st) {
            if (!(fl & 8)) {
                int a3 = base_q;
                for (int f = 0; f < n_img; ++f) {
                    const int t = (fl & 4) ? 0 : s_tot[f];
                    a3 += t;
                    s_frame_q0[base_f + f + 1] = a3;
                    s_frame_nkp[base_f + f] = (fl & 4) ? -1 : t;
                }
                st->f_write = base_f + n_img;
            }
            st->q_write = base_q + acc;
            st->flags |= fl;
        }
        info[0] = acc;
        info[1] = base_q;
        info[2] = base_f;
        info[3] = fl;
        if (h_out) {
            h_out[2 + n_img] = acc;
            h_out[0] = acc;
            h_out[1] = fl;
        }
    }
}

// keypoint records in canonical order: kp_src[g] = img << 4 | level ... packed with the selected entry
__global__ void __launch_bounds__(256) scatter_kernel(const Geo* __restrict__ gp, const uint32_t* __restrict__ sel,
                                                      const int32_t* __restrict__ sel_cnt,
                                                      const int32_t* __restrict__ kp_off, uint32_t* __restrict__ kp_src,
                                                      int32_t* __restrict__ q_frame, int32_t* __restrict__ kp_i,
                                                      size_t kp_cap, const int32_t* __restrict__ info, int32_t* __restrict__ s_q_frame) {
    const Geo& g = *gp;
    const int l = blockIdx.x, img = blockIdx.y;
    const OrbLevelGeom L = g.lv[l];
    if (info[0] == 0) return;                          // empty batch, or a batch dropped by a capacity flag
    const int base_q = info[1], base_f = info[2];
    const int m = sel_cnt[img * g.nlevels + l], off = kp_off[img * g.nlevels + l];
    const uint32_t* s = sel + (size_t)img * g.sel_img_words + L.sel_off;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const size_t o = (size_t)off + i;
        if (o >= kp_cap) continue;
        const uint32_t v = s[i];
        kp_src[2 * o] = v;
        kp_src[2 * o + 1] = ((uint32_t)img << 8) | (uint32_t)l;
        q_frame[o] = img;
        if (s_q_frame) s_q_frame[(size_t)base_q + o] = base_f + img;
        reinterpret_cast<int4*>(kp_i)[o] = make_int4((int)(v & 0xFFF), (int)((v >> 12) & 0xFFF), l, (int)(v >> 24));
    }
}

// ---- K6 -----------------------------------------------------------------------------------------------------
// GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) as ORB invokes it: the generic float separable filter
// (row: acc = k0*p0; acc = fma(k_i, p_i, acc) left to right; column: acc = k3*r3; acc = fma(k_{3+j}, r_{3+j} + r_{3-j},
// acc); round-half-even, saturate) -- SURVEY A.7.
// Tile = BT_W x BT_H pixels, one warp per strip of BS_H rows, one lane per 4 columns.  The whole box (tile + halo, rounded to the
// 16-byte alignment TMA needs in x) arrives by ONE tensor copy; elements outside the level arrive as zeros and the (at most 3)
// halo rows / columns that BORDER_REFLECT_101 defines are then written in place from their mirror positions inside the box.
// Every thread walks down its strip: the row pass of a new input row stays in registers (4 floats), the column pass reads
// the last 7 of them -- no shared-memory round trip for the intermediate, every input row converted and filtered once per strip.
constexpr int BT_W = 128, BT_H = 64, BS_H = 16, BT_THREADS = (BT_W / 4) * (BT_H / BS_H);
constexpr int BB_W = BT_W + 32, BB_H = BT_H + 8;   // box: columns tx0 - 16 .. tx0 + 143, rows ty0 - 4 .. ty0 + 67

__global__ void __launch_bounds__(BT_THREADS) blur_kernel(uint8_t* __restrict__ blur, const Geo* __restrict__ gp,
                                                          const FastMaps* __restrict__ maps, const uint32_t* __restrict__ tiles) {
    __shared__ __align__(128) uint8_t s_box[BB_H][BB_W];
    __shared__ __align__(8) uint64_t s_bar;

    const Geo& g = *gp;
    const uint32_t tref = __ldg(tiles + blockIdx.x);   // level | tx0 << 4 | ty0 << 16
    const int l = tref & 15, tx0 = (tref >> 4) & 0xFFF, ty0 = tref >> 16;
    const OrbLevelGeom& L = g.lv[l];
    const int img = blockIdx.y;
    uint8_t* dst = blur + (size_t)img * g.pyr_img_bytes + L.img_off;

    if (threadIdx.x == 0) {
        tma_mbar_init(&s_bar, 1);
        tma_mbar_expect_tx(&s_bar, BB_H * BB_W);
        tma_load_3d(&s_box[0][0], &maps->lv[l], tx0 - 16, ty0 - 4, img, &s_bar);
    }
    __syncthreads();
    tma_mbar_wait(&s_bar, 0);

    // BORDER_REFLECT_101 on the tiles that touch the level's border: rows first (all columns), then columns (all rows)
    const bool top = ty0 == 0, bottom = ty0 + BT_H + 3 > L.h, left = tx0 == 0, right = tx0 + BT_W + 3 > L.w;
    if (top || bottom) {
        for (int i = threadIdx.x; i < 6 * (BB_W / 4); i += BT_THREADS) {
            const int k = i / (BB_W / 4), cw = i - k * (BB_W / 4);
            int y, ys;   // image row to fill, its mirror
            if (k < 3) { y = -1 - k; ys = 1 + k; if (!top) continue; }
            else { y = L.h + (k - 3); ys = L.h - 2 - (k - 3); if (!bottom) continue; }
            const int r = y - (ty0 - 4), rs = ys - (ty0 - 4);
            if (r < 0 || r >= BB_H || rs < 0 || rs >= BB_H) continue;
            reinterpret_cast<uint32_t*>(&s_box[r][0])[cw] = reinterpret_cast<const uint32_t*>(&s_box[rs][0])[cw];
        }
        __syncthreads();
    }
    if (left || right) {
        for (int i = threadIdx.x; i < 6 * BB_H; i += BT_THREADS) {
            const int k = i / BB_H, r = i - k * BB_H;
            int x, xs;
            if (k < 3) { x = -1 - k; xs = 1 + k; if (!left) continue; }
            else { x = L.w + (k - 3); xs = L.w - 2 - (k - 3); if (!right) continue; }
            const int c = x - (tx0 - 16), cs = xs - (tx0 - 16);
            if (c < 0 || c >= BB_W || cs < 0 || cs >= BB_W) continue;
            s_box[r][c] = s_box[r][cs];
        }
        __syncthreads();
    }

    const int lane = threadIdx.x & 31, strip = threadIdx.x >> 5;
    const int gx = tx0 + 4 * lane, gy0 = ty0 + strip * BS_H;
    if (gx >= L.w || gy0 >= L.h) return;
    // getGaussianKernel(7, 2, CV_32F) as exact bit patterns
    const float k0 = __uint_as_float(0x3d8fafb1u), k1 = __uint_as_float(0x3e06387eu), k2 = __uint_as_float(0x3e434a39u),
                k3 = __uint_as_float(0x3e5d4ae0u);
    float h[7][4];   // row-pass results of the last 7 input rows (rotating window; all indices are compile-time after unrolling)
#pragma unroll
    for (int i = 0; i < BS_H + 6; ++i) {
        // input row gy0 - 3 + i == box row strip * BS_H + 1 + i; the thread's 12 bytes start at box column 4 * lane + 12
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(&s_box[strip * BS_H + 1 + i][0]) + lane + 3;
        const uint32_t w0 = rw[0], w1 = rw[1], w2 = rw[2];
        // u8 -> fp32 without the conversion pipe: byte b | 0x4B000000 is the float 2^23 + b, exactly; subtract 2^23
        float p[12];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            p[b] = __uint_as_float(__byte_perm(w0, 0x4B000000u, 0x7540 + b)) - 8388608.f;
            p[4 + b] = __uint_as_float(__byte_perm(w1, 0x4B000000u, 0x7540 + b)) - 8388608.f;
            p[8 + b] = __uint_as_float(__byte_perm(w2, 0x4B000000u, 0x7540 + b)) - 8388608.f;
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            float acc = __fmul_rn(k0, p[b + 1]);
            acc = __fmaf_rn(k1, p[b + 2], acc);
            acc = __fmaf_rn(k2, p[b + 3], acc);
            acc = __fmaf_rn(k3, p[b + 4], acc);
            acc = __fmaf_rn(k2, p[b + 5], acc);
            acc = __fmaf_rn(k1, p[b + 6], acc);
            acc = __fmaf_rn(k0, p[b + 7], acc);
            h[i % 7][b] = acc;
        }
        if (i >= 6) {
            const int j = i - 6, gy = gy0 + j;   // output row j: input rows j .. j + 6 of the strip = window slots (j + m) % 7
            if (gy < L.h) {
                uint32_t out = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    float acc = __fmul_rn(k3, h[(j + 3) % 7][b]);
                    acc = __fmaf_rn(k2, __fadd_rn(h[(j + 4) % 7][b], h[(j + 2) % 7][b]), acc);
                    acc = __fmaf_rn(k1, __fadd_rn(h[(j + 5) % 7][b], h[(j + 1) % 7][b]), acc);
                    acc = __fmaf_rn(k0, __fadd_rn(h[(j + 6) % 7][b], h[j % 7][b]), acc);
                    // round-half-even without F2I: acc + 1.5 * 2^23 has the rounded integer in its low mantissa bits (0 <= acc < 2^22)
                    const int v = __float_as_int(__fadd_rn(acc, 12582912.f)) - 0x4B400000;
                    out |= (uint32_t)min(max(v, 0), 255) << (8 * b);
                }
                *reinterpret_cast<uint32_t*>(dst + (size_t)gy * L.pitch + gx) = out;
            }
        }
    }
}

// ---- K5 + K7 ------------------------------------------------------------------------------------------------
// cv::fastAtan2 (degrees), float products/sums evaluated without contraction
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float p1 = __uint_as_float(0x4265226fu), p3 = __uint_as_float(0xc19556eeu), p5 = __uint_as_float(0x410e9fbfu),
                p7 = __uint_as_float(0xc0228ad9u);
    const float eps = __uint_as_float(0x25800000u);  // (float)DBL_EPSILON
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0.f) a = __fsub_rn(180.f, a);
    if (y < 0.f) a = __fsub_rn(360.f, a);
    return a;
}

// SAFE: the disc of the moments (radius half, read in whole words: half + 3) and the reach of the steered pattern (<= half*sqrt(2)
// + 1) stay inside the level for every keypoint runByImageBorder(edge) keeps, so no coordinate is ever reflected.  True for the
// reference's configuration (edge 62, patch 62).  The moments then run on packed bytes: a row of the disc is 16-17 aligned words,
// sum(u * I) over a word = u0 * dp4a(w, 0x01010101) + dp4a(w, 0x03020100), the disc boundary is a byte mask on the two end words.
template <bool SAFE>
__global__ void __launch_bounds__(256) describe_kernel(const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ blur,
                                                       const Geo* __restrict__ gp, const uint32_t* __restrict__ kp_src,
                                                       const int32_t* __restrict__ info, const int8_t* __restrict__ pattern,
                                                       float* __restrict__ kp_f, uint8_t* __restrict__ desc, float2* __restrict__ pt_out,
                                                       int stream_mode) {
    const int lane = threadIdx.x & 31;
    const int total = info[0];
    const size_t out_base = stream_mode ? (size_t)info[1] : 0;   // descriptors / points go to the stream position of the batch
    const Geo& g = *gp;
    const int half = g.half;
    const uint4* pat4 = reinterpret_cast<const uint4*>(pattern) + lane * 2;
    const uint4 w0 = __ldg(pat4), w1 = __ldg(pat4 + 1);
    const uint32_t words[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    for (int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < total; k += gridDim.x * (blockDim.x >> 5)) {
        const uint32_t v = kp_src[2 * k], fl = kp_src[2 * k + 1];
        const int x = v & 0xFFF, y = (v >> 12) & 0xFFF, l = fl & 0xFF, img = fl >> 8;
        const OrbLevelGeom L = g.lv[l];
        const uint8_t* im = pyr + (size_t)img * g.pyr_img_bytes + L.img_off;

        // K5: intensity centroid over the disc of radius `half` on the unblurred level
        int m10 = 0, m01 = 0;
        if (SAFE) {
            const int a = (x - half) & 3;
            const uint8_t* col0 = im + (x - half - a);                      // 4-byte aligned (rows are 16-byte aligned)
            for (int vv = -half + lane; vv <= half; vv += 32) {
                const int du = g.umax[vv < 0 ? -vv : vv];
                const int p_lo = a + half - du, p_hi = a + half + du;        // byte positions of the row's first / last disc pixel
                const int j_lo = p_lo >> 2, j_hi = p_hi >> 2;
                const uint32_t* row = reinterpret_cast<const uint32_t*>(col0 + (size_t)(y + vv) * L.pitch);
                int srow = 0, mrow = 0;
                for (int j = j_lo; j <= j_hi; ++j) {
                    uint32_t w = __ldg(row + j);
                    if (j == j_lo) w &= 0xFFFFFFFFu << (8 * (p_lo & 3));
                    if (j == j_hi) w &= 0xFFFFFFFFu >> (8 * (3 - (p_hi & 3)));
                    const int S = (int)__dp4a(w, 0x01010101u, 0u), T = (int)__dp4a(w, 0x03020100u, 0u);
                    mrow += (4 * j - a - half) * S + T;
                    srow += S;
                }
                m10 += mrow;
                m01 += vv * srow;
            }
        } else {
            for (int vv = -half; vv <= half; ++vv) {
                const int du = g.umax[vv < 0 ? -vv : vv];
                const int yy = reflect101(y + vv, L.h);
                const uint8_t* row = im + (size_t)yy * L.pitch;
                for (int u = -half + lane; u <= half; u += 32) {
                    if (u >= -du && u <= du) {
                        const int val = row[reflect101(x + u, L.w)];
                        m10 += u * val;
                        m01 += vv * val;
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m10 += __shfl_xor_sync(FULL, m10, o);
            m01 += __shfl_xor_sync(FULL, m01, o);
        }
        const float angle = fast_atan2_deg((float)m01, (float)m10);
        const float ptx = __fmul_rn((float)x, L.scale), pty = __fmul_rn((float)y, L.scale);
        if (lane == 0) {
            reinterpret_cast<float4*>(kp_f)[k] = make_float4(ptx, pty, __fmul_rn((float)g.patch, L.scale), angle);
            if (pt_out) pt_out[out_base + k] = make_float2(ptx, pty);
        }

        // K7: steered BRIEF on the blurred level (pattern = cv::RNG(0x34985739) points)
        const uint8_t* bl = blur + (size_t)img * g.pyr_img_bytes + L.img_off;
        const int cx = __float2int_rn(__fmul_rn(ptx, L.inv_scale)), cy = __float2int_rn(__fmul_rn(pty, L.inv_scale));
        const float th = __fmul_rn(angle, __uint_as_float(0x3c8efa35u));  // (float)(CV_PI / 180)
        const float a = (float)cos((double)th), b = (float)sin((double)th);
        const uint8_t* ctr = bl + (size_t)cy * L.pitch + cx;
        uint32_t byte = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            // word j holds the two points of bit j: (x0, y0, x1, y1) as int8
            // conversions without the quarter-rate XU pipe: int8 -> fp32 as (byte ^ 0x80) | 0x4B000000 = 2^23 + 128 + value, exactly;
            // cvRound as the low mantissa bits of v + 1.5 * 2^23 (round-half-even, |v| < 2^22)
            const uint32_t wv = words[j] ^ 0x80808080u;
            int val[2];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const float px = __fsub_rn(__uint_as_float(__byte_perm(wv, 0x4B000000u, 0x7540 + 2 * t)), 8388736.f);
                const float py = __fsub_rn(__uint_as_float(__byte_perm(wv, 0x4B000000u, 0x7541 + 2 * t)), 8388736.f);
                const int ix = __float_as_int(__fadd_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)), 12582912.f)) - 0x4B400000;
                const int iy = __float_as_int(__fadd_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)), 12582912.f)) - 0x4B400000;
                if (SAFE) val[t] = ctr[iy * L.pitch + ix];
                else val[t] = bl[(size_t)reflect101(cy + iy, L.h) * L.pitch + reflect101(cx + ix, L.w)];
            }
            byte |= (uint32_t)(val[0] < val[1]) << j;
        }
        desc[(out_base + k) * 32 + lane] = (uint8_t)byte;
    }
}

// host-side restatement of the geometry helpers (same arithmetic as oracle/orb_oracle.c)
int cv_round_f(float v) { return (int)lrintf(v); }
int cv_round_d(double v) { return (int)lrint(v); }

void build_axis_table(int src, int dst, int32_t* ofs, int32_t* c1) {
    const double scale = (double)src / (double)dst;
    for (int d = 0; d < dst; ++d) {
        const double val = scale * (d + 0.5) - 0.5;
        int o = (int)floor(val);
        int f = cv_round_d((val - o) * 256.0);
        if (o < 0) { o = 0; f = 0; }
        else if (o >= src - 1) { o = src - 1; f = 0; }
        ofs[d] = o;
        c1[d] = f;
    }
}

int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
OrbExtractor::OrbExtractor(const OrbConfig& cfg, int w, int h, int batch_cap) : cfg_(cfg), w_(w), h_(h), batch_cap_(batch_cap) {
    if (cfg.nlevels < 1 || cfg.nlevels > ORB_MAX_LEVELS) throw ArgError("nlevels out of range");
    if (w > ORB_MAX_DIM || h > ORB_MAX_DIM || w < 16 || h < 16) throw ArgError("image size out of range (16..4095)");
    if (batch_cap < 1 || batch_cap > 1024) throw ArgError("batch out of range (1..1024)");
    if (cfg.patch_size / 2 > 38 || cfg.patch_size < 2) throw ArgError("patch_size out of range");
    const int L = cfg.nlevels;
    lv_.resize(L);
    // quota per level (orb.cpp), SURVEY A.5
    {
        const float factor = (float)(1.0 / cfg.scale_factor);
        float nd = cfg.nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)L));
        int sum = 0;
        for (int l = 0; l < L - 1; ++l) { lv_[l].quota = cv_round_f(nd); sum += lv_[l].quota; nd *= factor; }
        lv_[L - 1].quota = cfg.nfeatures - sum > 0 ? cfg.nfeatures - sum : 0;
    }
    size_t img_off = 0, cand_off = 0, sel_off = 0, tab_off = 0;
    int tile_base = 0, btile_base = 0;
    std::vector<int32_t> tables;
    for (int l = 0; l < L; ++l) {
        OrbLevelGeom& g = lv_[l];
        g.scale = (float)pow((double)cfg.scale_factor, (double)l);
        g.inv_scale = 1.f / g.scale;
        g.w = cv_round_f((float)w * g.inv_scale);
        g.h = cv_round_f((float)h * g.inv_scale);
        if (g.w < 8 || g.h < 8) throw ArgError("image too small for the requested pyramid");
        g.pitch = (int)align_up((size_t)g.w, 16);
        g.img_off = img_off;
        img_off += align_up((size_t)g.pitch * g.h, 256);
        g.cand_cap = (g.w / 2 + 1) * (g.h / 2 + 1);  // strict 3x3 maxima: at most one per 2x2 block -> cannot overflow
        g.cand_off = cand_off;
        cand_off += (size_t)g.cand_cap;
        g.sel_cap = next_pow2(g.quota * 2 > g.quota + 256 ? g.quota * 2 : g.quota + 256);
        if (g.sel_cap > 8192) throw ArgError("nfeatures too large");
        g.sel_off = sel_off;
        sel_off += (size_t)g.sel_cap;
        g.tiles_x = cdiv(g.w, TILE_W);
        g.tiles_y = cdiv(g.h, TILE_H);
        g.tile_base = tile_base;
        tile_base += g.tiles_x * g.tiles_y;
        g.btiles_x = cdiv(g.w, BT_W);
        g.btiles_y = cdiv(g.h, BT_H);
        g.btile_base = btile_base;
        btile_base += g.btiles_x * g.btiles_y;
        g.tab_off = tab_off;
        if (l > 0) {
            tables.resize(tab_off + 2 * (size_t)(g.w + g.h));
            build_axis_table(lv_[l - 1].w, g.w, tables.data() + tab_off, tables.data() + tab_off + g.w);
            build_axis_table(lv_[l - 1].h, g.h, tables.data() + tab_off + 2 * g.w, tables.data() + tab_off + 2 * g.w + g.h);
            tab_off += 2 * (size_t)(g.w + g.h);
        }
    }
    total_tiles_ = tile_base;
    total_btiles_ = btile_base;
    {   // flat tile index -> level | tx0 << 4 | ty0 << 16 for FAST's and the blur's tilings
        std::vector<uint32_t> tf, tb;
        for (int l = 0; l < L; ++l) {
            for (int ty = 0; ty < lv_[l].tiles_y; ++ty)
                for (int tx = 0; tx < lv_[l].tiles_x; ++tx) tf.push_back((uint32_t)l | (uint32_t)(tx * TILE_W) << 4 | (uint32_t)(ty * TILE_H) << 16);
            for (int ty = 0; ty < lv_[l].btiles_y; ++ty)
                for (int tx = 0; tx < lv_[l].btiles_x; ++tx) tb.push_back((uint32_t)l | (uint32_t)(tx * BT_W) << 4 | (uint32_t)(ty * BT_H) << 16);
        }
        SLIDEO_CUDA(cudaMalloc(&d_tiles_fast_, tf.size() * 4));
        SLIDEO_CUDA(cudaMemcpy(d_tiles_fast_, tf.data(), tf.size() * 4, cudaMemcpyHostToDevice));
        SLIDEO_CUDA(cudaMalloc(&d_tiles_blur_, tb.size() * 4));
        SLIDEO_CUDA(cudaMemcpy(d_tiles_blur_, tb.data(), tb.size() * 4, cudaMemcpyHostToDevice));
    }
    pyr_img_bytes_ = img_off;
    cand_img_words_ = cand_off;
    sel_img_words_ = sel_off;
    kp_cap_ = (size_t)batch_cap * sel_off;

    const size_t B = (size_t)batch_cap;
    SLIDEO_CUDA(cudaMalloc(&d_pyr_, B * pyr_img_bytes_));
    SLIDEO_CUDA(cudaMalloc(&d_blur_, B * pyr_img_bytes_));
    SLIDEO_CUDA(cudaMalloc(&d_cand_, B * cand_img_words_ * 4));
    SLIDEO_CUDA(cudaMalloc(&d_sel_, B * sel_img_words_ * 4));
    SLIDEO_CUDA(cudaMalloc(&d_cand_cnt_, B * L * 4));
    SLIDEO_CUDA(cudaMalloc(&d_sel_cnt_, B * L * 4));
    SLIDEO_CUDA(cudaMalloc(&d_kp_off_, (B * L + 1) * 4));
    SLIDEO_CUDA(cudaMalloc(&d_frame_off_, (B + 1) * 4));
    SLIDEO_CUDA(cudaMalloc(&d_frame_nkp_, B * 4));
    SLIDEO_CUDA(cudaMalloc(&d_flags_, 4));
    SLIDEO_CUDA(cudaMalloc(&d_info_, 16));
    SLIDEO_CUDA(cudaMalloc(&d_kp_src_, kp_cap_ * 8));
    SLIDEO_CUDA(cudaMalloc(&d_q_frame_, kp_cap_ * 4));
    SLIDEO_CUDA(cudaMalloc(&d_kp_i_, kp_cap_ * 16));
    SLIDEO_CUDA(cudaMalloc(&d_kp_f_, kp_cap_ * 16));
    SLIDEO_CUDA(cudaMalloc(&d_desc_, kp_cap_ * 32));
    SLIDEO_CUDA(cudaMalloc(&d_tables_, (tables.size() + 4) * 4));
    if (!tables.empty())
        SLIDEO_CUDA(cudaMemcpy(d_tables_, tables.data(), tables.size() * 4, cudaMemcpyHostToDevice));
    SLIDEO_CUDA(cudaMallocHost(&h_pinned_, (B + 4) * 4));

    // sampling pattern (A.8) and umax (A.6)
    {
        const int half = cfg.patch_size / 2;
        int8_t pat[1024];
        uint64_t s = 0x34985739u;
        for (int i = 0; i < 1024; ++i) {
            s = (uint64_t)(uint32_t)s * 4164903690u + (s >> 32);
            pat[i] = (int8_t)(-half + (int)((uint32_t)s % (uint32_t)(2 * half + 1)));
        }
        SLIDEO_CUDA(cudaMalloc(&d_pattern_, 1024));
        SLIDEO_CUDA(cudaMemcpy(d_pattern_, pat, 1024, cudaMemcpyHostToDevice));
        int* umax = umax_;
        memset(umax_, 0, sizeof umax_);
        const int vmax = (int)floor(half * sqrt(2.0) / 2 + 1), vmin = (int)ceil(half * sqrt(2.0) / 2);
        for (int v = 0; v <= vmax; ++v) umax[v] = cv_round_d(sqrt((double)half * half - (double)v * v));
        for (int v = half, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
    }
    {   // see describe_kernel<SAFE>
        const int half = cfg.patch_size / 2;
        const int reach = (int)ceil(half * sqrt(2.0)) + 1;
        safe_ = cfg.edge_threshold >= half + 4 && cfg.edge_threshold >= reach + 1;
    }
    {   // tensor maps of the pyramid levels for the FAST tile loads (width = the padded row, so the box never depends on W % 4)
        FastMaps fm;
        memset(&fm, 0, sizeof fm);
        for (int l = 0; l < L; ++l)
            fm.lv[l] = tma_map_3d(CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, d_pyr_ + lv_[l].img_off, (uint64_t)lv_[l].pitch, (uint64_t)lv_[l].h, (uint64_t)batch_cap,
                                  (uint64_t)lv_[l].pitch, (uint64_t)pyr_img_bytes_, TILE_W + 32, TILE_H + 8);
        SLIDEO_CUDA(cudaMalloc(&fast_maps_, sizeof fm));   // the maps live in global memory (64-byte aligned by cudaMalloc)
        SLIDEO_CUDA(cudaMemcpy(fast_maps_, &fm, sizeof fm, cudaMemcpyHostToDevice));
        for (int l = 0; l < L; ++l)                        // the blur's boxes: BB_W x BB_H of the same tensors
            fm.lv[l] = tma_map_3d(CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, d_pyr_ + lv_[l].img_off, (uint64_t)lv_[l].pitch, (uint64_t)lv_[l].h, (uint64_t)batch_cap,
                                  (uint64_t)lv_[l].pitch, (uint64_t)pyr_img_bytes_, BB_W, BB_H);
        SLIDEO_CUDA(cudaMalloc(&blur_maps_, sizeof fm));
        SLIDEO_CUDA(cudaMemcpy(blur_maps_, &fm, sizeof fm, cudaMemcpyHostToDevice));
    }
    {
        Geo g;
        memset(&g, 0, sizeof g);
        g.nlevels = L; g.total_tiles = total_tiles_; g.edge = cfg_.edge_threshold; g.fast_thr = cfg_.fast_threshold;
        g.patch = cfg_.patch_size; g.half = cfg_.patch_size / 2;
        g.pyr_img_bytes = pyr_img_bytes_; g.cand_img_words = cand_img_words_; g.sel_img_words = sel_img_words_;
        memcpy(g.umax, umax_, sizeof g.umax);
        for (int l = 0; l < L; ++l) g.lv[l] = lv_[l];
        SLIDEO_CUDA(cudaMalloc(&d_geom_, sizeof g));
        SLIDEO_CUDA(cudaMemcpy(d_geom_, &g, sizeof g, cudaMemcpyHostToDevice));
    }
}

OrbExtractor::~OrbExtractor() {
    cudaFree(d_pyr_); cudaFree(d_blur_); cudaFree(d_cand_); cudaFree(d_sel_); cudaFree(d_cand_cnt_); cudaFree(d_sel_cnt_);
    cudaFree(d_kp_off_); cudaFree(d_frame_off_); cudaFree(d_frame_nkp_); cudaFree(d_flags_); cudaFree(d_info_); cudaFree(d_kp_src_);
    cudaFree(d_q_frame_); cudaFree(d_kp_i_); cudaFree(d_kp_f_); cudaFree(d_desc_); cudaFree(d_tables_);
    cudaFree(d_pattern_); cudaFree(d_geom_);
    cudaFree(fast_maps_);
    cudaFree(blur_maps_);
    cudaFree(d_tiles_fast_);
    cudaFree(d_tiles_blur_);
    cudaFreeHost(h_pinned_);
}

void OrbExtractor::enqueue(const uint8_t* d_src, int n, int stride, size_t frame_stride, int channels, cudaStream_t stream,
                           int* launches, const StreamSink* sink, int num_sms) {
    if (n < 1 || n > batch_cap_) throw ArgError("batch size out of range");
    if (channels != 1 && channels != 3) throw ArgError("channels must be 1 or 3");
    const int L = cfg_.nlevels;
    const Geo* g = static_cast<const Geo*>(d_geom_);
    int nl = 0;

    SLIDEO_CUDA(cudaMemsetAsync(d_cand_cnt_, 0, (size_t)n * L * 4, stream));
    SLIDEO_CUDA(cudaMemsetAsync(d_flags_, 0, 4, stream));
    {
        dim3 grid(cdiv(cdiv(w_, 16), 128), h_, n);
        gray_kernel<<<grid, 128, 0, stream>>>(d_src, stride, frame_stride, channels, d_pyr_, pyr_img_bytes_, w_, h_, lv_[0].pitch);
        ++nl;
    }
    for (int l = 1; l < L; ++l) {
        const OrbLevelGeom &s = lv_[l - 1], &d = lv_[l];
        dim3 grid(cdiv(cdiv(d.w, 4), 128), cdiv(d.h, RS_H), n);
        resize_kernel<<<grid, 128, 0, stream>>>(d_pyr_, pyr_img_bytes_, s.img_off, s.w, s.h, s.pitch, d.img_off, d.w, d.h,
                                                d.pitch, d_tables_ + d.tab_off);
        ++nl;
    }
    fast_kernel<<<dim3(total_tiles_, n), 256, 0, stream>>>(static_cast<const FastMaps*>(fast_maps_), g, d_tiles_fast_, d_cand_, d_cand_cnt_);
    ++nl;
    {
        int max_sel = 0;
        for (int l = 0; l < L; ++l) max_sel = lv_[l].sel_cap > max_sel ? lv_[l].sel_cap : max_sel;
        select_kernel<<<dim3(L, n), 256, (size_t)max_sel * 4, stream>>>(g, d_cand_, d_cand_cnt_, d_sel_, d_sel_cnt_, d_flags_);
        ++nl;
    }
    int32_t* d_hout = nullptr;
    if (!sink) SLIDEO_CUDA(cudaHostGetDevicePointer(&d_hout, h_pinned_, 0));
    scan_kernel<<<1, 256, 0, stream>>>(d_sel_cnt_, n, L, d_kp_off_, d_frame_off_, d_frame_nkp_, d_flags_, d_info_, d_hout,
                                       sink ? sink->st : nullptr, sink ? sink->frame_q0 : nullptr, sink ? sink->frame_nkp : nullptr,
                                       sink ? sink->q_cap : 0, sink ? sink->f_cap : 0, (int)kp_cap_);
    ++nl;
    scatter_kernel<<<dim3(L, n), 256, 0, stream>>>(g, d_sel_, d_sel_cnt_, d_kp_off_, d_kp_src_, d_q_frame_, d_kp_i_, kp_cap_, d_info_,
                                                   sink ? sink->q_frame : nullptr);
    ++nl;
    blur_kernel<<<dim3(total_btiles_, n), BT_THREADS, 0, stream>>>(d_blur_, g, static_cast<const FastMaps*>(blur_maps_), d_tiles_blur_);
    ++nl;
    {
        // the keypoint total of the batch lives on the device: a fixed grid of warps strides over it
        const size_t per_img = kp_cap_ / (size_t)batch_cap_;
        const int grid = (int)std::min<size_t>((size_t)num_sms * 8, (size_t)cdiv((int)std::min<size_t>(per_img * n, 1u << 30), 8));
        uint8_t* dd = sink ? sink->desc : d_desc_;
        float2* pp = sink ? sink->pt : nullptr;
        if (safe_) describe_kernel<true><<<grid, 256, 0, stream>>>(d_pyr_, d_blur_, g, d_kp_src_, d_info_, d_pattern_, d_kp_f_, dd, pp, sink ? 1 : 0);
        else describe_kernel<false><<<grid, 256, 0, stream>>>(d_pyr_, d_blur_, g, d_kp_src_, d_info_, d_pattern_, d_kp_f_, dd, pp, sink ? 1 : 0);
        ++nl;
    }
    SLIDEO_CUDA(cudaGetLastError());
    if (launches) *launches += nl;
}

int OrbExtractor::run(const uint8_t* d_src, int n, int stride, size_t frame_stride, int channels, cudaStream_t stream,
                      int* launches, int num_sms) {
    enqueue(d_src, n, stride, frame_stride, channels, stream, launches, nullptr, num_sms);
    SLIDEO_CUDA(cudaStreamSynchronize(stream));   // single images (pages, stage-level extraction): the caller wants the count
    const int total = h_pinned_[0], flags = h_pinned_[1];
    h_frame_off_.assign(h_pinned_ + 2, h_pinned_ + 2 + n + 1);
    if (flags & 1) throw CapacityError("FAST candidate capacity exceeded on at least one image");
    if (flags & 2) throw CapacityError("selected-keypoint capacity exceeded on at least one image");
    if ((flags & 4) || (size_t)total > kp_cap_) throw CapacityError("keypoint capacity exceeded");
    return total;
}

void OrbExtractor::run_stream(const uint8_t* d_src, int n, int stride, size_t frame_stride, int channels, cudaStream_t stream,
                              int* launches, const StreamSink& sink, int num_sms) {
    enqueue(d_src, n, stride, frame_stride, channels, stream, launches, &sink, num_sms);
}

}  // namespace slideo
