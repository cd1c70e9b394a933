// verify.cu -- K12: geometric verification of the per-frame candidates.
//
// Replaces the middle of match_images_with_frame (crates/matching-opencv/src/lib.rs:284-333): slides ranked by votes,
// take(40) -> per candidate `estimate_affine_partial_2d(slide kp -> frame kp, RANSAC, 3.0, 2000, 0.99, 10)`
// (image_utils.rs:45-60) -> rating = inlier count -> sort by rating, truncate(10), retain(rating > 50 && rating/best > 0.2).
// The algorithm restated is OpenCV's RANSACPointSetRegistrator::run with AffinePartial2DEstimatorCallback as pinned in
// oracle/ransac_oracle.c: cv::RNG((uint64)-1) sample sequence (a pure function of the correspondence count), closed-form
// similarity through two correspondences in double, fp32 residuals evaluated without contraction, adaptive iteration count.
// All hypotheses of a chunk are evaluated in parallel (one warp per hypothesis); the sequential best-so-far / niters logic is
// replayed over the chunk by one thread, so the result is identical to the sequential loop.
// Compiled with -fmad=false: no floating-point contraction anywhere in this file.
#include "verify.cuh"

#include <float.h>

namespace slideo {

namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int V_THREADS = 256;
constexpr int V_WARPS = V_THREADS / 32;
constexpr int V_SMEM_PTS = 2560;          // correspondences staged in shared memory (float4 each = 40 KB)
constexpr int V_CHUNK = 32;               // hypotheses per round (4 per warp)

// ---- A: candidates = the 40 pages with most votes (ties -> lower page index) ---------------------------------------
__global__ void __launch_bounds__(128) select_candidates_kernel(const int32_t* __restrict__ votes, int n_pages, int32_t* __restrict__ cand_page,
                                                                int32_t* __restrict__ cand_votes, int32_t* __restrict__ n_cand) {
    __shared__ unsigned long long s_best[4];
    const int f = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t* v = votes + (size_t)f * n_pages;
    // key = votes << 32 | (0xFFFFFFFF - page): unique per page, so round r simply takes the largest key below round r-1's
    unsigned long long prev = ~0ull;
    int n = 0;
    for (int r = 0; r < VERIFY_TOP_SLIDES; ++r) {
        unsigned long long best = 0;
        for (int p = threadIdx.x; p < n_pages; p += blockDim.x) {
            const int vv = v[p];
            if (vv > 0) {
                const unsigned long long key = ((unsigned long long)(uint32_t)vv << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)p);
                if (key < prev && key > best) best = key;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(FULL, best, o);
            best = other > best ? other : best;
        }
        if (lane == 0) s_best[warp] = best;
        __syncthreads();
        best = s_best[0];
        for (int w = 1; w < 4; ++w) best = s_best[w] > best ? s_best[w] : best;
        __syncthreads();
        if (best == 0) break;
        if (threadIdx.x == 0) {
            cand_page[(size_t)f * VERIFY_TOP_SLIDES + r] = (int32_t)(0xFFFFFFFFu - (uint32_t)best);
            cand_votes[(size_t)f * VERIFY_TOP_SLIDES + r] = (int32_t)(best >> 32);
        }
        prev = best;
        n = r + 1;
    }
    if (threadIdx.x == 0) n_cand[f] = n;
}

// ---- B: ordered correspondence lists: the voting matches of (frame, candidate page) in (query, rank) order ------------
__global__ void __launch_bounds__(V_THREADS) gather_matches_kernel(const uint32_t* __restrict__ keys, int k, const int32_t* __restrict__ frame_q0,
                                                                   const uint16_t* __restrict__ page_of, const int32_t* __restrict__ cand_page,
                                                                   const int32_t* __restrict__ cand_votes, const int32_t* __restrict__ n_cand, float ratio,
                                                                   uint2* __restrict__ corr) {
    __shared__ int s_warp[V_WARPS];
    __shared__ int s_base;
    const int c = blockIdx.x, f = blockIdx.y;
    if (c >= n_cand[f]) return;
    const int page = cand_page[(size_t)f * VERIFY_TOP_SLIDES + c];
    const int q0 = frame_q0[f], q1 = frame_q0[f + 1];
    const long long e0 = (long long)q0 * k, e1 = (long long)q1 * k;
    // the lists of one frame are packed back to back into the frame's own key range [q0 * k, q1 * k): a frame casts at most
    // one vote per k-NN entry, so the candidates' vote counts sum to no more than that
    long long off0 = e0;
    for (int cc = 0; cc < c; ++cc) off0 += cand_votes[(size_t)f * VERIFY_TOP_SLIDES + cc];
    uint2* out = corr + off0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (long long e = e0 + threadIdx.x; e - threadIdx.x < e1; e += V_THREADS) {
        bool take = false;
        uint32_t key = KEY_EMPTY;
        int q = 0;
        if (e < e1) {
            q = (int)(e / k);
            key = keys[e];
            const uint32_t best = keys[(long long)q * k];
            if (key != KEY_EMPTY) {
                const float d = (float)(key >> KEY_IDX_BITS), b = (float)(best >> KEY_IDX_BITS);
                take = d < __fmul_rn(b, ratio) && page_of[key & KEY_IDX_MASK] == (uint16_t)page;   // lib.rs:275
            }
        }
        const unsigned m = __ballot_sync(FULL, take);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += s_warp[w];
        if (take) out[off + __popc(m & ((1u << lane) - 1u))] = make_uint2((uint32_t)q, key & KEY_IDX_MASK);
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < V_WARPS; ++w) t += s_warp[w];
            s_base += t;
        }
        __syncthreads();
    }
}

// ---- C: RANSAC rating of one (frame, candidate) ------------------------------------------------------------------
__device__ __forceinline__ uint32_t rng_next(unsigned long long& s) {
    s = (unsigned long long)(uint32_t)s * 4164903690ull + (uint32_t)(s >> 32);
    return (uint32_t)s;
}

// RANSACUpdateNumIters(confidence, ep, 2, max_iters)
__device__ int update_num_iters(double p, double ep, int max_iters) {
    p = fmax(p, 0.); p = fmin(p, 1.);
    ep = fmax(ep, 0.); ep = fmin(ep, 1.);
    double num = fmax(1. - p, DBL_MIN);
    double denom = 1. - (1. - ep) * (1. - ep);   // std::pow(1 - ep, 2): correctly rounded square
    if (denom < DBL_MIN) return 0;
    num = log(num);
    denom = log(denom);
    return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : __double2int_rn(num / denom);
}

// The sample sequence of cv::RNG((uint64)-1) -- two distinct indices per iteration (getSubset; checkSubset never rejects a 2-point
// sample) -- is a pure function of the correspondence count n.  It is inherently serial (a multiply-with-carry generator and a
// rejection loop), and one thread spelling out 2000 pairs while its CTA waits used to be most of a small candidate's time.  The
// sequences of every n <= VERIFY_PAIRS_N (the wrong pages of a frame: a few dozen correspondences, never terminating early) are
// therefore built ONCE per context, one thread per n.
__global__ void __launch_bounds__(64) ransac_pairs_kernel(uint2* __restrict__ pairs, int max_iters) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < 3 || n > VERIFY_PAIRS_N) return;
    unsigned long long s = 0xFFFFFFFFFFFFFFFFull;
    uint2* out = pairs + (size_t)n * max_iters;
    for (int it = 0; it < max_iters; ++it) {
        const uint32_t i0 = rng_next(s) % (uint32_t)n;
        uint32_t i1;
        do { i1 = rng_next(s) % (uint32_t)n; } while (i1 == i0);
        out[it] = make_uint2(i0, i1);
    }
}

__global__ void __launch_bounds__(V_THREADS) ransac_kernel(const uint2* __restrict__ corr, const int32_t* __restrict__ frame_q0, int k,
                                                           const int32_t* __restrict__ cand_votes, const int32_t* __restrict__ n_cand,
                                                           const float2* __restrict__ frame_pt, const float2* __restrict__ pool_pt,
                                                           float thr, int max_iters, double confidence, int32_t* __restrict__ rating,
                                                           int32_t* __restrict__ best_it_out, const uint2* __restrict__ pair_table) {
    extern __shared__ __align__(16) uint8_t v_smem[];
    float4* s_pts = reinterpret_cast<float4*>(v_smem);                                        // [V_SMEM_PTS] (fx, fy, tx, ty)
    uint2* s_pairs = reinterpret_cast<uint2*>(v_smem + (size_t)V_SMEM_PTS * sizeof(float4));  // [V_CHUNK] sample indices of the current round (large n)
    __shared__ int s_good[V_THREADS];
    __shared__ int s_state[3];   // niters, max_good, iteration of the best model
    __shared__ unsigned long long s_rng;   // generator state behind the pairs spelled out so far (large n)

    const int c = blockIdx.x, f = blockIdx.y;
    if (threadIdx.x == 0 && best_it_out) best_it_out[(size_t)f * VERIFY_TOP_SLIDES + c] = -1;
    if (c >= n_cand[f]) {
        if (threadIdx.x == 0) rating[(size_t)f * VERIFY_TOP_SLIDES + c] = 0;
        return;
    }
    const int n = cand_votes[(size_t)f * VERIFY_TOP_SLIDES + c];
    long long off0 = (long long)frame_q0[f] * k;
    for (int cc = 0; cc < c; ++cc) off0 += cand_votes[(size_t)f * VERIFY_TOP_SLIDES + cc];
    const uint2* list = corr + off0;
    int32_t* out = rating + (size_t)f * VERIFY_TOP_SLIDES + c;
    if (n < 2) {          // fewer than modelPoints correspondences: no model, no inliers
        if (threadIdx.x == 0) *out = 0;
        return;
    }
    if (n == 2) {         // count == modelPoints: the model through both points, every point an inlier
        if (threadIdx.x == 0) *out = 2;
        return;
    }
    const bool in_smem = n <= V_SMEM_PTS;
    if (in_smem) {
        for (int i = threadIdx.x; i < n; i += V_THREADS) {
            const uint2 m = list[i];
            const float2 to = frame_pt[m.x], fr = pool_pt[m.y];
            s_pts[i] = make_float4(fr.x, fr.y, to.x, to.y);
        }
    }
    const bool small = n <= VERIFY_PAIRS_N && pair_table != nullptr;
    if (threadIdx.x == 0) {
        s_rng = 0xFFFFFFFFFFFFFFFFull;
        s_state[0] = max_iters > 1 ? max_iters : 1;
        s_state[1] = 0;
        s_state[2] = -1;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float t = (float)((double)thr * (double)thr);
    auto point = [&](int i) -> float4 {
        if (in_smem) return s_pts[i];
        const uint2 m = list[i];
        const float2 to = frame_pt[m.x], fr = pool_pt[m.y];
        return make_float4(fr.x, fr.y, to.x, to.y);
    };

    // closed-form model of a sample (AffinePartial2DEstimatorCallback::runKernel, double) and its inlier count over points [i0, n)
    // with stride `step` (Affine2DEstimatorCallback::computeError + findInliers: fp32, no contraction, err <= thr^2)
    auto count_inliers = [&](const uint2 pr, int i0, int step) {
        const float4 p0 = point((int)pr.x), p1 = point((int)pr.y);
        const double x1 = p0.x, y1 = p0.y, x2 = p1.x, y2 = p1.y, X1 = p0.z, Y1 = p0.w, X2 = p1.z, Y2 = p1.w;
        const double d = 1. / ((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2));
        const double S0 = d * ((X1 - X2) * (x1 - x2) + (Y1 - Y2) * (y1 - y2));
        const double S1 = d * ((Y1 - Y2) * (x1 - x2) - (X1 - X2) * (y1 - y2));
        const double S2 = d * ((Y1 - Y2) * (x1 * y2 - x2 * y1) - (X1 * y2 - X2 * y1) * (y1 - y2) - (X1 * x2 - X2 * x1) * (x1 - x2));
        const double S3 = d * (-(X1 - X2) * (x1 * y2 - x2 * y1) - (Y1 * x2 - Y2 * x1) * (x1 - x2) - (Y1 * y2 - Y2 * y1) * (y1 - y2));
        const float F0 = (float)S0, F1 = (float)(-S1), F2 = (float)S2, F3 = (float)S1, F4 = (float)S0, F5 = (float)S3;
        int good = 0;
        for (int i = i0; i < n; i += step) {
            const float4 p = point(i);
            const float a = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(F0, p.x), __fmul_rn(F1, p.y)), F2), p.z);
            const float b = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(F3, p.x), __fmul_rn(F4, p.y)), F5), p.w);
            const float e = __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
            good += e <= t ? 1 : 0;
        }
        return good;
    };
    // replay of the sequential loop over the hypotheses [it0, it0 + cnt) whose inlier counts sit in s_good: a hypothesis only
    // counts if the loop would still be running
    auto replay = [&](int it0, int cnt) {
        int ni = s_state[0], mg = s_state[1], bi = s_state[2];
        for (int h = 0; h < cnt; ++h) {
            const int it = it0 + h;
            if (it >= ni) break;
            const int good = s_good[h];
            if (good > (mg > 1 ? mg : 1)) {
                mg = good;
                bi = it;
                ni = update_num_iters(confidence, (double)(n - good) / n, ni);
            }
        }
        s_state[0] = ni;
        s_state[1] = mg;
        s_state[2] = bi;
    };

    if (small) {
        // few correspondences (the wrong pages among a frame's 40 candidates; the loop never terminates early for them): one
        // THREAD per hypothesis -- its own model, all n points from shared memory (broadcast reads) -- 256 hypotheses per round,
        // samples from the table
        const uint2* pairs = pair_table + (size_t)n * max_iters;
        for (int it0 = 0; ; it0 += V_THREADS) {
            const int niters = s_state[0];
            if (it0 >= niters) break;
            const int it = it0 + (int)threadIdx.x;
            s_good[threadIdx.x] = it < niters ? count_inliers(__ldg(pairs + it), 0, 1) : 0;
            __syncthreads();
            if (threadIdx.x == 0) replay(it0, V_THREADS);
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            *out = s_state[1];
            if (best_it_out) best_it_out[(size_t)f * VERIFY_TOP_SLIDES + c] = s_state[2];
        }
        return;
    }

    for (int it0 = 0; ; it0 += V_CHUNK) {
        const int niters = s_state[0];
        if (it0 >= niters) break;
        // many correspondences (the page the frame shows: the loop ends after a handful of hypotheses): one WARP per hypothesis,
        // lanes over the points; the samples of this round are spelled out now (not all max_iters up front)
        if (threadIdx.x == 0) {
            unsigned long long s = s_rng;
            for (int h = 0; h < V_CHUNK; ++h) {
                const uint32_t i0 = rng_next(s) % (uint32_t)n;
                uint32_t i1;
                do { i1 = rng_next(s) % (uint32_t)n; } while (i1 == i0);
                s_pairs[h] = make_uint2(i0, i1);
            }
            s_rng = s;
        }
        __syncthreads();
        for (int h = warp; h < V_CHUNK; h += V_WARPS) {
            int good = 0;
            if (it0 + h < niters) {
                good = count_inliers(s_pairs[h], lane, 32);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) good += __shfl_xor_sync(FULL, good, o);
            }
            if (lane == 0) s_good[h] = good;
        }
        __syncthreads();
        if (threadIdx.x == 0) replay(it0, V_CHUNK);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *out = s_state[1];
        if (best_it_out) best_it_out[(size_t)f * VERIFY_TOP_SLIDES + c] = s_state[2];
    }
}

// ---- D: gates (lib.rs:329-333): stable sort by rating, truncate(10), retain(rating > 50 && rating / best > 0.2) -----------
__global__ void __launch_bounds__(128) gate_kernel(const int32_t* __restrict__ cand_page, const int32_t* __restrict__ cand_votes,
                                                   const int32_t* __restrict__ rating, const int32_t* __restrict__ n_cand, int n_frames,
                                                   VerifyRecord* __restrict__ out, int32_t* __restrict__ survivor_cand) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    VerifyRecord r;
    const int n = n_cand[f];
    r.n_candidates = n;
    int order[VERIFY_TOP_SLIDES];
    for (int i = 0; i < VERIFY_TOP_SLIDES; ++i) {
        r.cand_page[i] = i < n ? cand_page[(size_t)f * VERIFY_TOP_SLIDES + i] : -1;
        r.cand_votes[i] = i < n ? cand_votes[(size_t)f * VERIFY_TOP_SLIDES + i] : 0;
        r.cand_rating[i] = i < n ? rating[(size_t)f * VERIFY_TOP_SLIDES + i] : 0;
        order[i] = i;
    }
    for (int i = 1; i < n; ++i) {   // stable insertion sort, descending rating
        const int o = order[i];
        int j = i - 1;
        while (j >= 0 && r.cand_rating[order[j]] < r.cand_rating[o]) { order[j + 1] = order[j]; --j; }
        order[j + 1] = o;
    }
    const int top = n < VERIFY_TOP_RATED ? n : VERIFY_TOP_RATED;
    const double best = top > 0 ? (double)r.cand_rating[order[0]] : 0.0;
    int ns = 0;
    for (int i = 0; i < VERIFY_TOP_RATED; ++i) { r.survivor_page[i] = -1; r.survivor_rating[i] = 0; }
    for (int i = 0; i < top; ++i) {
        const double v = (double)r.cand_rating[order[i]];
        if (v > 50.0 && v / best > 0.2) {
            r.survivor_page[ns] = r.cand_page[order[i]];
            r.survivor_rating[ns] = r.cand_rating[order[i]];
            if (survivor_cand) survivor_cand[(size_t)f * VERIFY_TOP_RATED + ns] = order[i];
            ++ns;
        }
    }
    r.n_survivors = ns;
    out[f] = r;
}


// ---- K14 a: Levenberg-Marquardt refinement of the survivors' RANSAC models (cv::LMSolver, 10 iterations) ------------------
__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
    return t;
}
__device__ __forceinline__ double block_max(double v, double* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t = fmax(t, s_red[w]);
    return t;
}
// solves the symmetric 4x4 system m x = b by Gaussian elimination with partial pivoting (double)
__device__ void solve4(const double (&m)[4][4], const double (&b)[4], double (&x)[4]) {
    double a[4][5];
    for (int i = 0; i < 4; ++i) { for (int j = 0; j < 4; ++j) a[i][j] = m[i][j]; a[i][4] = b[i]; }
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        for (int r = c + 1; r < 4; ++r) if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
        if (piv != c) for (int j = 0; j < 5; ++j) { const double t = a[c][j]; a[c][j] = a[piv][j]; a[piv][j] = t; }
        for (int r = c + 1; r < 4; ++r) {
            const double f = a[r][c] / a[c][c];
            for (int j = c; j < 5; ++j) a[r][j] -= f * a[c][j];
        }
    }
    for (int i = 3; i >= 0; --i) {
        double t = a[i][4];
        for (int j = i + 1; j < 4; ++j) t -= a[i][j] * x[j];
        x[i] = t / a[i][i];
    }
}

__global__ void __launch_bounds__(128) lm_refine_kernel(const PhotoArgs A) {
    __shared__ double s_red[4];
    __shared__ double s_h[4];       // parameters every thread evaluates residuals at
    const int j = blockIdx.x, f = blockIdx.y;
    const VerifyRecord& rec = A.d_records[f];
    if (j >= rec.n_survivors) return;
    const int c = A.d_survivor_cand[(size_t)f * VERIFY_TOP_RATED + j];
    const int n = A.d_cand_votes[(size_t)f * VERIFY_TOP_SLIDES + c];
    const int best_it = A.d_best_it[(size_t)f * VERIFY_TOP_SLIDES + c];
    long long off0 = (long long)A.d_frame_q0[f] * A.k;
    for (int cc = 0; cc < c; ++cc) off0 += A.d_cand_votes[(size_t)f * VERIFY_TOP_SLIDES + cc];
    const uint2* list = reinterpret_cast<const uint2*>(A.d_corr) + off0;
    auto point = [&](int i) -> float4 {
        const uint2 m = list[i];
        const float2 to = A.d_frame_pt[m.x], fr = A.d_pool_pt[m.y];
        return make_float4(fr.x, fr.y, to.x, to.y);
    };
    // the winning RANSAC model: replay the sample sequence up to best_it (pure function of n), closed form as in ransac_kernel
    if (threadIdx.x == 0) {
        unsigned long long s = 0xFFFFFFFFFFFFFFFFull;
        uint32_t i0 = 0, i1 = 0;
        for (int it = 0; it <= best_it; ++it) {
            i0 = rng_next(s) % (uint32_t)n;
            do { i1 = rng_next(s) % (uint32_t)n; } while (i1 == i0);
        }
        const float4 p0 = point((int)i0), p1 = point((int)i1);
        const double x1 = p0.x, y1 = p0.y, x2 = p1.x, y2 = p1.y, X1 = p0.z, Y1 = p0.w, X2 = p1.z, Y2 = p1.w;
        const double d = 1. / ((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2));
        s_h[0] = d * ((X1 - X2) * (x1 - x2) + (Y1 - Y2) * (y1 - y2));
        s_h[1] = d * ((Y1 - Y2) * (x1 - x2) - (X1 - X2) * (y1 - y2));
        s_h[2] = d * ((Y1 - Y2) * (x1 * y2 - x2 * y1) - (X1 * y2 - X2 * y1) * (y1 - y2) - (X1 * x2 - X2 * x1) * (x1 - x2));
        s_h[3] = d * (-(X1 - X2) * (x1 * y2 - x2 * y1) - (Y1 * x2 - Y2 * x1) * (x1 - x2) - (Y1 * y2 - Y2 * y1) * (y1 - y2));
    }
    __syncthreads();
    // inlier predicate of that model (== the RANSAC best mask): fp32, no contraction
    const float F0 = (float)s_h[0], F1 = (float)(-s_h[1]), F2 = (float)s_h[2], F3 = (float)s_h[1], F4 = (float)s_h[0], F5 = (float)s_h[3];
    auto inlier = [&](const float4& p) -> bool {
        const float a = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(F0, p.x), __fmul_rn(F1, p.y)), F2), p.z);
        const float b = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(F3, p.x), __fmul_rn(F4, p.y)), F5), p.w);
        return __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)) <= 9.0f;
    };
    // J^T J depends on the points only: [sxx 0 sx sy; 0 sxx -sy sx; sx -sy n 0; sy sx 0 n]
    double sxx = 0., sx = 0., sy = 0., cnt = 0.;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float4 p = point(i);
        if (inlier(p)) { sxx += (double)p.x * p.x + (double)p.y * p.y; sx += p.x; sy += p.y; cnt += 1.; }
    }
    sxx = block_sum(sxx, s_red); sx = block_sum(sx, s_red); sy = block_sum(sy, s_red); cnt = block_sum(cnt, s_red);
    // residual sums at parameters h: S = |r|^2, v = J^T r, rmax = |r|_inf
    auto residuals = [&](const double (&h)[4], double& S, double (&v)[4], double& rmax) {
        double aS = 0., a0 = 0., a1 = 0., a2 = 0., a3 = 0., am = 0.;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float4 p = point(i);
            if (inlier(p)) {
                const double Mx = p.x, My = p.y;
                const double rx = h[0] * Mx - h[1] * My + h[2] - (double)p.z;
                const double ry = h[1] * Mx + h[0] * My + h[3] - (double)p.w;
                aS += rx * rx + ry * ry;
                a0 += Mx * rx + My * ry;
                a1 += -My * rx + Mx * ry;
                a2 += rx;
                a3 += ry;
                am = fmax(am, fmax(fabs(rx), fabs(ry)));
            }
        }
        S = block_sum(aS, s_red); v[0] = block_sum(a0, s_red); v[1] = block_sum(a1, s_red); v[2] = block_sum(a2, s_red);
        v[3] = block_sum(a3, s_red); rmax = block_max(am, s_red);
    };
    const double Am[4][4] = {{sxx, 0., sx, sy}, {0., sxx, -sy, sx}, {sx, -sy, cnt, 0.}, {sy, sx, 0., cnt}};
    double x[4] = {s_h[0], s_h[1], s_h[2], s_h[3]}, v[4], S, rmax;
    residuals(x, S, v, rmax);
    double lambda = 1., lc = 0.75;
    for (int iter = 0; iter < 10;) {
        // every thread runs the identical scalar control flow on identical (block-reduced) values
        double Ap[4][4], d[4], xd[4];
        for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) Ap[a][b] = Am[a][b];
        for (int a = 0; a < 4; ++a) Ap[a][a] += lambda * Am[a][a];
        solve4(Ap, v, d);
        for (int a = 0; a < 4; ++a) xd[a] = x[a] - d[a];
        double Sd, vd[4], rmaxd;
        residuals(xd, Sd, vd, rmaxd);
        double dS = 0.;
        for (int a = 0; a < 4; ++a) {
            double t = 2. * v[a];
            for (int b = 0; b < 4; ++b) t -= Am[a][b] * d[b];
            dS += d[a] * t;
        }
        const double R = (S - Sd) / (fabs(dS) > DBL_EPSILON ? dS : 1.);
        if (R > 0.75) {
            lambda *= 0.5;
            if (lambda < lc) lambda = 0.;
        } else if (R < 0.25) {
            double t = 0.;
            for (int a = 0; a < 4; ++a) t += d[a] * v[a];
            double nu = (Sd - S) / (fabs(t) > DBL_EPSILON ? t : 1.) + 2.;
            nu = fmin(fmax(nu, 2.), 10.);
            if (lambda == 0.) {
                double maxval = DBL_EPSILON;
                for (int a = 0; a < 4; ++a) {   // diagonal of inverse(A)
                    double e[4] = {0., 0., 0., 0.}, col[4];
                    e[a] = 1.;
                    solve4(Am, e, col);
                    maxval = fmax(maxval, fabs(col[a]));
                }
                lambda = lc = 1. / maxval;
                nu *= 0.5;
            }
            lambda *= nu;
        }
        if (Sd < S) {
            S = Sd; rmax = rmaxd;
            for (int a = 0; a < 4; ++a) { x[a] = xd[a]; v[a] = vd[a]; }
        }
        ++iter;
        double dmax = 0.;
        for (int a = 0; a < 4; ++a) dmax = fmax(dmax, fabs(d[a]));
        if (!(iter < 10 && dmax >= (double)FLT_EPSILON && rmax >= (double)FLT_EPSILON)) break;
    }
    if (threadIdx.x == 0) {
        double* out = A.d_refined + ((size_t)f * VERIFY_TOP_RATED + j) * 4;
        out[0] = x[0]; out[1] = x[1]; out[2] = x[2]; out[3] = x[3];
    }
}

// ---- K14 b: warp (nearest, inverse map, AB_BITS = 10) + INTER_AREA small image + squared difference to the slide's small image
__global__ void __launch_bounds__(128) photometric_kernel(const PhotoArgs A) {
    __shared__ unsigned long long s_acc[4];
    const int f = blockIdx.z / VERIFY_TOP_RATED, j = blockIdx.z - f * VERIFY_TOP_RATED;
    const VerifyRecord& rec = A.d_records[f];
    if (j >= rec.n_survivors) return;
    const int page = rec.survivor_page[j];
    const PageGeom G = A.d_geom[A.d_page_class[page]];   // the slide's own size (the reference warps to slide_info.img.size())
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
    if (dy >= G.small_h) return;                          // block-uniform: the grid covers the largest small image of the deck
    unsigned long long acc = 0;
    if (dx < G.small_w) {
        const double* h = A.d_refined + ((size_t)f * VERIFY_TOP_RATED + j) * 4;
        const double M0 = h[0], M1 = -h[1], M2 = h[2], M3 = h[1], M4 = h[0], M5 = h[3];
        const uint8_t* frame = A.d_frames + (size_t)f * A.frame_stride;
        const int x0 = G.d_xoff[dx], x1 = G.d_xoff[dx + 1], y0 = G.d_yoff[dy], y1 = G.d_yoff[dy + 1];
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        for (int jj = y0; jj < y1; ++jj) {
            const int sy = G.d_ysi[jj];
            const float beta = G.d_ya[jj];
            const int X0 = __double2int_rn((M1 * (double)sy + M2) * 1024.) + 512;
            const int Y0 = __double2int_rn((M4 * (double)sy + M5) * 1024.) + 512;
            float b0 = 0.f, b1 = 0.f, b2 = 0.f;
            for (int kk = x0; kk < x1; ++kk) {
                const int sx = G.d_xsi[kk];
                const float al = G.d_xa[kk];
                const int X = (X0 + __double2int_rn(M0 * (double)sx * 1024.)) >> 10;
                const int Y = (Y0 + __double2int_rn(M3 * (double)sx * 1024.)) >> 10;
                float p0 = 0.f, p1 = 0.f, p2 = 0.f;   // BORDER_CONSTANT 0
                if (X >= 0 && X < A.frame_w && Y >= 0 && Y < A.frame_h) {
                    const uint8_t* px = frame + (size_t)Y * A.frame_stride_row + 3 * X;
                    p0 = (float)px[0]; p1 = (float)px[1]; p2 = (float)px[2];
                }
                b0 = __fadd_rn(b0, __fmul_rn(p0, al));
                b1 = __fadd_rn(b1, __fmul_rn(p1, al));
                b2 = __fadd_rn(b2, __fmul_rn(p2, al));
            }
            if (jj == y0) {
                s0 = __fmul_rn(beta, b0); s1 = __fmul_rn(beta, b1); s2 = __fmul_rn(beta, b2);
            } else {
                s0 = __fadd_rn(s0, __fmul_rn(beta, b0)); s1 = __fadd_rn(s1, __fmul_rn(beta, b1)); s2 = __fadd_rn(s2, __fmul_rn(beta, b2));
            }
        }
        const int g = A.d_page_small[A.d_page_small_off[page] + (size_t)dy * G.small_w + dx];
        const int v0 = min(max(__float2int_rn(s0), 0), 255) - g, v1 = min(max(__float2int_rn(s1), 0), 255) - g,
                  v2 = min(max(__float2int_rn(s2), 0), 255) - g;
        acc = (unsigned long long)(v0 * v0 + v1 * v1 + v2 * v2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
    if ((threadIdx.x & 31) == 0) s_acc[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long t = s_acc[0] + s_acc[1] + s_acc[2] + s_acc[3];
        if (t) atomicAdd(&A.d_sumsq[(size_t)f * VERIFY_TOP_RATED + j], t);
    }
}

}  // namespace

size_t verify_pairs_bytes() { return (size_t)(VERIFY_PAIRS_N + 1) * VERIFY_MAX_ITERS * sizeof(uint2); }
void verify_pairs_build(void* d_pairs, cudaStream_t stream) {
    ransac_pairs_kernel<<<cdiv(VERIFY_PAIRS_N + 1, 64), 64, 0, stream>>>((uint2*)d_pairs, VERIFY_MAX_ITERS);
    SLIDEO_CUDA(cudaGetLastError());
}

size_t verify_corr_bytes(long long total_entries) { return (size_t)(total_entries > 0 ? total_entries : 1) * sizeof(uint2); }

void verify_launch(const VerifyArgs& a, cudaStream_t stream, int* launches) {
    if (a.n_frames <= 0) return;
    select_candidates_kernel<<<a.n_frames, 128, 0, stream>>>(a.d_votes, a.n_pages, a.d_cand_page, a.d_cand_votes, a.d_n_cand);
    gather_matches_kernel<<<dim3(VERIFY_TOP_SLIDES, a.n_frames), V_THREADS, 0, stream>>>(a.d_keys, a.k, a.d_frame_q0, a.d_page_of, a.d_cand_page,
                                                                                     a.d_cand_votes, a.d_n_cand, a.ratio, (uint2*)a.d_corr);
    const size_t smem = (size_t)V_SMEM_PTS * sizeof(float4) + (size_t)V_CHUNK * sizeof(uint2);
    // function attributes are per device: set on every launch (cheap) so that ctxs on several GPUs of one process all get them
    SLIDEO_CUDA(cudaFuncSetAttribute(ransac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ransac_kernel<<<dim3(VERIFY_TOP_SLIDES, a.n_frames), V_THREADS, smem, stream>>>((const uint2*)a.d_corr, a.d_frame_q0, a.k, a.d_cand_votes,
                                                                                  a.d_n_cand, a.d_frame_pt, a.d_pool_pt, 3.0f, VERIFY_MAX_ITERS,
                                                                                  0.99, a.d_rating, a.d_best_it, a.d_pairs);
    gate_kernel<<<cdiv(a.n_frames, 128), 128, 0, stream>>>(a.d_cand_page, a.d_cand_votes, a.d_rating, a.d_n_cand, a.n_frames, a.d_out, a.d_survivor_cand);
    SLIDEO_CUDA(cudaGetLastError());
    if (launches) *launches += 4;
}

}  // namespace slideo

namespace slideo {

void photometric_launch(const PhotoArgs& a, cudaStream_t stream, int* launches) {
    if (a.n_frames <= 0) return;
    SLIDEO_CUDA(cudaMemsetAsync(a.d_sumsq, 0, (size_t)a.n_frames * VERIFY_TOP_RATED * sizeof(unsigned long long), stream));
    lm_refine_kernel<<<dim3(VERIFY_TOP_RATED, a.n_frames), 128, 0, stream>>>(a);
    photometric_kernel<<<dim3(cdiv(a.max_small_w, 128), a.max_small_h, a.n_frames * VERIFY_TOP_RATED), 128, 0, stream>>>(a);
    SLIDEO_CUDA(cudaGetLastError());
    if (launches) *launches += 2;
}

}  // namespace slideo
