/*
 * ransac_oracle.c -- CPU restatement (plain C) of the geometric verification step of the reference's per-frame path.
 * TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / reference arm).
 *
 * What it restates
 *   reference call sites: crates/matching-opencv/src/image_utils.rs:45-60 (estimate_affine_partial_2d(from, to, inliers,
 *                         RANSAC, 3.0, 2000, 0.99, 10)) called per candidate slide at crates/matching-opencv/src/lib.rs:297-311;
 *                         rating = number of inlier flags; gates lib.rs:329-333.
 *   arithmetic          : OpenCV calib3d (third party, not under /root/reference): cv::estimateAffinePartial2D ->
 *                         RANSACPointSetRegistrator::run with AffinePartial2DEstimatorCallback (modelPoints = 2), cv::RNG
 *                         seeded with (uint64)-1.  The returned inlier mask is the RANSAC best mask (the LM refinement only
 *                         touches the matrix), so the rating needs the RANSAC loop only.
 *   parity pin          : tests/test_oracle_ransac.py compares masks and counts with cv2.estimateAffinePartial2D on seeded
 *                         correspondence sets (golden vectors in tests/golden/ransac.npz) and live.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t s; } cvrng_t;
static uint32_t rng_next(cvrng_t* r) {
    r->s = (uint64_t)(uint32_t)r->s * 4164903690u + (uint32_t)(r->s >> 32);
    return (uint32_t)r->s;
}
static int rng_uniform(cvrng_t* r, int a, int b) { return a == b ? a : (int)(rng_next(r) % (uint32_t)(b - a) + a); }

static int cv_round_d(double v) { return (int)lrint(v); }

/* RANSACUpdateNumIters (calib3d/ptsetreg.cpp) */
int ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
    p = p > 0. ? p : 0.; p = p < 1. ? p : 1.;
    ep = ep > 0. ? ep : 0.; ep = ep < 1. ? ep : 1.;
    double num = 1. - p > DBL_MIN ? 1. - p : DBL_MIN;
    double denom = 1. - pow(1. - ep, model_points);
    if (denom < DBL_MIN) return 0;
    num = log(num);
    denom = log(denom);
    return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : cv_round_d(num / denom);
}

/* the i-th RANSAC sample (two distinct indices) for a set of `count` correspondences; pure function of (count, iteration) */
void ransac_sample_sequence(int count, int n_iters, int32_t* idx_pairs /* n_iters x 2 */) {
    cvrng_t rng = {0xFFFFFFFFFFFFFFFFull};
    for (int it = 0; it < n_iters; ++it) {
        int i0 = rng_uniform(&rng, 0, count), i1;
        do { i1 = rng_uniform(&rng, 0, count); } while (i1 == i0);
        idx_pairs[2 * it] = i0;
        idx_pairs[2 * it + 1] = i1;
    }
}

/* AffinePartial2DEstimatorCallback::runKernel: closed-form similarity through two correspondences (double) */
static void partial_affine_from_2(const float* f0, const float* f1, const float* t0, const float* t1, double* M) {
    double x1 = f0[0], y1 = f0[1], x2 = f1[0], y2 = f1[1];
    double X1 = t0[0], Y1 = t0[1], X2 = t1[0], Y2 = t1[1];
    double d = 1. / ((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2));
    double S0 = d * ((X1 - X2) * (x1 - x2) + (Y1 - Y2) * (y1 - y2));
    double S1 = d * ((Y1 - Y2) * (x1 - x2) - (X1 - X2) * (y1 - y2));
    double S2 = d * ((Y1 - Y2) * (x1 * y2 - x2 * y1) - (X1 * y2 - X2 * y1) * (y1 - y2) - (X1 * x2 - X2 * x1) * (x1 - x2));
    double S3 = d * (-(X1 - X2) * (x1 * y2 - x2 * y1) - (Y1 * x2 - Y2 * x1) * (x1 - x2) - (Y1 * y2 - Y2 * y1) * (y1 - y2));
    M[0] = M[4] = S0; M[1] = -S1; M[2] = S2; M[3] = S1; M[5] = S3;
}

/* Affine2DEstimatorCallback::computeError + findInliers: fp32, err <= thr^2.  fma != 0: contracted evaluation. */
static int count_inliers(const float* from, const float* to, int n, const double* M, float t, int fma, uint8_t* mask) {
    float F0 = (float)M[0], F1 = (float)M[1], F2 = (float)M[2], F3 = (float)M[3], F4 = (float)M[4], F5 = (float)M[5];
    int nz = 0;
    for (int i = 0; i < n; ++i) {
        float fx = from[2 * i], fy = from[2 * i + 1], tx = to[2 * i], ty = to[2 * i + 1], a, b, e;
        if (fma) {
            a = fmaf(F1, fy, F0 * fx) + F2 - tx;   /* placeholder order, refined against cv2 in the tests */
            b = fmaf(F4, fy, F3 * fx) + F5 - ty;
            e = fmaf(b, b, a * a);
        } else {
            a = F0 * fx + F1 * fy + F2 - tx;
            b = F3 * fx + F4 * fy + F5 - ty;
            e = a * a + b * b;
        }
        int f = e <= t;
        if (mask) mask[i] = (uint8_t)f;
        nz += f;
    }
    return nz;
}

/* cv::estimateAffinePartial2D(..., RANSAC, thr, max_iters, confidence, refineIters) restricted to what the reference uses:
 * returns the number of inliers (the "rating" of lib.rs:310), fills mask (n) and the RANSAC model (6 doubles, unrefined).
 * n < 2 -> 0 inliers, mask zero.  n == 2 -> all inliers. */
int ransac_affine_partial(const float* from, const float* to, int n, double thr, int max_iters, double confidence, int fma,
                          uint8_t* mask, double* model, int* iters_run) {
    if (mask) memset(mask, 0, (size_t)(n > 0 ? n : 0));
    if (iters_run) *iters_run = 0;
    if (n < 2) return 0;
    double M[6], best[6] = {0, 0, 0, 0, 0, 0};
    if (n == 2) {
        partial_affine_from_2(from, from + 2, to, to + 2, best);
        if (mask) memset(mask, 1, 2);
        if (model) memcpy(model, best, sizeof best);
        return 2;
    }
    uint8_t* cur = (uint8_t*)malloc((size_t)n);
    float t = (float)(thr * thr);
    int niters = max_iters > 1 ? max_iters : 1, max_good = 0, it;
    cvrng_t rng = {0xFFFFFFFFFFFFFFFFull};
    for (it = 0; it < niters; ++it) {
        int i0 = rng_uniform(&rng, 0, n), i1;
        do { i1 = rng_uniform(&rng, 0, n); } while (i1 == i0);
        partial_affine_from_2(from + 2 * i0, from + 2 * i1, to + 2 * i0, to + 2 * i1, M);
        int good = count_inliers(from, to, n, M, t, fma, cur);
        if (good > (max_good > 1 ? max_good : 1)) {
            max_good = good;
            memcpy(best, M, sizeof M);
            if (mask) memcpy(mask, cur, (size_t)n);
            niters = ransac_update_num_iters(confidence, (double)(n - good) / n, 2, niters);
        }
    }
    free(cur);
    if (iters_run) *iters_run = it;
    if (model) memcpy(model, best, sizeof best);
    return max_good;
}
