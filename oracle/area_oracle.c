/*
 * area_oracle.c -- CPU restatement (plain C) of the changed-frame prefilter of the reference.
 * TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / reference arm).
 *
 * What it restates
 *   reference call sites: crates/matching-opencv/src/image_utils.rs:8-19  to_small_image  (resize(..., INTER_AREA) to ~300x400 area)
 *                         crates/matching-opencv/src/image_utils.rs:21-27 compute_similarity (1 - norm2(a, b, NORM_L2) / sqrt(255^2*3*p))
 *                         crates/matching-opencv/src/video_capture.rs:86-102 MarkSimilarIter (changed iff similarity < 0.98)
 *   arithmetic          : OpenCV imgproc resize INTER_AREA for a non-integer scale (third party, not under /root/reference):
 *                         computeResizeAreaTab + ResizeArea_Invoker<uchar, float>: per destination pixel, float accumulation of
 *                         source * alpha along x in table order, then beta-weighted accumulation over the source rows, then
 *                         saturate_cast<uchar> (round half to even).
 *   parity pin          : tests/test_oracle_area.py against cv2.resize(INTER_AREA) / cv2.norm (golden CRCs + live).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int si, di; float alpha; } area_tab_t;

/* computeResizeAreaTab (imgproc/resize.cpp), cn folded in by the caller */
static int area_tab(int ssize, int dsize, double scale, area_tab_t* tab) {
    int k = 0;
    for (int dx = 0; dx < dsize; ++dx) {
        double fsx1 = dx * scale, fsx2 = fsx1 + scale;
        double cell = scale < ssize - fsx1 ? scale : ssize - fsx1;
        int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
        sx2 = sx2 < ssize - 1 ? sx2 : ssize - 1;
        sx1 = sx1 < sx2 ? sx1 : sx2;
        if (sx1 - fsx1 > 1e-3) { tab[k].di = dx; tab[k].si = sx1 - 1; tab[k++].alpha = (float)((sx1 - fsx1) / cell); }
        for (int sx = sx1; sx < sx2; ++sx) { tab[k].di = dx; tab[k].si = sx; tab[k++].alpha = (float)(1.0 / cell); }
        if (fsx2 - sx2 > 1e-3) {
            double a = fsx2 - sx2; a = a < 1.0 ? a : 1.0; a = a < cell ? a : cell;
            tab[k].di = dx; tab[k].si = sx2; tab[k++].alpha = (float)(a / cell);
        }
    }
    return k;
}

void area_small_size(int w, int h, int* sw, int* sh) {   /* image_utils.rs:10-16, all in f32, `as i32` truncates */
    float factor = sqrtf((float)(300 * 400) / (float)(w * h));
    *sw = (int)((float)w * factor);
    *sh = (int)((float)h * factor);
}

/* cv::resize(src, dst, dsize, 0, 0, INTER_AREA), 8-bit, cn interleaved channels, shrinking by a non-integer factor.
 * fma != 0 evaluates the accumulations contracted (to identify which form the OpenCV build uses). */
void area_resize_u8(const uint8_t* src, int sw, int sh, int sstride, int cn, uint8_t* dst, int dw, int dh, int fma) {
    double scale_x = (double)sw / dw, scale_y = (double)sh / dh;
    area_tab_t* xtab = (area_tab_t*)malloc(sizeof(area_tab_t) * (size_t)(sw + dw) * 2);
    area_tab_t* ytab = (area_tab_t*)malloc(sizeof(area_tab_t) * (size_t)(sh + dh) * 2);
    int xn = area_tab(sw, dw, scale_x, xtab), yn = area_tab(sh, dh, scale_y, ytab);
    int n = dw * cn;
    float* buf = (float*)malloc(sizeof(float) * (size_t)n * 2);
    float* sum = buf + n;
    int prev_dy = ytab[0].di;
    for (int i = 0; i < n; ++i) sum[i] = 0.f;
    for (int j = 0; j < yn; ++j) {
        float beta = ytab[j].alpha;
        int dy = ytab[j].di, sy = ytab[j].si;
        const uint8_t* S = src + (size_t)sy * sstride;
        for (int i = 0; i < n; ++i) buf[i] = 0.f;
        for (int k = 0; k < xn; ++k) {
            float alpha = xtab[k].alpha;
            for (int c = 0; c < cn; ++c) {
                float* b = &buf[xtab[k].di * cn + c];
                float s = (float)S[xtab[k].si * cn + c];
                *b = fma ? fmaf(s, alpha, *b) : *b + s * alpha;
            }
        }
        if (dy != prev_dy) {
            uint8_t* D = dst + (size_t)prev_dy * n;
            for (int i = 0; i < n; ++i) {
                long v = lrintf(sum[i]);
                D[i] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
                sum[i] = beta * buf[i];
            }
            prev_dy = dy;
        } else {
            for (int i = 0; i < n; ++i) sum[i] = fma ? fmaf(beta, buf[i], sum[i]) : sum[i] + beta * buf[i];
        }
    }
    uint8_t* D = dst + (size_t)prev_dy * n;
    for (int i = 0; i < n; ++i) {
        long v = lrintf(sum[i]);
        D[i] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
    }
    free(buf); free(xtab); free(ytab);
}

/* compute_similarity (image_utils.rs:21-27): f32 result from the exact integer sum of squared differences */
float area_similarity(const uint8_t* a, const uint8_t* b, int w, int h, int cn) {
    uint64_t ss = 0;
    size_t n = (size_t)w * h * cn;
    for (size_t i = 0; i < n; ++i) { int d = (int)a[i] - (int)b[i]; ss += (uint64_t)(d * d); }
    double error_l2 = sqrt((double)ss);
    int p = h * w;
    float max_error = sqrtf((255.0f * 255.0f * 3.0f) * (float)p);
    return 1.0f - (float)error_l2 / max_error;
}
