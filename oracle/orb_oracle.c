/*
 * orb_oracle.c -- CPU restatement (plain C) of the ORB feature extractor exactly as the
 * reference configures it.  TEST INFRASTRUCTURE ONLY: imported by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline leg as the *checker*; never linked
 * into or called from the product path (slideo_b200/).
 *
 * What it restates
 *   reference call site : crates/matching-opencv/src/feature_extractor.rs:12-46
 *                         ORB::create(2000, 1.2, 8, 62, 0, 2, FAST_SCORE, 62, 20).detectAndCompute
 *   arithmetic          : third-party OpenCV (pinned 4.5.2 by the reference: .github/workflows/ci.yml:18,
 *                         opencv crate 0.52.0: crates/matching-opencv/Cargo.toml:8), which is NOT under
 *                         /root/reference.  The published algorithm (features2d/orb.cpp, fast.cpp,
 *                         imgproc resize INTER_LINEAR_EXACT, sepFilter) is restated here from
 *                         SURVEY.md Appendix A.
 *   parity pin          : tests/test_oracle_orb.py checks every stage and the end-to-end
 *                         keypoints+descriptors against cv2 4.13.0 (the only runnable OpenCV here)
 *                         on seeded images and, when /root/reference is mounted, on the reference's
 *                         fixtures data/matchings/test1 (PNG files); golden anchors live in tests/golden/.
 *                         Version skew 4.5.2 -> 4.13.0 is unavoidable and stated in DESIGN.md.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: every fused multiply-add below is explicit).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORB_MAX_LEVELS 16

/* ---------------------------------------------------------------------------------------------
 * A.1  BGR -> gray   (OpenCV cvtColor BGR2GRAY 8u: fixed point, 15 fractional bits)
 * ------------------------------------------------------------------------------------------- */
void orb_gray_from_bgr(const uint8_t* bgr, int w, int h, int stride, uint8_t* gray) {
    for (int y = 0; y < h; ++y) {
        const uint8_t* p = bgr + (size_t)y * stride;
        uint8_t* g = gray + (size_t)y * w;
        for (int x = 0; x < w; ++x)
            g[x] = (uint8_t)((p[3 * x] * 3735 + p[3 * x + 1] * 19235 + p[3 * x + 2] * 9798 + (1 << 14)) >> 15);
    }
}

/* ---------------------------------------------------------------------------------------------
 * A.2  pyramid geometry
 * ------------------------------------------------------------------------------------------- */
static int cv_round_f(float v) { return (int)lrintf(v); }        /* round-half-even (default FE mode) */
static int cv_round_d(double v) { return (int)lrint(v); }

float orb_level_scale(int level, float scale_factor) { return (float)pow((double)scale_factor, (double)level); }

void orb_level_size(int w, int h, int level, float scale_factor, int* lw, int* lh) {
    float inv = 1.f / orb_level_scale(level, scale_factor);
    *lw = cv_round_f((float)w * inv);
    *lh = cv_round_f((float)h * inv);
}

/* per-level keypoint quota (orb.cpp: nfeatures distributed geometrically over the levels) */
void orb_level_quota(int nfeatures, int nlevels, float scale_factor, int* quota) {
    float factor = (float)(1.0 / scale_factor);
    float nd = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        quota[l] = cv_round_f(nd);
        sum += quota[l];
        nd *= factor;
    }
    quota[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
}

/* ---------------------------------------------------------------------------------------------
 * A.3  resize INTER_LINEAR_EXACT for 8-bit, 8.8 fixed point, horizontal then vertical
 * ------------------------------------------------------------------------------------------- */
static void lin_exact_axis(int src, int dst, int* ofs, int* c1) {
    double scale = (double)src / (double)dst;
    for (int d = 0; d < dst; ++d) {
        double val = scale * (d + 0.5) - 0.5;
        int o = (int)floor(val);
        int f = cv_round_d((val - o) * 256.0);
        if (o < 0) { o = 0; f = 0; }
        else if (o >= src - 1) { o = src - 1; f = 0; }
        ofs[d] = o;
        c1[d] = f;
    }
}

void orb_resize_linear_exact(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) {
    int* xo = (int*)malloc(sizeof(int) * dw * 2);
    int* xc = xo + dw;
    int* yo = (int*)malloc(sizeof(int) * dh * 2);
    int* yc = yo + dh;
    lin_exact_axis(sw, dw, xo, xc);
    lin_exact_axis(sh, dh, yo, yc);
    uint16_t* r0 = (uint16_t*)malloc(sizeof(uint16_t) * dw * 2);
    uint16_t* r1 = r0 + dw;
    for (int y = 0; y < dh; ++y) {
        int o = yo[y], f = yc[y];
        const uint8_t* s0 = src + (size_t)o * sw;
        const uint8_t* s1 = src + (size_t)(f ? o + 1 : o) * sw;
        for (int x = 0; x < dw; ++x) {
            int ox = xo[x], fx = xc[x];
            int ox1 = fx ? ox + 1 : ox;
            r0[x] = (uint16_t)(s0[ox] * (256 - fx) + s0[ox1] * fx);
            r1[x] = (uint16_t)(s1[ox] * (256 - fx) + s1[ox1] * fx);
        }
        uint8_t* d = dst + (size_t)y * dw;
        for (int x = 0; x < dw; ++x) {
            uint32_t v = ((uint32_t)r0[x] * (uint32_t)(256 - f) + (uint32_t)r1[x] * (uint32_t)f + (1u << 15)) >> 16;
            d[x] = (uint8_t)(v > 255 ? 255 : v);
        }
    }
    free(xo); free(yo); free(r0);
}

/* ---------------------------------------------------------------------------------------------
 * A.4  FAST-9/16 score map.  score(x,y) = 0 for non-corners, else s-1 where
 *      s = max over the 16 arcs of 9 contiguous circle pixels of max(min d, min -d), corner iff s > thr
 * ------------------------------------------------------------------------------------------- */
static const int FAST_DX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int FAST_DY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

void orb_fast_score_map(const uint8_t* img, int w, int h, int thr, uint8_t* score) {
    memset(score, 0, (size_t)w * h);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            int c = img[(size_t)y * w + x];
            int d[25];
            for (int k = 0; k < 16; ++k) d[k] = c - img[(size_t)(y + FAST_DY[k]) * w + x + FAST_DX[k]];
            for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
            int best = 0;
            for (int k = 0; k < 16; ++k) {
                int mn = d[k], mx = d[k];
                for (int j = 1; j < 9; ++j) {
                    if (d[k + j] < mn) mn = d[k + j];
                    if (d[k + j] > mx) mx = d[k + j];
                }
                if (mn > best) best = mn;          /* all nine darker than centre by >= mn */
                if (-mx > best) best = -mx;        /* all nine brighter */
            }
            if (best > thr) score[(size_t)y * w + x] = (uint8_t)(best - 1);
        }
}

/* 3x3 non-max suppression (strict >), raster order output. Returns count (may exceed cap: truncated) */
int orb_fast_nms(const uint8_t* score, int w, int h, int* xs, int* ys, int* sc, int cap) {
    int n = 0;
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            int s = score[(size_t)y * w + x];
            if (!s) continue;
            const uint8_t* p = score + (size_t)y * w + x;
            if (s > p[-1] && s > p[1] && s > p[-w - 1] && s > p[-w] && s > p[-w + 1] && s > p[w - 1] && s > p[w] &&
                s > p[w + 1]) {
                if (n < cap) { xs[n] = x; ys[n] = y; sc[n] = s; }
                ++n;
            }
        }
    return n;
}

/* ---------------------------------------------------------------------------------------------
 * A.6  orientation: intensity centroid over the radius-15*... disc (half patch = 31), fastAtan2
 * ------------------------------------------------------------------------------------------- */
static void orb_umax(int half, int* umax) {
    /* orb.cpp: umax[v] = cvRound(sqrt(half^2 - v^2)) for v <= vmax, then symmetric fix-up */
    int vmax = (int)floor(half * sqrt(2.0) / 2 + 1);
    int vmin = (int)ceil(half * sqrt(2.0) / 2);
    for (int v = 0; v <= vmax; ++v) umax[v] = cv_round_d(sqrt((double)half * half - (double)v * v));
    for (int v = half, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
}

float orb_fast_atan2(float y, float x) {
    const float R = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * R, p3 = -0.3258083974640975f * R, p5 = 0.1555786518463281f * R,
                p7 = -0.04432655554792128f * R;
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)2.2204460492503131e-16);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)2.2204460492503131e-16);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

float orb_ic_angle(const uint8_t* img, int w, int x, int y, int half, const int* umax) {
    int m01 = 0, m10 = 0;
    const uint8_t* c = img + (size_t)y * w + x;
    for (int u = -half; u <= half; ++u) m10 += u * c[u];
    for (int v = 1; v <= half; ++v) {
        int vsum = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int lo = c[u + v * w], hi = c[u - v * w];
            vsum += lo - hi;
            m10 += u * (lo + hi);
        }
        m01 += v * vsum;
    }
    return orb_fast_atan2((float)m01, (float)m10);
}

/* ---------------------------------------------------------------------------------------------
 * A.7  blur: separable 7-tap Gaussian sigma 2, float path with FMA, BORDER_REFLECT_101
 * ------------------------------------------------------------------------------------------- */
static const uint32_t GAUSS7_BITS[4] = {0x3d8fafb1u, 0x3e06387eu, 0x3e434a39u, 0x3e5d4ae0u};
static float gk(int i) {
    float f;
    uint32_t b = GAUSS7_BITS[i < 4 ? i : 6 - i];
    memcpy(&f, &b, 4);
    return f;
}
static int reflect101(int p, int n) {
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * (n - 1) - p;
    }
    return p;
}

void orb_blur7(const uint8_t* src, int w, int h, uint8_t* dst) {
    float* rows = (float*)malloc(sizeof(float) * (size_t)w * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = src + (size_t)y * w;
        float* r = rows + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            float acc = gk(0) * (float)s[reflect101(x - 3, w)];
            for (int i = 1; i < 7; ++i) acc = fmaf(gk(i), (float)s[reflect101(x - 3 + i, w)], acc);
            r[x] = acc;
        }
    }
    for (int y = 0; y < h; ++y) {
        const float* r3 = rows + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            float acc = gk(3) * r3[x];
            for (int j = 1; j <= 3; ++j) {
                float a = rows[(size_t)reflect101(y + j, h) * w + x];
                float b = rows[(size_t)reflect101(y - j, h) * w + x];
                acc = fmaf(gk(3 + j), a + b, acc);
            }
            int v = cv_round_f(acc);
            dst[(size_t)y * w + x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    }
    free(rows);
}

/* ---------------------------------------------------------------------------------------------
 * A.8  descriptor: 512 sample points from cv::RNG(0x34985739) in [-31,31]^2 (patchSize != 31 branch)
 * ------------------------------------------------------------------------------------------- */
void orb_pattern(int half, int npoints, int* xy /* npoints*2, x first */) {
    uint64_t s = 0x34985739u;
    for (int i = 0; i < npoints * 2; ++i) {
        s = (uint64_t)(uint32_t)s * 4164903690u + (s >> 32);
        uint32_t r = (uint32_t)s;
        xy[i] = -half + (int)(r % (uint32_t)(2 * half + 1));
    }
}

void orb_describe(const uint8_t* blurred, int w, int h, int cx, int cy, float angle_deg, const int* pat,
                  uint8_t* desc /* 32 */) {
    float th = angle_deg * (float)(3.14159265358979323846 / 180.0);
    float a = (float)cos((double)th), b = (float)sin((double)th);
    for (int n = 0; n < 32; ++n) {
        int byte = 0;
        for (int j = 0; j < 8; ++j) {
            int v[2];
            for (int t = 0; t < 2; ++t) {
                int i = 16 * n + 2 * j + t;
                float px = (float)pat[2 * i], py = (float)pat[2 * i + 1];
                int ix = cv_round_f(px * a - py * b);
                int iy = cv_round_f(px * b + py * a);
                v[t] = blurred[(size_t)reflect101(cy + iy, h) * w + reflect101(cx + ix, w)];
            }
            byte |= (v[0] < v[1]) << j;
        }
        desc[n] = (uint8_t)byte;
    }
}

/* ---------------------------------------------------------------------------------------------
 * Full detectAndCompute on an 8-bit gray image, canonical output order (octave, y, x)  (A.9).
 *   kp_i   : n x 4 int32  {x_level, y_level, octave, score}
 *   kp_f   : n x 4 float  {pt.x, pt.y (level-0 coordinates), size, angle_deg}
 *   desc   : n x 32 uint8
 * Returns the number of keypoints (<= cap; -1 if cap was too small).
 * ------------------------------------------------------------------------------------------- */
typedef struct { int x, y, s; } cand_t;

int orb_detect_and_compute(const uint8_t* gray, int w, int h, int nfeatures, float scale_factor, int nlevels,
                           int edge_threshold, int patch_size, int fast_threshold, int32_t* kp_i, float* kp_f,
                           uint8_t* desc, int cap) {
    if (nlevels > ORB_MAX_LEVELS) return -2;
    int half = patch_size / 2;
    int umax[64];
    orb_umax(half, umax);
    int pat[1024];
    orb_pattern(half, 512, pat);
    int quota[ORB_MAX_LEVELS];
    orb_level_quota(nfeatures, nlevels, scale_factor, quota);

    uint8_t* prev = NULL;
    int pw = 0, ph = 0, n = 0;
    for (int l = 0; l < nlevels; ++l) {
        int lw, lh;
        orb_level_size(w, h, l, scale_factor, &lw, &lh);
        uint8_t* img = (uint8_t*)malloc((size_t)lw * lh);
        if (l == 0) memcpy(img, gray, (size_t)w * h);
        else orb_resize_linear_exact(prev, pw, ph, img, lw, lh);
        free(prev);
        prev = img; pw = lw; ph = lh;

        uint8_t* score = (uint8_t*)malloc((size_t)lw * lh);
        orb_fast_score_map(img, lw, lh, fast_threshold, score);
        int ccap = lw * lh / 4 + 16;
        int* xs = (int*)malloc(sizeof(int) * 3 * ccap);
        int *ys = xs + ccap, *sc = ys + ccap;
        int nc = orb_fast_nms(score, lw, lh, xs, ys, sc, ccap);
        free(score);
        /* runByImageBorder(edge_threshold) */
        int m = 0, hist[256];
        memset(hist, 0, sizeof hist);
        for (int i = 0; i < nc; ++i)
            if (xs[i] >= edge_threshold && xs[i] < lw - edge_threshold && ys[i] >= edge_threshold &&
                ys[i] < lh - edge_threshold) {
                xs[m] = xs[i]; ys[m] = ys[i]; sc[m] = sc[i];
                hist[sc[m]]++;
                ++m;
            }
        /* retainBest(quota): keep everything >= the quota-th largest score (ties retained) */
        int thr = 0;
        if (m > quota[l]) {
            if (quota[l] == 0) thr = 1 << 30;
            else {
                int acc = 0;
                for (thr = 255; thr > 0; --thr) { acc += hist[thr]; if (acc >= quota[l]) break; }
            }
        }
        float scale = orb_level_scale(l, scale_factor);
        float inv = 1.f / scale;
        uint8_t* blurred = NULL;
        for (int i = 0; i < m; ++i) {       /* raster order == canonical (y, x) order within the level */
            if (sc[i] < thr) continue;
            if (n >= cap) { free(xs); free(prev); free(blurred); return -1; }
            if (!blurred) { blurred = (uint8_t*)malloc((size_t)lw * lh); orb_blur7(img, lw, lh, blurred); }
            float ang = orb_ic_angle(img, lw, xs[i], ys[i], half, umax);
            float ptx = (float)xs[i] * scale, pty = (float)ys[i] * scale;
            kp_i[4 * n] = xs[i]; kp_i[4 * n + 1] = ys[i]; kp_i[4 * n + 2] = l; kp_i[4 * n + 3] = sc[i];
            kp_f[4 * n] = ptx; kp_f[4 * n + 1] = pty; kp_f[4 * n + 2] = (float)patch_size * scale; kp_f[4 * n + 3] = ang;
            orb_describe(blurred, lw, lh, cv_round_f(ptx * inv), cv_round_f(pty * inv), ang, pat, desc + (size_t)32 * n);
            ++n;
        }
        free(blurred);
        free(xs);
    }
    free(prev);
    return n;
}
