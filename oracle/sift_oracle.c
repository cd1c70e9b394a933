/*
 * sift_oracle.c -- CPU restatement (plain C) of OpenCV's SIFT (cv::SIFT::create() defaults) for the north_star's
 * "SIFT DoG pyramid + 128-d float descriptor variant".  TEST INFRASTRUCTURE ONLY: imported by tests/ and by
 * bench.py's cpu_baseline leg as the *checker*; never linked into or called from the product path (slideo_b200/).
 *
 * What it restates
 *   reference call site : none -- SIFT does not exist in hediet/slideo; it would sit where
 *                         crates/matching-opencv/src/feature_extractor.rs:29-46 calls ORB::detectAndCompute.
 *   arithmetic          : third-party OpenCV features2d/src/sift.dispatch.cpp + sift.simd.hpp (4.x), with
 *                         nfeatures 0, nOctaveLayers 3, contrastThreshold 0.04, edgeThreshold 10, sigma 1.6,
 *                         descriptorType CV_32F; plus imgproc GaussianBlur (float sepFilter path) and
 *                         core hal::exp32f / fastAtan2 / magnitude32f as the AVX2 build evaluates them (FMA).
 *   parity pin          : tests/test_oracle_sift.py -- leaf functions against cv2 4.13.0 bit for bit
 *                         (getGaussianKernel, GaussianBlur on float images, cv2.exp, cv2.phase, cv2.magnitude, the 2x
 *                         INTER_LINEAR upsample), end to end against cv2.SIFT_create().detectAndCompute on seeded
 *                         images + committed golden vectors, at the tolerance cv2 shows against ITSELF between its
 *                         own code paths (setUseOptimized on/off: 2118 of 2120 keypoints, descriptors +-1 LSB on
 *                         0.008 % of the elements) -- OpenCV's compiler-contracted float code has no single
 *                         bit-exact definition.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: every fused multiply-add below is explicit).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SIFT_DESCR_WIDTH 4
#define SIFT_DESCR_HIST_BINS 8
#define SIFT_INIT_SIGMA 0.5f
#define SIFT_IMG_BORDER 5
#define SIFT_MAX_INTERP_STEPS 5
#define SIFT_ORI_HIST_BINS 36
#define SIFT_ORI_SIG_FCTR 1.5f
#define SIFT_ORI_RADIUS 4.5f /* 3 * SIFT_ORI_SIG_FCTR */
#define SIFT_ORI_PEAK_RATIO 0.8f
#define SIFT_DESCR_SCL_FCTR 3.f
#define SIFT_DESCR_MAG_THR 0.2f
#define SIFT_INT_DESCR_FCTR 512.f
#define SIFT_N_OCTAVE_LAYERS 3
#define SIFT_CONTRAST_THRESHOLD 0.04
#define SIFT_EDGE_THRESHOLD 10.0
#define SIFT_SIGMA 1.6
#define SIFT_MAX_OCTAVES 16

static int cv_round_f(float v) { return (int)lrintf(v); }
static int cv_round_d(double v) { return (int)lrint(v); }
static int cv_floor_f(float v) { return (int)floorf(v); }

/* ---------------------------------------------------------------------------------------------
 * S.1  Gaussian taps: getGaussianKernel(n, sigma, CV_32F) -> getGaussianKernelBitExact (softdouble == IEEE double
 *      for + * /; exp() is the only libm call) ; n = cvRound(sigma*4*2 + 1) | 1 for CV_32F images.
 * ------------------------------------------------------------------------------------------- */
int sift_gauss_ksize(double sigma) { return cv_round_d(sigma * 4 * 2 + 1) | 1; }

void sift_gauss_taps(double sigma, int n, float* taps) {
    double* v = (double*)malloc(sizeof(double) * (size_t)n);
    double scale2x = -0.125 / (sigma * sigma); /* x below is doubled */
    int n2 = (n - 1) / 2;
    double sum = 0;
    for (int i = 0, x = 1 - n; i < n2; ++i, x += 2) {
        v[i] = exp((double)(x * x) * scale2x);
        sum += v[i];
    }
    sum *= 2;
    sum += 1;
    double mul1 = 1.0 / sum, sum2 = 0;
    double* r = (double*)malloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < n2; ++i) {
        double t = v[i] * mul1;
        r[i] = t;
        r[n - 1 - i] = t;
        sum2 += t;
    }
    sum2 *= 2;
    r[n2] = 1.0 * mul1;
    sum2 += r[n2];
    r[n2] += 1.0 - sum2;
    for (int i = 0; i < n; ++i) taps[i] = (float)r[i];
    free(v);
    free(r);
}

/* ---------------------------------------------------------------------------------------------
 * S.2  GaussianBlur on a float image: sepFilter2D, BORDER_REFLECT_101.
 *      rows   : RowVec_32f       s = x0*k0; s = fma(x_i, k_i, s), i ascending
 *      columns: SymmColumnVec_32f s = c*k0;  s = fma(r[+j] + r[-j], k_j, s), j ascending
 * ------------------------------------------------------------------------------------------- */
static int border101(int p, int len) {
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p;
        else p = 2 * len - 2 - p;
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

static float mul_add_unfused(float a, float b, float c) { /* the scalar remainder loops of filter.simd.hpp: mul, then add */
    volatile float m = a * b;
    return c + m;
}

void sift_gauss_blur(const float* src, int w, int h, const float* taps, int n, float* dst) {
    int r = n / 2;
    /* RowVec_32f vectorises down to 4 lanes, SymmColumnVec_32f to 8 (AVX2); the remaining columns run OpenCV's scalar
     * loops, which are NOT contracted (measured against cv2: tests/test_oracle_sift.py::test_blur_equals_cv2) */
    const int row_tail = w - w % 4, col_tail = w - w % 8;
    float* rows = (float*)malloc(sizeof(float) * (size_t)w * h);
    int* xi = (int*)malloc(sizeof(int) * (size_t)(w + 2 * r));
    for (int x = -r; x < w + r; ++x) xi[x + r] = border101(x, w);
    for (int y = 0; y < h; ++y) {
        const float* s = src + (size_t)y * w;
        float* d = rows + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            const int* ix = xi + x;
            float acc = s[ix[0]] * taps[0];
            if (x < row_tail)
                for (int i = 1; i < n; ++i) acc = fmaf(s[ix[i]], taps[i], acc);
            else
                for (int i = 1; i < n; ++i) acc = mul_add_unfused(s[ix[i]], taps[i], acc);
            d[x] = acc;
        }
    }
    for (int y = 0; y < h; ++y) {
        const float* c = rows + (size_t)y * w;
        float* d = dst + (size_t)y * w;
        for (int x = 0; x < w; ++x) d[x] = c[x] * taps[r];
        for (int j = 1; j <= r; ++j) {
            const float* a = rows + (size_t)border101(y + j, h) * w;
            const float* b = rows + (size_t)border101(y - j, h) * w;
            float k = taps[r + j];
            for (int x = 0; x < col_tail; ++x) d[x] = fmaf(a[x] + b[x], k, d[x]);
            for (int x = col_tail; x < w; ++x) d[x] = mul_add_unfused(a[x] + b[x], k, d[x]);
        }
    }
    free(rows);
    free(xi);
}

/* ---------------------------------------------------------------------------------------------
 * S.3  createInitialImage: u8 gray -> float, resize x2 INTER_LINEAR (weights 0.25/0.75: every product and sum of
 *      8-bit inputs is exact in fp32, so the op order is immaterial), blur with sig_diff.
 * ------------------------------------------------------------------------------------------- */
void sift_upsample2(const uint8_t* gray, int w, int h, float* dst) {
    int W = 2 * w, H = 2 * h;
    int* xo = (int*)malloc(sizeof(int) * (size_t)W);
    float* xa = (float*)malloc(sizeof(float) * (size_t)W);
    for (int dx = 0; dx < W; ++dx) {
        float fx = (float)((dx + 0.5) * 0.5 - 0.5);
        int sx = cv_floor_f(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= w - 1) { fx = 0; sx = w - 1; }
        xo[dx] = sx;
        xa[dx] = fx;
    }
    float* hrow0 = (float*)malloc(sizeof(float) * (size_t)W);
    float* hrow1 = (float*)malloc(sizeof(float) * (size_t)W);
    for (int dy = 0; dy < H; ++dy) {
        float fy = (float)((dy + 0.5) * 0.5 - 0.5);
        int sy = cv_floor_f(fy);
        fy -= sy;
        if (sy < 0) { fy = 0; sy = 0; }
        if (sy >= h - 1) { fy = 0; sy = h - 1; }
        int sy1 = sy + 1 < h ? sy + 1 : h - 1;
        const uint8_t* s0 = gray + (size_t)sy * w;
        const uint8_t* s1 = gray + (size_t)sy1 * w;
        for (int dx = 0; dx < W; ++dx) {
            int sx = xo[dx], sx1 = sx + 1 < w ? sx + 1 : w - 1;
            float a1 = xa[dx], a0 = 1.f - a1;
            hrow0[dx] = (float)s0[sx] * a0 + (float)s0[sx1] * a1;
            hrow1[dx] = (float)s1[sx] * a0 + (float)s1[sx1] * a1;
        }
        float b1 = fy, b0 = 1.f - fy;
        float* d = dst + (size_t)dy * W;
        for (int dx = 0; dx < W; ++dx) d[dx] = hrow0[dx] * b0 + hrow1[dx] * b1;
    }
    free(xo); free(xa); free(hrow0); free(hrow1);
}

/* ---------------------------------------------------------------------------------------------
 * S.4  hal leaf functions as the SIMD (FMA) build evaluates them
 * ------------------------------------------------------------------------------------------- */
/* cv::hal::exp32f: 64-entry table of 2^(j/64)*A0 and a cubic in the remainder */
#define EXPTAB_SCALE 6
#define EXPTAB_MASK ((1 << EXPTAB_SCALE) - 1)
#define EXPPOLY_32F_A0 .9670371139572337719125840413672004409288e-2
static float g_exptab[EXPTAB_MASK + 1];
static int g_exptab_ready = 0;
float sift_exp32f(float x) {
    if (!g_exptab_ready) {
        for (int j = 0; j <= EXPTAB_MASK; ++j) g_exptab[j] = (float)(exp2((double)j / (1 << EXPTAB_SCALE)) * EXPPOLY_32F_A0);
        g_exptab_ready = 1;
    }
    const double exp_prescale = 1.4426950408889634073599246810019 * (1 << EXPTAB_SCALE);
    const double exp_postscale = 1. / (1 << EXPTAB_SCALE);
    const double exp_max_val = 3000. * (1 << EXPTAB_SCALE);
    const float A4 = (float)(1.000000000000002438532970795181890933776 / EXPPOLY_32F_A0),
                A3 = (float)(.6931471805521448196800669615864773144641 / EXPPOLY_32F_A0),
                A2 = (float)(.2402265109513301490103372422686535526573 / EXPPOLY_32F_A0),
                A1 = (float)(.5550339366753125211915322047004666939128e-1 / EXPPOLY_32F_A0);
    const float minval = (float)(-exp_max_val / exp_prescale), maxval = (float)(exp_max_val / exp_prescale);
    float x0 = fminf(fmaxf(x, minval), maxval);
    x0 *= (float)exp_prescale;
    int xi = cv_round_f(x0);
    x0 = (x0 - (float)xi) * (float)exp_postscale;
    int t = (xi >> EXPTAB_SCALE) + 127;
    t = t < 0 ? 0 : t > 255 ? 255 : t;
    uint32_t bits = (uint32_t)t << 23;
    float p2;
    memcpy(&p2, &bits, 4);
    float y = g_exptab[xi & EXPTAB_MASK] * p2;
    float z = x0 + A1;
    z = fmaf(z, x0, A2);
    z = fmaf(z, x0, A3);
    z = fmaf(z, x0, A4);
    return z * y;
}

/* cv::hal::fastAtan2 (array version, v_atan_f32, degrees) */
float sift_fast_atan2(float y, float x) {
    const float R = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * R, p3 = -0.3258083974640975f * R, p5 = 0.1555786518463281f * R,
                p7 = -0.04432655554792128f * R;
    float ax = fabsf(x), ay = fabsf(y);
    float c = fminf(ax, ay) / (fmaxf(ax, ay) + (float)2.2204460492503131e-16);
    float cc = c * c;
    float a = fmaf(fmaf(fmaf(cc, p7, p5), cc, p3), cc, p1) * c;
    if (!(ax >= ay)) a = 90.f - a;
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* cv::hal::magnitude32f */
float sift_magnitude(float x, float y) { return sqrtf(fmaf(x, x, y * y)); }

/* ---------------------------------------------------------------------------------------------
 * S.5  pyramid
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int n_octaves;
    int w[SIFT_MAX_OCTAVES], h[SIFT_MAX_OCTAVES];
    float* gauss[SIFT_MAX_OCTAVES][SIFT_N_OCTAVE_LAYERS + 3];
    float* dog[SIFT_MAX_OCTAVES][SIFT_N_OCTAVE_LAYERS + 2];
} sift_pyr;

int sift_num_octaves(int w, int h) { /* on the DOUBLED base image, firstOctave = -1 */
    int m = 2 * (w < h ? w : h);
    return cv_round_d(log((double)m) / log(2.) - 2) + 1;
}

void sift_layer_sigmas(double* sig) { /* sig[0..5]; sig[0] is the blur of the initial image (sig_diff) */
    float s = (float)SIFT_SIGMA;
    float d = s * s - SIFT_INIT_SIGMA * SIFT_INIT_SIGMA * 4;
    sig[0] = (double)sqrtf(d > 0.01f ? d : 0.01f);
    double k = pow(2., 1. / SIFT_N_OCTAVE_LAYERS);
    for (int i = 1; i < SIFT_N_OCTAVE_LAYERS + 3; ++i) {
        double sig_prev = pow(k, (double)(i - 1)) * SIFT_SIGMA;
        double sig_total = sig_prev * k;
        sig[i] = sqrt(sig_total * sig_total - sig_prev * sig_prev);
    }
}

static void sift_build(const uint8_t* gray, int w, int h, sift_pyr* P) {
    double sig[SIFT_N_OCTAVE_LAYERS + 3];
    float taps[SIFT_N_OCTAVE_LAYERS + 3][64];
    int nt[SIFT_N_OCTAVE_LAYERS + 3];
    sift_layer_sigmas(sig);
    for (int i = 0; i < SIFT_N_OCTAVE_LAYERS + 3; ++i) {
        nt[i] = sift_gauss_ksize(sig[i]);
        sift_gauss_taps(sig[i], nt[i], taps[i]);
    }
    memset(P, 0, sizeof(*P));
    P->n_octaves = sift_num_octaves(w, h);
    int W = 2 * w, H = 2 * h;
    for (int o = 0; o < P->n_octaves; ++o) {
        P->w[o] = W;
        P->h[o] = H;
        size_t px = (size_t)W * H;
        for (int i = 0; i < SIFT_N_OCTAVE_LAYERS + 3; ++i) P->gauss[o][i] = (float*)malloc(sizeof(float) * (px ? px : 1));
        for (int i = 0; i < SIFT_N_OCTAVE_LAYERS + 2; ++i) P->dog[o][i] = (float*)malloc(sizeof(float) * (px ? px : 1));
        if (o == 0) {
            float* up = (float*)malloc(sizeof(float) * px);
            sift_upsample2(gray, w, h, up);
            sift_gauss_blur(up, W, H, taps[0], nt[0], P->gauss[0][0]);
            free(up);
        } else { /* resize(INTER_NEAREST) of layer nOctaveLayers of the previous octave to (cols/2, rows/2): resizeNN,
                    sx = min(floor(x * ifx), cols - 1) with ifx = 1 / ((double)dst_cols / src_cols) */
            const float* s = P->gauss[o - 1][SIFT_N_OCTAVE_LAYERS];
            int pw = P->w[o - 1], ph = P->h[o - 1];
            double ifx = W ? 1. / ((double)W / pw) : 0, ify = H ? 1. / ((double)H / ph) : 0;
            for (int y = 0; y < H; ++y) {
                int sy = (int)floor(y * ify);
                if (sy > ph - 1) sy = ph - 1;
                for (int x = 0; x < W; ++x) {
                    int sx = (int)floor(x * ifx);
                    if (sx > pw - 1) sx = pw - 1;
                    P->gauss[o][0][(size_t)y * W + x] = s[(size_t)sy * pw + sx];
                }
            }
        }
        for (int i = 1; i < SIFT_N_OCTAVE_LAYERS + 3; ++i) sift_gauss_blur(P->gauss[o][i - 1], W, H, taps[i], nt[i], P->gauss[o][i]);
        for (int i = 0; i < SIFT_N_OCTAVE_LAYERS + 2; ++i)
            for (size_t p = 0; p < px; ++p) P->dog[o][i][p] = P->gauss[o][i + 1][p] - P->gauss[o][i][p];
        W /= 2;
        H /= 2;
    }
}

static void sift_free(sift_pyr* P) {
    for (int o = 0; o < P->n_octaves; ++o) {
        for (int i = 0; i < SIFT_N_OCTAVE_LAYERS + 3; ++i) free(P->gauss[o][i]);
        for (int i = 0; i < SIFT_N_OCTAVE_LAYERS + 2; ++i) free(P->dog[o][i]);
    }
}

/* debug / stage parity: copy one pyramid image (kind 0 gauss, 1 dog) */
int sift_pyramid_image(const uint8_t* gray, int w, int h, int kind, int octave, int layer, float* out, int* ow, int* oh) {
    sift_pyr P;
    sift_build(gray, w, h, &P);
    if (octave >= P.n_octaves) { sift_free(&P); return -1; }
    *ow = P.w[octave];
    *oh = P.h[octave];
    if (out) memcpy(out, kind ? P.dog[octave][layer] : P.gauss[octave][layer], sizeof(float) * (size_t)P.w[octave] * P.h[octave]);
    sift_free(&P);
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * S.6  adjustLocalExtrema: <= 5 Newton steps on the 3-D quadratic, contrast and edge tests
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave;
} sift_kp;

#ifndef SIFT_TAIL_SMOOTH
#define SIFT_TAIL_SMOOTH(a, b, c) fmaf((c), 6.f / 16.f, (a) * (1.f / 16.f) + (b) * (4.f / 16.f))
#endif
#define AT(img, r, c) (img)[(size_t)(r) * W + (c)]

static int sift_adjust(const sift_pyr* P, int octv, int* layer, int* r_, int* c_, sift_kp* kpt) {
    const float img_scale = 1.f / 255;
    const float deriv_scale = img_scale * 0.5f;
    const float second_deriv_scale = img_scale;
    const float cross_deriv_scale = img_scale * 0.25f;
    const int W = P->w[octv], H = P->h[octv];
    float xi = 0, xr = 0, xc = 0, contr = 0;
    int i = 0, r = *r_, c = *c_, l = *layer;
    for (; i < SIFT_MAX_INTERP_STEPS; ++i) {
        const float *img = P->dog[octv][l], *prev = P->dog[octv][l - 1], *next = P->dog[octv][l + 1];
        float dD0 = (AT(img, r, c + 1) - AT(img, r, c - 1)) * deriv_scale;
        float dD1 = (AT(img, r + 1, c) - AT(img, r - 1, c)) * deriv_scale;
        float dD2 = (AT(next, r, c) - AT(prev, r, c)) * deriv_scale;
        float v2 = AT(img, r, c) * 2;
        float dxx = (AT(img, r, c + 1) + AT(img, r, c - 1) - v2) * second_deriv_scale;
        float dyy = (AT(img, r + 1, c) + AT(img, r - 1, c) - v2) * second_deriv_scale;
        float dss = (AT(next, r, c) + AT(prev, r, c) - v2) * second_deriv_scale;
        float dxy = (AT(img, r + 1, c + 1) - AT(img, r + 1, c - 1) - AT(img, r - 1, c + 1) + AT(img, r - 1, c - 1)) * cross_deriv_scale;
        float dxs = (AT(next, r, c + 1) - AT(next, r, c - 1) - AT(prev, r, c + 1) + AT(prev, r, c - 1)) * cross_deriv_scale;
        float dys = (AT(next, r + 1, c) - AT(next, r - 1, c) - AT(prev, r + 1, c) + AT(prev, r - 1, c)) * cross_deriv_scale;
        /* Matx33f H(dxx, dxy, dxs, dxy, dyy, dys, dxs, dys, dss); X = H.solve(dD, DECOMP_LU): Cramer's rule in float */
        float a00 = dxx, a01 = dxy, a02 = dxs, a10 = dxy, a11 = dyy, a12 = dys, a20 = dxs, a21 = dys, a22 = dss;
        /* OpenCV's sift.simd.hpp is compiled per ISA with FMA and GCC's default -ffp-contract=fast: `p*q - r*s` becomes
         * fma(p, q, -(r*s)), `u - p*q` chains become fma too.  FMS(p,q,r,s) below is that form. */
#ifdef VAR_A
#define FMS(p, q, r, s) fmaf(-(r), (s), (p) * (q))
#else
#define FMS(p, q, r, s) fmaf((p), (q), -((r) * (s)))
#endif
#ifdef VAR_B
#define OUT3(p, m1, q, m2, r, m3) fmaf((r), (m3), fmaf(-(q), (m2), (p) * (m1)))
#else
#define OUT3(p, m1, q, m2, r, m3) fmaf((r), (m3), fmaf((p), (m1), -((q) * (m2))))
#endif
        float det = OUT3(a00, FMS(a11, a22, a21, a12), a01, FMS(a10, a22, a20, a12), a02, FMS(a10, a21, a20, a11));
        float X0 = 0, X1 = 0, X2 = 0;
        if (det != 0) {
            float d = 1 / det;
            X0 = d * OUT3(dD0, FMS(a11, a22, a12, a21), a01, FMS(dD1, a22, a12, dD2), a02, FMS(dD1, a21, a11, dD2));
            X1 = d * OUT3(a00, FMS(dD1, a22, a12, dD2), dD0, FMS(a10, a22, a12, a20), a02, FMS(a10, dD2, dD1, a20));
            X2 = d * OUT3(a00, FMS(a11, dD2, dD1, a21), a01, FMS(a10, dD2, dD1, a20), dD0, FMS(a10, a21, a11, a20));
        }
        xi = -X2;
        xr = -X1;
        xc = -X0;
        if (fabsf(xi) < 0.5f && fabsf(xr) < 0.5f && fabsf(xc) < 0.5f) break;
        if (fabsf(xi) > (float)(INT32_MAX / 3) || fabsf(xr) > (float)(INT32_MAX / 3) || fabsf(xc) > (float)(INT32_MAX / 3)) return 0;
        c += cv_round_f(xc);
        r += cv_round_f(xr);
        l += cv_round_f(xi);
        if (l < 1 || l > SIFT_N_OCTAVE_LAYERS || c < SIFT_IMG_BORDER || c >= W - SIFT_IMG_BORDER || r < SIFT_IMG_BORDER ||
            r >= H - SIFT_IMG_BORDER)
            return 0;
    }
    if (i >= SIFT_MAX_INTERP_STEPS) return 0;
    {
        const float *img = P->dog[octv][l], *prev = P->dog[octv][l - 1], *next = P->dog[octv][l + 1];
        float dD0 = (AT(img, r, c + 1) - AT(img, r, c - 1)) * deriv_scale;
        float dD1 = (AT(img, r + 1, c) - AT(img, r - 1, c)) * deriv_scale;
        float dD2 = (AT(next, r, c) - AT(prev, r, c)) * deriv_scale;
        float t = fmaf(dD2, xi, fmaf(dD1, xr, dD0 * xc)); /* Matx::dot: s += a[i]*b[i] */
        contr = fmaf(AT(img, r, c), img_scale, t * 0.5f);
        if (fabsf(contr) * SIFT_N_OCTAVE_LAYERS < (float)SIFT_CONTRAST_THRESHOLD) return 0;
        float v2 = AT(img, r, c) * 2.f;
        float dxx = (AT(img, r, c + 1) + AT(img, r, c - 1) - v2) * second_deriv_scale;
        float dyy = (AT(img, r + 1, c) + AT(img, r - 1, c) - v2) * second_deriv_scale;
        float dxy = (AT(img, r + 1, c + 1) - AT(img, r + 1, c - 1) - AT(img, r - 1, c + 1) + AT(img, r - 1, c - 1)) * cross_deriv_scale;
        float tr = dxx + dyy;
        float det = FMS(dxx, dyy, dxy, dxy);
        const float e = (float)SIFT_EDGE_THRESHOLD;
        if (det <= 0 || tr * tr * e >= (e + 1) * (e + 1) * det) return 0;
    }
    kpt->x = (c + xc) * (1 << octv);
    kpt->y = (r + xr) * (1 << octv);
    kpt->octave = octv + (l << 8) + (cv_round_d((xi + 0.5) * 255) << 16);
    kpt->size = (float)SIFT_SIGMA * exp2f((l + xi) / SIFT_N_OCTAVE_LAYERS) * (1 << octv) * 2;
    kpt->response = fabsf(contr);
    *layer = l;
    *r_ = r;
    *c_ = c;
    return 1;
}

/* ---------------------------------------------------------------------------------------------
 * S.7  calcOrientationHist: 36-bin gradient histogram, accumulated in raster order, smoothed [1 4 6 4 1]/16
 * ------------------------------------------------------------------------------------------- */
static float sift_ori_hist(const float* img, int W, int H, int px, int py, int radius, float sigma, float* hist) {
    const int n = SIFT_ORI_HIST_BINS;
    float expf_scale = -1.f / (2.f * sigma * sigma);
    float tmp[SIFT_ORI_HIST_BINS + 4];
    float* temphist = tmp + 2;
    for (int i = 0; i < n; ++i) temphist[i] = 0.f;
    for (int i = -radius; i <= radius; ++i) {
        int y = py + i;
        if (y <= 0 || y >= H - 1) continue;
        for (int j = -radius; j <= radius; ++j) {
            int x = px + j;
            if (x <= 0 || x >= W - 1) continue;
            float dx = AT(img, y, x + 1) - AT(img, y, x - 1);
            float dy = AT(img, y - 1, x) - AT(img, y + 1, x);
            float wgt = sift_exp32f((float)(i * i + j * j) * expf_scale);
            float ori = sift_fast_atan2(dy, dx);
            float mag = sift_magnitude(dx, dy);
            int bin = cv_round_f((n / 360.f) * ori);
            if (bin >= n) bin -= n;
            if (bin < 0) bin += n;
            temphist[bin] += wgt * mag;
        }
    }
    temphist[-1] = temphist[n - 1];
    temphist[-2] = temphist[n - 2];
    temphist[n] = temphist[0];
    temphist[n + 1] = temphist[1];
    for (int i = 0; i < n; ++i) { /* bins 0..31: the SIMD loop (nested v_fma); bins 32..35: its scalar remainder */
        float a = temphist[i - 2] + temphist[i + 2], b = temphist[i - 1] + temphist[i + 1], c = temphist[i];
        if (i < 32) hist[i] = fmaf(a, 1.f / 16.f, fmaf(b, 4.f / 16.f, c * (6.f / 16.f)));
        else hist[i] = SIFT_TAIL_SMOOTH(a, b, c);
    }
    float maxval = hist[0];
    for (int i = 1; i < n; ++i) maxval = maxval > hist[i] ? maxval : hist[i];
    return maxval;
}

/* ---------------------------------------------------------------------------------------------
 * S.8  findScaleSpaceExtrema: 26-neighbour extrema (ties allowed), refinement, orientation peaks
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    sift_kp* v;
    int n, cap;
} kp_vec;
static void kp_push(kp_vec* K, const sift_kp* k) {
    if (K->n == K->cap) {
        K->cap = K->cap ? 2 * K->cap : 4096;
        K->v = (sift_kp*)realloc(K->v, sizeof(sift_kp) * (size_t)K->cap);
    }
    K->v[K->n++] = *k;
}

static void sift_extrema(const sift_pyr* P, kp_vec* K) {
    const int threshold = (int)floor(0.5 * SIFT_CONTRAST_THRESHOLD / SIFT_N_OCTAVE_LAYERS * 255);
    const int n = SIFT_ORI_HIST_BINS;
    float hist[SIFT_ORI_HIST_BINS];
    for (int o = 0; o < P->n_octaves; ++o) {
        const int W = P->w[o], H = P->h[o];
        for (int i = 1; i <= SIFT_N_OCTAVE_LAYERS; ++i) {
            const float *img = P->dog[o][i], *prev = P->dog[o][i - 1], *next = P->dog[o][i + 1];
            for (int r = SIFT_IMG_BORDER; r < H - SIFT_IMG_BORDER; ++r)
                for (int c = SIFT_IMG_BORDER; c < W - SIFT_IMG_BORDER; ++c) {
                    float val = AT(img, r, c);
                    if (!(fabsf(val) > (float)threshold)) continue;
                    int ext = 1;
                    if (val > 0) {
                        for (int dr = -1; dr <= 1 && ext; ++dr)
                            for (int dc = -1; dc <= 1; ++dc)
                                if (!(val >= AT(img, r + dr, c + dc) && val >= AT(prev, r + dr, c + dc) && val >= AT(next, r + dr, c + dc))) {
                                    ext = 0;
                                    break;
                                }
                    } else {
                        for (int dr = -1; dr <= 1 && ext; ++dr)
                            for (int dc = -1; dc <= 1; ++dc)
                                if (!(val <= AT(img, r + dr, c + dc) && val <= AT(prev, r + dr, c + dc) && val <= AT(next, r + dr, c + dc))) {
                                    ext = 0;
                                    break;
                                }
                    }
                    if (!ext) continue;
                    sift_kp kpt;
                    int r1 = r, c1 = c, layer = i;
                    if (!sift_adjust(P, o, &layer, &r1, &c1, &kpt)) continue;
                    float scl_octv = kpt.size * 0.5f / (1 << o);
                    float omax = sift_ori_hist(P->gauss[o][layer], W, H, c1, r1, cv_round_f(SIFT_ORI_RADIUS * scl_octv),
                                               SIFT_ORI_SIG_FCTR * scl_octv, hist);
                    float mag_thr = omax * SIFT_ORI_PEAK_RATIO;
                    for (int j = 0; j < n; ++j) {
                        int l = j > 0 ? j - 1 : n - 1;
                        int r2 = j < n - 1 ? j + 1 : 0;
                        if (hist[j] > hist[l] && hist[j] > hist[r2] && hist[j] >= mag_thr) {
                            float bin = j + 0.5f * (hist[l] - hist[r2]) / (hist[l] - 2 * hist[j] + hist[r2]);
                            bin = bin < 0 ? n + bin : bin >= n ? bin - n : bin;
                            kpt.angle = fmaf(-(360.f / n), bin, 360.f);
                            if (fabsf(kpt.angle - 360.f) < 1.1920928955078125e-7f) kpt.angle = 0.f;
                            kp_push(K, &kpt);
                        }
                    }
                }
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * S.9  KeyPointsFilter::removeDuplicatedSorted (KeyPoint_LessThan total order, then unique on (pt, size, angle))
 * ------------------------------------------------------------------------------------------- */
static int kp_less(const void* a_, const void* b_) {
    const sift_kp *a = (const sift_kp*)a_, *b = (const sift_kp*)b_;
    if (a->x != b->x) return a->x < b->x ? -1 : 1;
    if (a->y != b->y) return a->y < b->y ? -1 : 1;
    if (a->size != b->size) return a->size > b->size ? -1 : 1;
    if (a->angle != b->angle) return a->angle < b->angle ? -1 : 1;
    if (a->response != b->response) return a->response > b->response ? -1 : 1;
    if (a->octave != b->octave) return a->octave > b->octave ? -1 : 1;
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * S.10 calcSIFTDescriptor: 4x4x8 histogram with trilinear interpolation, accumulated in raster order
 * ------------------------------------------------------------------------------------------- */
static void sift_descriptor(const float* img, int W, int H, float ptx, float pty, float ori, float scl, float* dst) {
    const int d = SIFT_DESCR_WIDTH, n = SIFT_DESCR_HIST_BINS;
    int px = cv_round_f(ptx), py = cv_round_f(pty);
    float cos_t = cosf(ori * (float)(3.14159265358979323846 / 180));
    float sin_t = sinf(ori * (float)(3.14159265358979323846 / 180));
    float bins_per_rad = n / 360.f;
    float exp_scale = -1.f / (d * d * 0.5f);
    float hist_width = SIFT_DESCR_SCL_FCTR * scl;
    int radius = cv_round_f(hist_width * 1.4142135623730951f * (d + 1) * 0.5f);
    int diag = (int)sqrt((double)W * W + (double)H * H);
    if (radius > diag) radius = diag;
    cos_t /= hist_width;
    sin_t /= hist_width;
    float hist[(SIFT_DESCR_WIDTH + 2) * (SIFT_DESCR_WIDTH + 2) * (SIFT_DESCR_HIST_BINS + 2)];
    float raw[SIFT_DESCR_WIDTH * SIFT_DESCR_WIDTH * SIFT_DESCR_HIST_BINS];
    memset(hist, 0, sizeof(hist));
    for (int i = -radius; i <= radius; ++i)
        for (int j = -radius; j <= radius; ++j) {
            float c_rot = j * cos_t - i * sin_t;
            float r_rot = j * sin_t + i * cos_t;
            float rbin = r_rot + d / 2 - 0.5f;
            float cbin = c_rot + d / 2 - 0.5f;
            int r = py + i, c = px + j;
            if (rbin > -1 && rbin < d && cbin > -1 && cbin < d && r > 0 && r < H - 1 && c > 0 && c < W - 1) {
                float dx = AT(img, r, c + 1) - AT(img, r, c - 1);
                float dy = AT(img, r - 1, c) - AT(img, r + 1, c);
                float wgt = sift_exp32f((c_rot * c_rot + r_rot * r_rot) * exp_scale);
                float o = sift_fast_atan2(dy, dx);
                float mag = sift_magnitude(dx, dy) * wgt;
                float obin = (o - ori) * bins_per_rad;
                int r0 = cv_floor_f(rbin), c0 = cv_floor_f(cbin), o0 = cv_floor_f(obin);
                rbin -= r0;
                cbin -= c0;
                obin -= o0;
                if (o0 < 0) o0 += n;
                if (o0 >= n) o0 -= n;
                float v_r1 = mag * rbin, v_r0 = mag - v_r1;
                float v_rc11 = v_r1 * cbin, v_rc10 = v_r1 - v_rc11;
                float v_rc01 = v_r0 * cbin, v_rc00 = v_r0 - v_rc01;
                float v_rco111 = v_rc11 * obin, v_rco110 = v_rc11 - v_rco111;
                float v_rco101 = v_rc10 * obin, v_rco100 = v_rc10 - v_rco101;
                float v_rco011 = v_rc01 * obin, v_rco010 = v_rc01 - v_rco011;
                float v_rco001 = v_rc00 * obin, v_rco000 = v_rc00 - v_rco001;
                int idx = ((r0 + 1) * (d + 2) + c0 + 1) * (n + 2) + o0;
                hist[idx] += v_rco000;
                hist[idx + 1] += v_rco001;
                hist[idx + (n + 2)] += v_rco010;
                hist[idx + (n + 3)] += v_rco011;
                hist[idx + (d + 2) * (n + 2)] += v_rco100;
                hist[idx + (d + 2) * (n + 2) + 1] += v_rco101;
                hist[idx + (d + 3) * (n + 2)] += v_rco110;
                hist[idx + (d + 3) * (n + 2) + 1] += v_rco111;
            }
        }
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
            int idx = ((i + 1) * (d + 2) + (j + 1)) * (n + 2);
            hist[idx] += hist[idx + n];
            hist[idx + 1] += hist[idx + n + 1];
            for (int k = 0; k < n; ++k) raw[(i * d + j) * n + k] = hist[idx + k];
        }
    const int len = d * d * n;
    /* nrm2: the AVX2 build sums 8 lanes with FMA, then reduces; the tolerance of the end-to-end pin covers the order */
    float nrm2 = 0;
    for (int k = 0; k < len; ++k) nrm2 = fmaf(raw[k], raw[k], nrm2);
    float thr = sqrtf(nrm2) * SIFT_DESCR_MAG_THR;
    nrm2 = 0;
    for (int k = 0; k < len; ++k) {
        float val = raw[k] < thr ? raw[k] : thr;
        raw[k] = val;
        nrm2 = fmaf(val, val, nrm2);
    }
    float s = sqrtf(nrm2);
    nrm2 = SIFT_INT_DESCR_FCTR / (s > 1.1920928955078125e-7f ? s : 1.1920928955078125e-7f);
    for (int k = 0; k < len; ++k) {
        int v = cv_round_f(raw[k] * nrm2);
        dst[k] = (float)(v < 0 ? 0 : v > 255 ? 255 : v);
    }
}

/* ---------------------------------------------------------------------------------------------
 * S.11 SIFT::detectAndCompute (gray 8-bit input).  Output order = OpenCV's (KeyPoint_LessThan after dedup).
 *      kp_f: n x 5 {pt.x, pt.y, size, angle, response}, kp_octave: n packed octaves, desc: n x 128 floats.
 *      Returns the keypoint count (may exceed cap; only cap entries are written).
 * ------------------------------------------------------------------------------------------- */
int sift_detect_and_compute(const uint8_t* gray, int w, int h, float* kp_f, int32_t* kp_octave, float* desc, int cap) {
    sift_pyr P;
    kp_vec K = {0, 0, 0};
    sift_build(gray, w, h, &P);
    sift_extrema(&P, &K);
    int n = K.n;
    if (n > 1) {
        qsort(K.v, (size_t)n, sizeof(sift_kp), kp_less);
        int i = 0;
        for (int j = 1; j < n; ++j) {
            const sift_kp *a = &K.v[i], *b = &K.v[j];
            if (a->x != b->x || a->y != b->y || a->size != b->size || a->angle != b->angle) K.v[++i] = K.v[j];
        }
        n = i + 1;
    }
    for (int i = 0; i < n; ++i) { /* firstOctave = -1: back to input-image coordinates */
        sift_kp* k = &K.v[i];
        k->octave = (k->octave & ~255) | ((k->octave - 1) & 255);
        k->x *= 0.5f;
        k->y *= 0.5f;
        k->size *= 0.5f;
    }
    for (int i = 0; i < n && i < cap; ++i) {
        const sift_kp* k = &K.v[i];
        if (kp_f) {
            kp_f[5 * i + 0] = k->x;
            kp_f[5 * i + 1] = k->y;
            kp_f[5 * i + 2] = k->size;
            kp_f[5 * i + 3] = k->angle;
            kp_f[5 * i + 4] = k->response;
        }
        if (kp_octave) kp_octave[i] = k->octave;
        if (desc) { /* calcDescriptors: unpackOctave, image of (octave - firstOctave, layer) */
            int octave = k->octave & 255, layer = (k->octave >> 8) & 255;
            octave = octave < 128 ? octave : (-128 | octave);
            float scale = octave >= 0 ? 1.f / (1 << octave) : (float)(1 << -octave);
            float size = k->size * scale;
            int o = octave + 1;
            float angle = 360.f - k->angle;
            if (fabsf(angle - 360.f) < 1.1920928955078125e-7f) angle = 0.f;
            sift_descriptor(P.gauss[o][layer], P.w[o], P.h[o], k->x * scale, k->y * scale, angle, size * 0.5f, desc + (size_t)128 * i);
        }
    }
    free(K.v);
    sift_free(&P);
    return n;
}
