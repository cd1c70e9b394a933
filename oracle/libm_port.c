/* libm_port.c -- TEST INFRASTRUCTURE.  glibc's float exp2 / sin / cos restated in plain C: the algorithms slideo_b200/csrc/sift.cu
 * evaluates on the device (glibc_exp2f / glibc_sincosf) in place of the libm calls of OpenCV's SIFT (the keypoint size uses
 * powf(2, x) == exp2f, calcSIFTDescriptor uses cosf / sinf of the orientation).  Source of the algorithms: glibc
 * sysdeps/ieee754/flt-32/e_exp2f.c + exp2f_data, s_sincosf.h + s_sincosf_data.c (the ARM optimized-routines code, glibc >= 2.28);
 * third party, not under /root/reference.  libm_port_mismatches() pins the restatement against the libm of the machine it runs
 * on (tests/test_oracle_sift.py::test_libm_port_equals_libm): 0 mismatches on every 7th float of [0, 1.6] / every 5th of [0, 6.4]
 * with glibc 2.39, with and without FMA contraction. */
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline uint32_t asuint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline uint64_t asuint64(double f) { uint64_t u; memcpy(&u, &f, 8); return u; }
static inline double asdouble(uint64_t u) { double f; memcpy(&f, &u, 8); return f; }

static const uint64_t EXP2F_TAB[32] = {
    0x3ff0000000000000, 0x3fefd9b0d3158574, 0x3fefb5586cf9890f, 0x3fef9301d0125b51, 0x3fef72b83c7d517b, 0x3fef54873168b9aa,
    0x3fef387a6e756238, 0x3fef1e9df51fdee1, 0x3fef06fe0a31b715, 0x3feef1a7373aa9cb, 0x3feedea64c123422, 0x3feece086061892d,
    0x3feebfdad5362a27, 0x3feeb42b569d4f82, 0x3feeab07dd485429, 0x3feea47eb03a5585, 0x3feea09e667f3bcd, 0x3fee9f75e8ec5f74,
    0x3feea11473eb0187, 0x3feea589994cce13, 0x3feeace5422aa0db, 0x3feeb737b0cdc5e5, 0x3feec49182a3f090, 0x3feed503b23e255d,
    0x3feee89f995ad3ad, 0x3feeff76f2fb5e47, 0x3fef199bdd85529c, 0x3fef3720dcef9069, 0x3fef5818dcfba487, 0x3fef7c97337b9b5f,
    0x3fefa4afa2a490da, 0x3fefd0765b6e4540};

/* e_exp2f.c, |x| < 128 */
float libm_port_exp2f(float x) {
    const double SHIFT = 0x1.8p+52 / 32;
    const double C0 = 0x1.c6af84b912394p-5, C1 = 0x1.ebfce50fac4f3p-3, C2 = 0x1.62e42ff0c52d6p-1;
    double xd = (double)x, kd = xd + SHIFT;
    uint64_t ki = asuint64(kd);
    kd -= SHIFT;
    double r = xd - kd;
    uint64_t t = EXP2F_TAB[ki % 32] + (ki << 47);
    double s = asdouble(t), z = C0 * r + C1, r2 = r * r, y = C2 * r + 1;
    y = z * r2 + y;
    return (float)(y * s);
}

typedef struct { double sign[4], hpi_inv, hpi, c0, c1, c2, c3, c4, s1, s2, s3; } sincos_t;
static const sincos_t SC[2] = {
    {{1.0, -1.0, -1.0, 1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, 0x1p0, -0x1.ffffffd0c621cp-2, 0x1.55553e1068f19p-5,
     -0x1.6c087e89a359dp-10, 0x1.99343027bf8c3p-16, -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13},
    {{1.0, -1.0, -1.0, 1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, -0x1p0, 0x1.ffffffd0c621cp-2, -0x1.55553e1068f19p-5,
     0x1.6c087e89a359dp-10, -0x1.99343027bf8c3p-16, -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13}};
static inline uint32_t abstop12(float x) { return (asuint(x) >> 20) & 0x7ff; }
static inline float sinf_poly(double x, double x2, const sincos_t* p, int n) {
    if ((n & 1) == 0) {
        double x3 = x * x2, s1 = p->s2 + x2 * p->s3, x7 = x3 * x2, s = x + x3 * p->s1;
        return (float)(s + x7 * s1);
    }
    double x4 = x2 * x2, c2 = p->c3 + x2 * p->c4, c1 = p->c1 + x2 * p->c2, x6 = x4 * x2, c = p->c0 + x2 * c1;
    return (float)(c + x6 * c2);
}
static inline double reduce_fast(double x, const sincos_t* p, int* np) {
    double r = x * p->hpi_inv;
    int n = ((int32_t)r + 0x800000) >> 24;
    *np = n;
    return x - n * p->hpi;
}
/* s_sinf.c / s_cosf.c, |y| < 120 */
float libm_port_sinf(float y) {
    double x = y, s;
    int n;
    const sincos_t* p = &SC[0];
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
        s = x * x;
        if (abstop12(y) < abstop12(0x1p-12f)) return y;
        return sinf_poly(x, s, p, 0);
    }
    x = reduce_fast(x, p, &n);
    s = p->sign[n & 3];
    if (n & 2) p = &SC[1];
    return sinf_poly(x * s, x * x, p, n);
}
float libm_port_cosf(float y) {
    double x = y, s;
    int n;
    const sincos_t* p = &SC[0];
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
        double x2 = x * x;
        if (abstop12(y) < abstop12(0x1p-12f)) return 1.0f;
        return sinf_poly(x, x2, p, 1);
    }
    x = reduce_fast(x, p, &n);
    s = p->sign[n & 3];
    if (n & 2) p = &SC[1];
    return sinf_poly(x * s, x * x, p, n ^ 1);
}

/* which: 0 exp2f, 1 sinf, 2 cosf.  Every `step`-th float of [lo, hi] (lo, hi >= 0): number of inputs on which the restatement and
 * this machine's libm disagree in any bit. */
long libm_port_mismatches(int which, float lo, float hi, uint32_t step) {
    long bad = 0;
    for (uint32_t u = asuint(lo); u <= asuint(hi); u += step) {
        float x;
        memcpy(&x, &u, 4);
        float a = which == 0 ? libm_port_exp2f(x) : which == 1 ? libm_port_sinf(x) : libm_port_cosf(x);
        float b = which == 0 ? exp2f(x) : which == 1 ? sinf(x) : cosf(x);
        bad += asuint(a) != asuint(b);
    }
    return bad;
}
