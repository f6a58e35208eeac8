/*
 * orc_trig.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * sin / cos / atan2 for the arm22 model (reference call sites:
 * src/utils.jl:97-99 diff_angles, src/models/arm22.jl:22-27,69-70,122-131).
 *
 * The reference evaluates these with Julia Base's pure-Julia ports of FreeBSD
 * msun (base/special/trig.jl, base/special/rem_pio2.jl; Julia is NOT vendored in
 * /root/reference and is not pinned: Project.toml only says julia >= 1.6).  This
 * file restates the published msun algorithm (k_sin.c, k_cos.c, e_rem_pio2.c
 * medium path, s_atan.c, e_atan2.c) with explicit fma() exactly where Julia's
 * @horner/evalpoly uses muladd, and plain rounded ops elsewhere.  PARITY
 * UNPINNED vs Julia (<= 1 ulp expected); it is validated against the host libm
 * to <= 1 ulp (tests/test_oracle_trig.py) and is bit-identical to the separately
 * written device version (sssp_b200/csrc/trig.cuh) on the tested domain.
 *
 * Domain: |x| <= 2^20*pi/2 uses the Cody-Waite reduction below; larger |x|
 * (unreachable for joint angles) falls back to libm.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "mrmp_oracle.h"

static inline uint32_t hi_word(double x) {
    uint64_t u;
    memcpy(&u, &x, 8);
    return (uint32_t)(u >> 32);
}

/* k_sin.c / k_cos.c coefficients */
static const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                    S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                    S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
static const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                    C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                    C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;

/* sin on [-pi/4,pi/4] of the double-double (y, ylo) */
static double k_sin(double y, double ylo) {
    double y2 = y * y;
    double y4 = y2 * y2;
    double r = fma(y2, fma(y2, S4, S3), S2) + y2 * y4 * fma(y2, S6, S5);
    double y3 = y2 * y;
    return y - ((y2 * (0.5 * ylo - y3 * r) - ylo) - y3 * S1);
}

static double k_cos(double y, double ylo) {
    double y2 = y * y;
    double y4 = y2 * y2;
    double r = y2 * fma(y2, fma(y2, C3, C2), C1) + y4 * y4 * fma(y2, fma(y2, C6, C5), C4);
    double hy2 = 0.5 * y2;
    double w = 1.0 - hy2;
    return w + (((1.0 - w) - hy2) + (y2 * r - y * ylo));
}

static const double INVPIO2 = 6.36619772367581382433e-01, PIO2_1 = 1.57079632673412561417e+00,
                    PIO2_1T = 6.07710050650619224932e-11, PIO2_2 = 6.07710050630396597660e-11,
                    PIO2_2T = 2.02226624879595063154e-21, PIO2_3 = 2.02226624871116645580e-21,
                    PIO2_3T = 8.47842766036889956997e-32;

/* e_rem_pio2.c medium path; returns n mod 4 in the low bits, y = x - n*pi/2 */
static int rem_pio2_medium(double x, double *y0, double *y1) {
    uint32_t ix = hi_word(x) & 0x7fffffffu;
    double fn = rint(x * INVPIO2);
    int n = (int)fn;
    double r = x - fn * PIO2_1;
    double w = fn * PIO2_1T;
    int j = (int)(ix >> 20);
    double y = r - w;
    int i = j - (int)((hi_word(y) >> 20) & 0x7ff);
    if (i > 16) {
        double t = r;
        w = fn * PIO2_2;
        r = t - w;
        w = fn * PIO2_2T - ((t - r) - w);
        y = r - w;
        i = j - (int)((hi_word(y) >> 20) & 0x7ff);
        if (i > 49) {
            t = r;
            w = fn * PIO2_3;
            r = t - w;
            w = fn * PIO2_3T - ((t - r) - w);
            y = r - w;
        }
    }
    *y0 = y;
    *y1 = (r - y) - w;
    return n;
}

#define TRIG_MEDIUM_MAX 1647099.3291652855 /* 2^20 * pi/2 */

double orc_sin(double x) {
    uint32_t ix = hi_word(x) & 0x7fffffffu;
    if (ix <= 0x3fe921fbu) {          /* |x| ~<= pi/4 */
        if (ix < 0x3e500000u) return x; /* |x| < 2^-26 */
        return k_sin(x, 0.0);
    }
    if (!(fabs(x) <= TRIG_MEDIUM_MAX)) return sin(x);
    double y0, y1;
    int n = rem_pio2_medium(x, &y0, &y1);
    switch (n & 3) {
    case 0: return k_sin(y0, y1);
    case 1: return k_cos(y0, y1);
    case 2: return -k_sin(y0, y1);
    default: return -k_cos(y0, y1);
    }
}

double orc_cos(double x) {
    uint32_t ix = hi_word(x) & 0x7fffffffu;
    if (ix <= 0x3fe921fbu) {
        if (ix < 0x3e46a09eu) return 1.0; /* |x| < 2^-27 * sqrt(2) */
        return k_cos(x, 0.0);
    }
    if (!(fabs(x) <= TRIG_MEDIUM_MAX)) return cos(x);
    double y0, y1;
    int n = rem_pio2_medium(x, &y0, &y1);
    switch (n & 3) {
    case 0: return k_cos(y0, y1);
    case 1: return -k_sin(y0, y1);
    case 2: return -k_cos(y0, y1);
    default: return k_sin(y0, y1);
    }
}

/* s_atan.c */
static const double ATANHI[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01,
                                 9.82793723247329054082e-01, 1.57079632679489655800e+00};
static const double ATANLO[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17,
                                 1.39033110312309984516e-17, 6.12323399573676603587e-17};
static const double AT[11] = {
    3.33333333333329318027e-01,  -1.99999999998764832476e-01, 1.42857142725034663711e-01,
    -1.11111104054623557880e-01, 9.09088713343650656196e-02,  -7.69187620504482999495e-02,
    6.66107313738753120669e-02,  -5.83357013379057348645e-02, 4.97687799461593236017e-02,
    -3.65315727442169155270e-02, 1.62858201153657823623e-02};

static inline double atan_pq(double x) {
    double x2 = x * x;
    double x4 = x2 * x2;
    double p = x2 * fma(x4, fma(x4, fma(x4, fma(x4, fma(x4, AT[10], AT[8]), AT[6]), AT[4]), AT[2]),
                        AT[0]);
    double q = x4 * fma(x4, fma(x4, fma(x4, fma(x4, AT[9], AT[7]), AT[5]), AT[3]), AT[1]);
    return p + q;
}

static double orc_atan(double x) {
    double ax = fabs(x);
    if (isnan(x)) return x;
    if (ax >= 0x1p66) return copysign(1.5707963267948966, x);
    if (ax < 0.4375) {
        if (ax < 0x1p-27) return x;
        return x - x * atan_pq(x);
    }
    int id;
    double t;
    if (ax < 1.1875) {
        if (ax < 0.6875) {
            id = 0;
            t = (2.0 * ax - 1.0) / (2.0 + ax);
        } else {
            id = 1;
            t = (ax - 1.0) / (ax + 1.0);
        }
    } else {
        if (ax < 2.4375) {
            id = 2;
            t = (ax - 1.5) / (1.0 + 1.5 * ax);
        } else {
            id = 3;
            t = -1.0 / ax;
        }
    }
    double z = ATANHI[id] - ((t * atan_pq(t) - ATANLO[id]) - t);
    return copysign(z, x);
}

/* e_atan2.c */
double orc_atan2(double y, double x) {
    static const double PI = 3.1415926535897931160e+00, PI_LO = 1.2246467991473531772e-16;
    if (isnan(x) || isnan(y)) return NAN;
    if (x == 1.0) return orc_atan(y);
    int m = 2 * (signbit(x) ? 1 : 0) + (signbit(y) ? 1 : 0);
    if (y == 0.0) {
        if (m < 2) return y;
        return m == 2 ? PI : -PI;
    }
    if (x == 0.0) return copysign(PI / 2, y);
    if (isinf(x)) {
        if (isinf(y)) {
            switch (m) {
            case 0: return PI / 4;
            case 1: return -PI / 4;
            case 2: return 3 * (PI / 4);
            default: return -3 * (PI / 4);
            }
        }
        switch (m) {
        case 0: return 0.0;
        case 1: return -0.0;
        case 2: return PI;
        default: return -PI;
        }
    }
    if (isinf(y)) return copysign(PI / 2, y);
    int32_t k = ((int32_t)(hi_word(y) & 0x7fffffffu) - (int32_t)(hi_word(x) & 0x7fffffffu)) >> 20;
    double z;
    if (k > 60) {
        z = PI / 2 + 0.5 * PI_LO;
        m &= 1;
    } else if (x < 0 && k < -60) {
        z = 0.0;
    } else {
        z = orc_atan(fabs(y / x));
    }
    switch (m) {
    case 0: return z;
    case 1: return -z;
    case 2: return PI - (z - PI_LO);
    default: return (z - PI_LO) - PI;
    }
}
