// trig.cuh -- sin / cos / atan2 for the arm22 model on sm_100a.
//
// Reference call sites: diff_angles (src/utils.jl:97-99) and the joint positions of
// src/models/arm22.jl:22-27,69-70,122-131.  Julia evaluates them with its own msun-derived
// kernels (base/special/trig.jl, rem_pio2.jl; not vendored in the reference, not pinned).
// CUDA's libdevice sin/cos/atan2 differ from those by up to 1-2 ulp, which would flip
// verdicts that sit on a threshold, so the published msun algorithm (k_sin, k_cos, the medium
// branch of rem_pio2, atan, atan2) is restated here with fused multiply-adds exactly where
// Julia's evalpoly uses muladd and individually rounded operations elsewhere.  sin and cos of
// one angle share a single argument reduction (sincos), which returns the same two values as
// separate calls.
//
// Domain: |x| <= 2^20 * pi/2 (joint angles are O(10)); beyond it the device falls back to
// libdevice and bit parity with the CPU restatement (which uses libm there) is not claimed.
#pragma once
#include <string.h>

#include "geom.cuh"

namespace mrmp {

MRMP_HD unsigned hi_word(double x) {
#ifdef __CUDA_ARCH__
    return (unsigned)__double2hiint(x);
#else
    unsigned long long u;
    memcpy(&u, &x, 8);
    return (unsigned)(u >> 32);
#endif
}

// sin(y + ylo) for |y| <= pi/4: odd polynomial of degree 13 (msun k_sin.c)
MRMP_HD double sin_kernel(double y, double ylo) {
    const double y2 = dmul(y, y);
    const double y4 = dmul(y2, y2);
    const double hi = dfma(y2, dfma(y2, 2.75573137070700676789e-06, -1.98412698298579493134e-04),
                           8.33333333332248946124e-03);
    const double lo = dfma(y2, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    const double r = dadd(hi, dmul(dmul(y2, y4), lo));
    const double y3 = dmul(y2, y);
    const double inner = dsub(dmul(y2, dsub(dmul(0.5, ylo), dmul(y3, r))), ylo);
    return dsub(y, dsub(inner, dmul(y3, -1.66666666666666324348e-01)));
}

// cos(y + ylo) for |y| <= pi/4 (msun k_cos.c)
MRMP_HD double cos_kernel(double y, double ylo) {
    const double y2 = dmul(y, y);
    const double y4 = dmul(y2, y2);
    const double p1 = dfma(y2, dfma(y2, 2.48015872894767294178e-05, -1.38888888888741095749e-03),
                           4.16666666666666019037e-02);
    const double p2 = dfma(y2, dfma(y2, -1.13596475577881948265e-11, 2.08757232129817482790e-09),
                           -2.75573143513906633035e-07);
    const double r = dadd(dmul(y2, p1), dmul(dmul(y4, y4), p2));
    const double hy2 = dmul(0.5, y2);
    const double w = dsub(1.0, hy2);
    return dadd(w, dadd(dsub(dsub(1.0, w), hy2), dsub(dmul(y2, r), dmul(y, ylo))));
}

// x = n*pi/2 + (y0 + y1), |y0| <= pi/4: Cody-Waite with 33+33+53-bit pieces of pi/2
// (msun e_rem_pio2.c, medium branch)
MRMP_HD int reduce_pio2(double x, double &y0, double &y1) {
    const int ex = (int)((hi_word(x) & 0x7fffffffu) >> 20);
    const double fn = rint(dmul(x, 6.36619772367581382433e-01));
    double r = dsub(x, dmul(fn, 1.57079632673412561417e+00));
    double w = dmul(fn, 6.07710050650619224932e-11);
    double y = dsub(r, w);
    if (ex - (int)((hi_word(y) >> 20) & 0x7ff) > 16) {  // cancellation: second piece
        double t = r;
        w = dmul(fn, 6.07710050630396597660e-11);
        r = dsub(t, w);
        w = dsub(dmul(fn, 2.02226624879595063154e-21), dsub(dsub(t, r), w));
        y = dsub(r, w);
        if (ex - (int)((hi_word(y) >> 20) & 0x7ff) > 49) {  // third piece
            t = r;
            w = dmul(fn, 2.02226624871116645580e-21);
            r = dsub(t, w);
            w = dsub(dmul(fn, 8.47842766036889956997e-32), dsub(dsub(t, r), w));
            y = dsub(r, w);
        }
    }
    y0 = y;
    y1 = dsub(dsub(r, y), w);
    return (int)fn;
}

constexpr double kTrigMediumMax = 1647099.3291652855;  // 2^20 * pi/2

// s = sin(x), c = cos(x)
MRMP_HD void sincos_rn(double x, double &s, double &c) {
    const unsigned ix = hi_word(x) & 0x7fffffffu;
    if (ix <= 0x3fe921fbu) {  // |x| <~ pi/4: no reduction
        s = ix < 0x3e500000u ? x : sin_kernel(x, 0.0);    // |x| < 2^-26
        c = ix < 0x3e46a09eu ? 1.0 : cos_kernel(x, 0.0);  // |x| < 2^-27 sqrt(2)
        return;
    }
    if (!(fabs(x) <= kTrigMediumMax)) {  // out of the documented domain (also NaN / inf)
        s = sin(x);
        c = cos(x);
        return;
    }
    double y0, y1;
    const int n = reduce_pio2(x, y0, y1);
    const double ks = sin_kernel(y0, y1), kc = cos_kernel(y0, y1);
    switch (n & 3) {
        case 0: s = ks; c = kc; break;
        case 1: s = kc; c = -ks; break;
        case 2: s = -ks; c = -kc; break;
        default: s = -kc; c = ks; break;
    }
}

// odd/even split of the degree-11 atan polynomial in x^2 (msun s_atan.c)
MRMP_HD double atan_poly(double x) {
    const double x2 = dmul(x, x);
    const double x4 = dmul(x2, x2);
    double p = dfma(x4, 1.62858201153657823623e-02, 4.97687799461593236017e-02);
    p = dfma(x4, p, 6.66107313738753120669e-02);
    p = dfma(x4, p, 9.09088713343650656196e-02);
    p = dfma(x4, p, 1.42857142725034663711e-01);
    p = dfma(x4, p, 3.33333333333329318027e-01);
    double q = dfma(x4, -3.65315727442169155270e-02, -5.83357013379057348645e-02);
    q = dfma(x4, q, -7.69187620504482999495e-02);
    q = dfma(x4, q, -1.11111104054623557880e-01);
    q = dfma(x4, q, -1.99999999998764832476e-01);
    return dadd(dmul(x2, p), dmul(x4, q));
}

MRMP_HD double atan_rn(double x) {
    const double ax = fabs(x);
    if (isnan(x)) return x;
    if (ax >= 73786976294838206464.0) return copysign(1.5707963267948966, x);  // 2^66
    if (ax < 0.4375) {
        if (ax < 7.450580596923828125e-09) return x;  // 2^-27
        return dsub(x, dmul(x, atan_poly(x)));
    }
    double hi, lo, t;
    if (ax < 0.6875) {  // around atan(1/2)
        hi = 4.63647609000806093515e-01;
        lo = 2.26987774529616870924e-17;
        t = ddiv(dsub(dmul(2.0, ax), 1.0), dadd(2.0, ax));
    } else if (ax < 1.1875) {  // around atan(1)
        hi = 7.85398163397448278999e-01;
        lo = 3.06161699786838301793e-17;
        t = ddiv(dsub(ax, 1.0), dadd(ax, 1.0));
    } else if (ax < 2.4375) {  // around atan(3/2)
        hi = 9.82793723247329054082e-01;
        lo = 1.39033110312309984516e-17;
        t = ddiv(dsub(ax, 1.5), dadd(1.0, dmul(1.5, ax)));
    } else {  // around atan(inf)
        hi = 1.57079632679489655800e+00;
        lo = 6.12323399573676603587e-17;
        t = ddiv(-1.0, ax);
    }
    const double z = dsub(hi, dsub(dsub(dmul(t, atan_poly(t)), lo), t));
    return copysign(z, x);
}

// atan(y, x) (msun e_atan2.c)
MRMP_HD double atan2_rn(double y, double x) {
    const double kPi = 3.1415926535897931160e+00, kPiLo = 1.2246467991473531772e-16;
    if (isnan(x) || isnan(y)) return dnan();
    if (x == 1.0) return atan_rn(y);
    const bool xneg = signbit(x), yneg = signbit(y);
    if (y == 0.0) return xneg ? (yneg ? -kPi : kPi) : y;
    if (x == 0.0) return copysign(dmul(kPi, 0.5), y);
    if (isinf(x)) {
        if (isinf(y)) {
            const double q = dmul(kPi, 0.25);
            const double a = xneg ? dmul(3.0, q) : q;
            return yneg ? -a : a;
        }
        const double a = xneg ? kPi : 0.0;
        return yneg ? -a : a;
    }
    if (isinf(y)) return copysign(dmul(kPi, 0.5), y);
    const int k = ((int)(hi_word(y) & 0x7fffffffu) - (int)(hi_word(x) & 0x7fffffffu)) >> 20;
    double z;
    bool fold = xneg;
    if (k > 60) {  // |y/x| > 2^60
        z = dadd(dmul(kPi, 0.5), dmul(0.5, kPiLo));
        fold = false;
    } else if (xneg && k < -60) {  // 0 > |y|/x > -2^-60
        z = 0.0;
    } else {
        z = atan_rn(fabs(ddiv(y, x)));
    }
    if (!fold) return yneg ? -z : z;
    return yneg ? dsub(dsub(z, kPiLo), kPi) : dsub(kPi, dsub(z, kPiLo));
}

// src/utils.jl:97-99: atan(sin(t1 - t2), cos(t1 - t2))
MRMP_HD double diff_angles(double t1, double t2) {
    double s, c;
    sincos_rn(dsub(t1, t2), s, c);
    return atan2_rn(s, c);
}

}  // namespace mrmp
