// geom.cuh -- FP64 geometry primitives of the connect/collide hot path (device side).
//
// Restates src/utils.jl:17-94 of the reference for sm_100a.  Every arithmetic step uses
// the round-to-nearest intrinsics (__dadd_rn, __dmul_rn, ...), which nvcc never contracts
// into FMA, so results are bit-identical to the reference's un-fused Julia arithmetic
// independent of -fmad.
//
// Two exact (verdict-preserving) transformations are used throughout:
//  (1) sqrt elimination.  The reference compares  norm(p) = sqrt(sum p_k^2)  with a
//      threshold.  RN(sqrt(.)) is monotone, so  sqrt(s) < thr  <=>  s < T2  where
//      T2 = min{ s : RN(sqrt(s)) >= thr } is precomputed once per threshold on the host
//      (abi.cu: sq_lt_threshold).  Julia's generic_norm2 only leaves the plain
//      sum-of-squares path when the sum under/overflows; s == 0 catches the underflow
//      case and takes norm_slow().
//  (2) endpoint short-cut of the segment-point distance: for e == 1 / e == 0 the
//      reference evaluates (e*a + (1-e)*b) - c, which is exactly a-c / b-c.
//
// The primitives are __host__ __device__: the device side is the product; the host side
// exists only so that tests/hostprobe can run THIS source on the CPU (no GPU in the build
// container) and compare it bit for bit with the CPU checker.  No library entry point calls them
// on the host.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define MRMP_HD __host__ __device__ __forceinline__

namespace mrmp {

#ifdef __CUDA_ARCH__
MRMP_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
MRMP_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
MRMP_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
MRMP_HD double ddiv(double a, double b) { return __ddiv_rn(a, b); }
MRMP_HD double dsqrt(double a) { return __dsqrt_rn(a); }
MRMP_HD double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }
MRMP_HD double dnan() { return __longlong_as_double(0x7ff8000000000000LL); }
#else
// host build (tests/hostprobe only): volatile blocks contraction / reassociation
MRMP_HD double dadd(double a, double b) { volatile double r = a + b; return r; }
MRMP_HD double dsub(double a, double b) { volatile double r = a - b; return r; }
MRMP_HD double dmul(double a, double b) { volatile double r = a * b; return r; }
MRMP_HD double ddiv(double a, double b) { volatile double r = a / b; return r; }
MRMP_HD double dsqrt(double a) { return sqrt(a); }
MRMP_HD double dfma(double a, double b, double c) { return fma(a, b, c); }
MRMP_HD double dnan() { return NAN; }
#endif

// LinearAlgebra.dot on 2-3 element vectors: x1*y1 + x2*y2 (+ x3*y3), left to right
template <int D>
MRMP_HD double dotd(const double *a, const double *b) {
    double s = dmul(a[0], b[0]);
#pragma unroll
    for (int k = 1; k < D; ++k) s = dadd(s, dmul(a[k], b[k]));
    return s;
}

// sum of squares, left to right (fast path of LinearAlgebra.generic_norm2)
template <int D>
MRMP_HD double sumsq(const double *v) {
    double s = dmul(v[0], v[0]);
#pragma unroll
    for (int k = 1; k < D; ++k) s = dadd(s, dmul(v[k], v[k]));
    return s;
}

// generic_norm2 when the plain sum of squares is 0: either all-zero or underflow
template <int D>
__host__ __device__ __noinline__ double norm_slow(const double *v) {
    double maxabs = fabs(v[0]);
#pragma unroll
    for (int k = 1; k < D; ++k) {
        double a = fabs(v[k]);
        maxabs = (isnan(maxabs) || isnan(a)) ? dnan() : (a > maxabs ? a : maxabs);
    }
    if (maxabs == 0.0 || isinf(maxabs)) return maxabs;
    double t = ddiv(fabs(v[0]), maxabs);
    double s = dmul(t, t);
#pragma unroll
    for (int k = 1; k < D; ++k) {
        t = ddiv(fabs(v[k]), maxabs);
        s = dadd(s, dmul(t, t));
    }
    return dmul(maxabs, dsqrt(s));
}

// full norm value (used where the distance itself is returned to the host)
template <int D>
MRMP_HD double norm_val(const double *v) {
    double s = sumsq<D>(v);
    if (s == 0.0) return norm_slow<D>(v);
    return dsqrt(s);
}

// norm(v) < thr.  TAB: T2 = sq_lt_threshold(thr) is available (sqrt-free exact compare);
// otherwise the square root is taken like the reference does.
template <int D, bool TAB = true>
MRMP_HD bool norm_lt(const double *v, double s, double thr, double T2) {
    if (s == 0.0) return norm_slow<D>(v) < thr;
    if constexpr (TAB) return s < T2;
    return dsqrt(s) < thr;
}

// src/utils.jl:44-57: segment [a,b] to point c.  u = a-b and uu = dot(u,u) are passed in
// (shared between several c).
// verdict dist(a,b,c) < thr; s_probe receives the squared distance (NaN propagation of
// segseg_lt's four-way minimum)
template <int D, bool TAB = true>
MRMP_HD bool segpt_verdict(const double *a, const double *b, const double *u, double uu,
                           const double *c, double thr, double T2, double &s_probe) {
    double bc[D], ac[D], p[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        bc[k] = dsub(b[k], c[k]);
        ac[k] = dsub(a[k], c[k]);
    }
    const double df0 = dotd<D>(u, bc);  // :46
    const double df1 = dotd<D>(u, ac);  // :47
    if (dmul(df0, df1) < 0.0) {         // :50
        const double e = ddiv(-df0, uu);  // :51
        const double f = dsub(1.0, e);
#pragma unroll
        for (int k = 0; k < D; ++k)
            p[k] = dsub(dadd(dmul(e, a[k]), dmul(f, b[k])), c[k]);  // :56
    } else if (df0 >= 0.0) {  // :52  e = 0  ->  p = b - c
#pragma unroll
        for (int k = 0; k < D; ++k) p[k] = bc[k];
    } else {  // e = 1  ->  p = a - c
#pragma unroll
        for (int k = 0; k < D; ++k) p[k] = ac[k];
    }
    const double s = sumsq<D>(p);
    s_probe = s;
    return norm_lt<D, TAB>(p, s, thr, T2);
}

// verdict  dist(a,b,c) < thr
template <int D, bool TAB = true>
MRMP_HD bool segpt_lt(const double *a, const double *b, const double *u,
                                         double uu, const double *c, double thr, double T2) {
    double probe;
    return segpt_verdict<D, TAB>(a, b, u, uu, c, thr, T2, probe);
}

// Branch-free statement of :44-57 for SIMT callers: in a warp nearly every case (interior foot
// point / either end) is taken by some lane, so one predicated select of e costs no more than
// divergent branches and leaves one straight dependency chain per point that the caller can
// interleave with its neighbours.  p = e*a + (1-e)*b - c is evaluated as :56 writes it for all
// three values of e, which also reproduces the reference's NaN / Inf propagation (0 * b).
// Returns sumsq(p).
template <int D>
MRMP_HD double segpt_foot_core(const double *a, const double *b, const double *c, double uu,
                               double df0, double df1, double *p) {
    const bool inside = dmul(df0, df1) < 0.0;  // :50
    const double e_in = ddiv(-df0, uu);        // :51 (only used when inside)
    const double e = inside ? e_in : (df0 >= 0.0 ? 0.0 : 1.0);  // :49-54
    const double f = dsub(1.0, e);
#pragma unroll
    for (int x = 0; x < D; ++x) p[x] = dsub(dadd(dmul(e, a[x]), dmul(f, b[x])), c[x]);  // :56
    return sumsq<D>(p);
}

template <int D>
MRMP_HD double segpt_foot_sq(const double *a, const double *b, const double *u, double uu,
                             const double *c, double *p) {
    double bc[D], ac[D];
#pragma unroll
    for (int x = 0; x < D; ++x) {
        bc[x] = dsub(b[x], c[x]);
        ac[x] = dsub(a[x], c[x]);
    }
    const double df0 = dotd<D>(u, bc);  // :46
    const double df1 = dotd<D>(u, ac);  // :47
    return segpt_foot_core<D>(a, b, c, uu, df0, df1, p);
}

template <int D, bool TAB = true>
MRMP_HD bool segpt_lt_simt(const double *a, const double *b, const double *u, double uu,
                           const double *c, double thr, double T2) {
    double p[D];
    const double s = segpt_foot_sq<D>(a, b, u, uu, c, p);
    return norm_lt<D, TAB>(p, s, thr, T2);
}

// src/utils.jl:59-94: verdict  dist(p11,p12,p21,p22) < thr
template <int D, bool TAB = true>
MRMP_HD bool segseg_lt(const double *p11, const double *p12, const double *p21,
                                          const double *p22, double thr, double T2) {
    double v1[D], v2[D], w[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        v1[k] = dsub(p12[k], p11[k]);  // :68
        v2[k] = dsub(p22[k], p21[k]);  // :69
        w[k] = dsub(p21[k], p11[k]);
    }
    const double D1 = dotd<D>(w, v1);   // :70
    const double D2 = dotd<D>(w, v2);   // :71
    const double Dv = dotd<D>(v1, v2);  // :72
    const double V1 = dotd<D>(v1, v1);  // :73
    const double V2 = dotd<D>(v2, v2);  // :74
    const double Dd = dsub(dmul(V1, V2), dmul(Dv, Dv));  // :75
    if (Dd > 0.0) {
        // t = RN(num / Dd) lies in [0, 1]  <=>  0 <= num <= Dd  (Dd > 0; a quotient above 1
        // exceeds 1 by at least ulp(Dd)/Dd > 2^-53 and cannot round down to 1), so the two
        // divisions of :77-78 are only carried out when the interior solution is taken
        const double n1 = dsub(dmul(D1, V2), dmul(D2, Dv));
        const double n2 = dsub(dmul(D1, Dv), dmul(D2, V1));
        if (0.0 <= n1 && n1 <= Dd && 0.0 <= n2 && n2 <= Dd) {
            const double t1 = ddiv(n1, Dd);  // :77
            const double t2 = ddiv(n2, Dd);  // :78
            double q[D];
#pragma unroll
            for (int k = 0; k < D; ++k)
                q[k] = dsub(dadd(p11[k], dmul(t1, v1[k])), dadd(p21[k], dmul(t2, v2[k])));  // :80-82
            return norm_lt<D, TAB>(q, sumsq<D>(q), thr, T2);
        }
    }
    // :86-93  minimum of the four endpoint-to-segment distances (NaN-propagating).
    // The four evaluations of dist(a,b,c) (:44-57) run branch-free and side by side
    // (segpt_foot_core): four independent dependency chains for the FP64 pipe.  Their
    // difference vectors are shared: RN(x - y) = -RN(y - x) exactly, a negated operand negates
    // every product and sum of an un-fused dot exactly, and signs of zero cannot reach a verdict
    // (they only meet comparisons against 0, squares, and products with finite values), so
    //   a - b  is  -v1 / -v2,   p11 - p21 = -w,   p21 - p11 = w,
    //   p22 - p11 = -(p11 - p22),  p22 - p12 = -(p12 - p22),  p21 - p12 = -(p12 - p21),
    // and dot(a - b, a - c) of the first / third call is D1 / -D2.
    double g0[D], g1[D], g2[D], nv1[D], nv2[D], ng0[D], ng1[D], ng2[D];
#pragma unroll
    for (int x = 0; x < D; ++x) {
        g0[x] = dsub(p12[x], p21[x]);
        g1[x] = dsub(p12[x], p22[x]);
        g2[x] = dsub(p11[x], p22[x]);
        nv1[x] = -v1[x];
        nv2[x] = -v2[x];
        ng0[x] = -g0[x];
        ng1[x] = -g1[x];
        ng2[x] = -g2[x];
    }
    double s[4], p[D];
    bool hit = false;
    // dist(p11, p12, p21): u = -v1, b - c = g0, a - c = -w
    s[0] = segpt_foot_core<D>(p11, p12, p21, V1, dotd<D>(nv1, g0), D1, p);
    hit |= norm_lt<D, TAB>(p, s[0], thr, T2);
    // dist(p11, p12, p22): u = -v1, b - c = g1, a - c = g2
    s[1] = segpt_foot_core<D>(p11, p12, p22, V1, dotd<D>(nv1, g1), dotd<D>(nv1, g2), p);
    hit |= norm_lt<D, TAB>(p, s[1], thr, T2);
    // dist(p21, p22, p11): u = -v2, b - c = -g2, a - c = w
    s[2] = segpt_foot_core<D>(p21, p22, p11, V2, dotd<D>(nv2, ng2), -D2, p);
    hit |= norm_lt<D, TAB>(p, s[2], thr, T2);
    // dist(p21, p22, p12): u = -v2, b - c = -g1, a - c = -g0
    s[3] = segpt_foot_core<D>(p21, p22, p12, V2, dotd<D>(nv2, ng1), dotd<D>(nv2, ng0), p);
    hit |= norm_lt<D, TAB>(p, s[3], thr, T2);
    const double nan_probe = dadd(dadd(s[0], s[1]), dadd(s[2], s[3]));
    return hit && !isnan(nan_probe);
}


// ================================================================================================
// Verdict filter for segseg_lt (conflict-table kernel)
// ================================================================================================
// segseg_filter() evaluates the distance of two segments with a cheap, numerically stable
// formulation (FMA, per-segment reciprocals, clamped foot points; no division) and returns
//    1  the reference arithmetic is certain to say  dist < thr,
//    0  it is certain to say  !(dist < thr),
//   -1  undecided: the caller runs segseg_lt (the reference's own sequence of roundings).
// It is the classic floating-point filter of computational geometry: it never replaces the
// reference arithmetic where the two could disagree.
//
// Why the answers 0 / 1 are certain.  Domain: both segments finite with |coordinate| <= 2 (the
// reference's field is the unit square / cube), thr >= 2^-18.  Let d* be the distance of the real
// segments, d_ref what utils.jl:59-94 computes, d_f what the filter computes.
//  (a) Fallback branch of the reference (:86-93, minimum of four segment-point distances :44-57):
//      every step is a sum / product / quotient of quantities bounded by the coordinate scale with
//      no cancellation that is later amplified (the foot-point parameter e = -df0/uu enters the
//      point as e*u, so its error is scaled back by |u|); |d_ref - d*| and |d_f - d*| are both
//      below 1e-13 in the domain.
//  (b) Interior branch (:76-84): t1, t2 come from D = V1*V2 - Dv^2, which cancels for nearly
//      parallel segments.  With sin^2(angle) >= 2^-20 the error of the two points is below
//      12*eps*scale/sin^2 < 1e-7, and whichever branch the reference takes (its own t in or just
//      outside [0,1]) its value is within 1e-7 of d*: the two branches are the same continuous
//      function.  Below that conditioning the filter answers -1, except when one segment is a
//      point in the reference's own arithmetic (V == 0 exactly => D = 0 - Dv^2 <= 0 => fallback).
//  The filter answers 1 only for d_f < thr - 2^-20 and 0 only for d_f > thr + 2^-20 (or, in 2-D, 1
//  when the segments properly cross and are well conditioned: d* = 0).  2^-20 ~ 9.5e-7 exceeds
//  the combined error bound by an order of magnitude.  tests/test_hostprobe.py runs THIS source on
//  the CPU against segseg_lt on random, adversarial near-threshold, parallel, degenerate and
//  crossing pairs: the filter must never contradict it.
constexpr double kFilterBand = 9.5367431640625e-07;    // 2^-20
constexpr double kFilterSin2Min = 9.5367431640625e-07; // 2^-20
constexpr double kFilterThrMin = 3.814697265625e-06;   // 2^-18
constexpr double kFilterCoordMax = 2.0;

// Per-segment constant of the filter: 1/V with V = dot(b-a, b-a) exactly as the reference
// computes it (:73-74); 0 when V == 0 because a == b; -1 (filter disabled for this segment) when
// a coordinate is non-finite or beyond the domain, or when V is zero / subnormal-small although
// a != b (the reference's own arithmetic degenerates there: e = -df0/0).
template <int D>
MRMP_HD double seg_filter_inv(const double *a, const double *b) {
    double v[D];
    bool in = true, same = true;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        v[k] = dsub(b[k], a[k]);
        in &= fabs(a[k]) <= kFilterCoordMax && fabs(b[k]) <= kFilterCoordMax;  // NaN fails
        same &= v[k] == 0.0;
    }
    if (!in) return -1.0;
    if (same) return 0.0;
    const double V = dotd<D>(v, v);
    if (!(V >= 1e-200)) return -1.0;
    return ddiv(1.0, V);
}

// (plain selects: the filter's operands are finite inside its domain, and fmin / fmax would spend
// instructions on NaN handling)
MRMP_HD double clamp01(double t) { return t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t); }
MRMP_HD double min2(double a, double b) { return a < b ? a : b; }

// squared distance of the point (origin + r) from the segment origin + t*v, t in [0,1]
template <int D>
MRMP_HD double filt_ptseg_sq(const double *r, const double *v, double inv) {
    double dt = dmul(r[0], v[0]);
#pragma unroll
    for (int k = 1; k < D; ++k) dt = dfma(r[k], v[k], dt);
    const double t = clamp01(dmul(dt, inv));
    double p = dfma(-t, v[0], r[0]);
    double s = dmul(p, p);
#pragma unroll
    for (int k = 1; k < D; ++k) {
        p = dfma(-t, v[k], r[k]);
        s = dfma(p, p, s);
    }
    return s;
}

template <int D>
MRMP_HD int segseg_filter(const double *p11, const double *p12, const double *p21,
                          const double *p22, double inv1, double inv2, double thr) {
    if (!((inv1 >= 0.0) & (inv2 >= 0.0) & (thr >= kFilterThrMin))) return -1;
    double v1[D], v2[D], w[D], g0[D], ng2[D], nw[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        v1[k] = dsub(p12[k], p11[k]);
        v2[k] = dsub(p22[k], p21[k]);
        w[k] = dsub(p21[k], p11[k]);   // p21 seen from p11
        ng2[k] = dsub(p22[k], p11[k]); // p22 seen from p11
        g0[k] = dsub(p12[k], p21[k]);  // p12 seen from p21
        nw[k] = -w[k];                 // p11 seen from p21
    }
    const bool degen = (inv1 == 0.0) | (inv2 == 0.0);
    bool well;
    if constexpr (D == 2) {
        const double cr = dfma(v1[0], v2[1], -dmul(v1[1], v2[0]));
        const double c1 = dfma(v1[0], w[1], -dmul(v1[1], w[0]));
        const double c2 = dfma(v2[0], w[1], -dmul(v2[1], w[0]));
        well = dmul(dmul(cr, cr), dmul(inv1, inv2)) > kFilterSin2Min;
        // proper crossing: the end points of each segment lie on different sides of the other
        const bool cross = (dmul(c1, dadd(c1, cr)) <= 0.0) & (dmul(c2, dadd(c2, cr)) <= 0.0);
        if (well & cross) return 1;  // d* = 0, thr >= 2^-18
    } else {
        // closest points of the two LINES: interior of both segments => the reference may take
        // :76-84, whose value the filter does not model in 3-D
        double V1 = dmul(v1[0], v1[0]), V2 = dmul(v2[0], v2[0]), Dv = dmul(v1[0], v2[0]);
        double D1 = dmul(w[0], v1[0]), D2 = dmul(w[0], v2[0]);
#pragma unroll
        for (int k = 1; k < D; ++k) {
            V1 = dfma(v1[k], v1[k], V1);
            V2 = dfma(v2[k], v2[k], V2);
            Dv = dfma(v1[k], v2[k], Dv);
            D1 = dfma(w[k], v1[k], D1);
            D2 = dfma(w[k], v2[k], D2);
        }
        const double VV = dmul(V1, V2);
        const double Dd = dfma(-Dv, Dv, VV);
        well = Dd > dmul(kFilterSin2Min, VV);
        const double n1 = dfma(D1, V2, -dmul(D2, Dv));
        const double n2 = dfma(D1, Dv, -dmul(D2, V1));
        const double slack = dmul(Dd, 0.0009765625);  // 2^-10
        const bool inner = (n1 >= -slack) & (n1 <= dadd(Dd, slack)) & (n2 >= -slack) &
                           (n2 <= dadd(Dd, slack));
        if (well & inner) return -1;
    }
    if (!(well | degen)) return -1;
    const double s0 = filt_ptseg_sq<D>(w, v1, inv1);    // p21 against segment 1
    const double s1 = filt_ptseg_sq<D>(ng2, v1, inv1);  // p22 against segment 1
    const double s2 = filt_ptseg_sq<D>(nw, v2, inv2);   // p11 against segment 2
    const double s3 = filt_ptseg_sq<D>(g0, v2, inv2);   // p12 against segment 2
    const double s = min2(min2(s0, s1), min2(s2, s3));
    if (!(s == s)) return -1;  // cannot happen inside the domain; never decide on a NaN
    const double lo = dsub(thr, kFilterBand), hi = dadd(thr, kFilterBand);
    if (s < dmul(lo, lo)) return 1;
    if (s > dmul(hi, hi)) return 0;
    return -1;
}

// half length of an edge, rounded up (broad-phase reach)
template <int D>
MRMP_HD double half_len_up(const double *a, const double *b) {
    double v[D];
#pragma unroll
    for (int x = 0; x < D; ++x) v[x] = b[x] - a[x];
    double s = v[0] * v[0];
#pragma unroll
    for (int x = 1; x < D; ++x) s += v[x] * v[x];
    return 0.5 * sqrt(s) * (1.0 + 1e-9);
}

// ================================================================================================
// Conservative FP32 broad phase of the conflict-table kernel
// ================================================================================================
// bound32_far() answers "these two segments are certainly farther apart than r_i + r_j + 2^-20"
// from single-precision copies of the row (midpoint m, unit direction u, half length h, reach
// rc = r_i + 2^-20) and of the column (midpoint, reach h_c + r_j), using the capsule bound
//   dist(seg_r, seg_c) >= dist(m_c, seg_r) - h_c,
//   dist(m_c, seg_r)^2 = |w_perp|^2 + max(|w_par| - h_r, 0)^2,   w = m_c - m_r.
// Only pairs it does not reject reach the FP64 stages, so a wrong "far" would change a verdict:
// the test must be one-sided.  Domain: |coordinate| <= 2 (callers pass rc = +inf / h = +inf for
// anything else, which never tests far).  Error budget of d2 (eps = 2^-24): midpoints rounded to
// float 2*eps each, |dw| <= 4.8e-7 per component, d(w2) <= 2|w||dw| + 3 eps w2, d(par) <=
// |dw| + |w| eps + 3 eps |w|, d(perp2) <= d(w2) + 2|par| d(par), d(ex^2) <= 2 ex (d(par) + eps
// |par|); the sum stays below 2^-20 + 2^-16 * w2 for every |w| <= 5.7 (at |w| = 0.05 / 0.2 / 1 /
// 5.7 the terms add up to 4e-7 / 1.2e-6 / 6e-6 / 7.3e-5 against 1.0e-6 / 1.6e-6 / 1.6e-5 /
// 4.9e-4).  h_r, rc, h_c are rounded UP on conversion and the reach is accumulated with
// round-up instructions, so the right-hand side only grows.  NaN compares false: never far.
struct RowB32 {
    float m[3], u[3], h, rc;
};

MRMP_HD float f32_up(double x) {
#ifdef __CUDA_ARCH__
    return __double2float_ru(x);
#else
    float f = (float)x;
    return ((double)f < x) ? nextafterf(f, INFINITY) : f;
#endif
}
MRMP_HD float fadd_up(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_ru(a, b);
#else
    return nextafterf(a + b, INFINITY);
#endif
}
MRMP_HD float fma_up(float a, float b, float c) {
#ifdef __CUDA_ARCH__
    return __fmaf_ru(a, b, c);
#else
    return nextafterf(fmaf(a, b, c), INFINITY);
#endif
}

constexpr float kB32Abs = 9.5367431640625e-07f;  // 2^-20
constexpr float kB32Rel = 1.52587890625e-05f;    // 2^-16

// (cx, cy, cz, ch) = column midpoint and reach (half length + agent radius, rounded up)
template <int D>
MRMP_HD bool bound32_far(const RowB32 &r, float cx, float cy, float cz, float ch) {
    const float w0 = cx - r.m[0], w1 = cy - r.m[1];
    float w2 = fmaf(w1, w1, w0 * w0);
    float par = fmaf(w1, r.u[1], w0 * r.u[0]);
    if constexpr (D == 3) {
        const float wz = cz - r.m[2];
        w2 = fmaf(wz, wz, w2);
        par = fmaf(wz, r.u[2], par);
    }
    const float ex = fmaxf(fabsf(par) - r.h, 0.0f);
    float d2 = fmaf(ex, ex, fmaf(-par, par, w2));
    d2 = fmaf(-kB32Rel, w2, d2);
    const float reach = fadd_up(r.rc, ch);
    const float rr = fma_up(reach, reach, kB32Abs);
    return d2 > rr;
}

// row side of the bound from the FP64 edge; `in_domain` false => never far
template <int D>
MRMP_HD RowB32 bound32_row(const double *a, const double *b, double reach_r, bool in_domain) {
    RowB32 r;
    double len2 = 0.0, v[D];
#pragma unroll
    for (int x = 0; x < D; ++x) {
        v[x] = b[x] - a[x];
        len2 += v[x] * v[x];
    }
    const double len = sqrt(len2);
    const bool dir = len2 > 0.0 && len2 < 1e300;
    const double invl = dir ? 1.0 / len : 0.0;
    r.m[2] = 0.0f;
    r.u[2] = 0.0f;
#pragma unroll
    for (int x = 0; x < D; ++x) {
        r.m[x] = (float)((a[x] + b[x]) * 0.5);
        r.u[x] = (float)(v[x] * invl);  // a zero-length edge keeps u = 0: disc of the midpoint
    }
    r.h = f32_up(0.5 * len * (1.0 + 1e-9));
    r.rc = in_domain ? f32_up(reach_r) : INFINITY;
    return r;
}

}  // namespace mrmp
