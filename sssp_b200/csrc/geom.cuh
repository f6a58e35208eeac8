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

}  // namespace mrmp
