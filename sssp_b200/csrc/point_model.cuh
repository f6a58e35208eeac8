// point_model.cuh -- device bodies of the StatePoint2D / StatePoint3D closures.
// Reference: src/models/point2d.jl:42-69, point3d.jl:47-80 (gen_connect),
//            src/models/point.jl:5-36 (gen_collide).
#pragma once
#include "mrmp_internal.cuh"

namespace mrmp {

// Accessor over the (possibly smem-staged) connect segment of the ctx blob.
struct ConSeg {
    const double *s;  // points at element con_off of the blob (or its smem copy)
    int obs, rads, t2, g, mp, n_obs;
    __device__ __forceinline__ ConSeg(const CtxView &cv, const double *seg)
        : s(seg),
          obs(cv.obs_off - cv.con_off),
          rads(cv.rads_off - cv.con_off),
          t2(cv.t2_obs_off - cv.con_off),
          g(cv.g_obs_off - cv.con_off),
          mp(cv.mp),
          n_obs(cv.n_obs) {}
    __device__ __forceinline__ double rad(int i) const { return s[rads + i]; }
    __device__ __forceinline__ double oc(int k, int m) const { return s[obs + k * mp + m]; }
    __device__ __forceinline__ double t2_im(int i, int m) const { return s[t2 + i * mp + m]; }
    __device__ __forceinline__ double g_im(int i, int m) const { return s[g + i * mp + m]; }
};

// Accessor over the collide segment.
struct ColSeg {
    const double *s;
    int rads, t2, g, n, has_tab;
    __device__ __forceinline__ ColSeg(const CtxView &cv, const double *seg)
        : s(seg),
          rads(cv.crads_off - cv.col_off),
          t2(cv.t2_pair_off - cv.col_off),
          g(cv.g_pair_off - cv.col_off),
          n(cv.n_agents),
          has_tab(cv.has_pair_tab) {}
    __device__ __forceinline__ double rad(int i) const { return s[rads + i]; }
    // rads[i] + rads[j]  (point.jl:11)
    __device__ __forceinline__ double thr(int i, int j) const { return dadd(rad(i), rad(j)); }
    __device__ __forceinline__ double t2_ij(int i, int j) const { return s[t2 + i * n + j]; }
    __device__ __forceinline__ double g_ij(int i, int j) const { return s[g + i * n + j]; }
};

// ---- conservative culling (kernels_cross.cu, kernels_prm.cu) -----------------------------
// A segment whose midpoint is farther from a point (or another midpoint) than
//   thr + kPad + half length(s)
// cannot produce `dist < thr` in the reference's arithmetic: the true distance then exceeds
// thr by kPad = 2^-20, while the reference's computed distance is within a few ulp of the
// coordinate magnitude (< 2^-28 for |coordinates| < 2^20) of the true one.  Callers disable
// the cull for larger or non-finite coordinates.
constexpr double kPad = 9.5367431640625e-07;  // 2^-20
constexpr double kCullCoordMax = 1048576.0;   // 2^20

template <int D>
__device__ __forceinline__ void load_state(const double *__restrict__ base, int64_t k, double *q) {
    if constexpr (D == 2) {
        const double2 t = __ldg(reinterpret_cast<const double2 *>(base) + k);
        q[0] = t.x;
        q[1] = t.y;
    } else {
#pragma unroll
        for (int c = 0; c < D; ++c) q[c] = __ldg(base + (int64_t)D * k + c);
    }
}

// any(x -> (x < rads[i] || 1 - rads[i] < x), q)   point2d.jl:51,60 / point3d.jl:55,67
template <int D>
__device__ __forceinline__ bool out_of_field(const double *q, double r) {
    const double hi = dsub(1.0, r);
    bool bad = false;
#pragma unroll
    for (int k = 0; k < D; ++k) bad |= (q[k] < r) | (hi < q[k]);
    return bad;
}

// connect(q, i): point2d.jl:49-57, point3d.jl:54-64
template <int D>
__device__ __forceinline__ bool point_connect_point(const ConSeg &cs, const double *q, int i) {
    const double r = cs.rad(i);
    bool ok = !out_of_field<D>(q, r);
    for (int m = 0; m < cs.n_obs; ++m) {
        double v[D];
#pragma unroll
        for (int k = 0; k < D; ++k) v[k] = dsub(q[k], cs.oc(k, m));
        const double s = sumsq<D>(v);
        bool hit;
        if (s == 0.0)
            hit = norm_slow<D>(v) < dadd(cs.oc(D, m), r);  // dist(q,o) < o.r + rads[i]
        else
            hit = s < cs.t2_im(i, m);
        ok &= !hit;
    }
    return ok;
}

// connect(q_from, q_to, i): point2d.jl:59-66, point3d.jl:66-77
template <int D>
__device__ __forceinline__ bool point_connect_edge(const ConSeg &cs, const double *qf,
                                                   const double *qt, int i) {
    const double r = cs.rad(i);
    bool ok = !out_of_field<D>(qt, r);  // q_to only
    double u[D];
#pragma unroll
    for (int k = 0; k < D; ++k) u[k] = dsub(qf[k], qt[k]);
    const double uu = dotd<D>(u, u);
    for (int m = 0; m < cs.n_obs; ++m) {
        double c[D];
#pragma unroll
        for (int k = 0; k < D; ++k) c[k] = cs.oc(k, m);
        // dist(q_from,q_to,o) < o.r + rads[i]   (branchy form: most lanes of a list kernel take
        // an end-point case, which skips the division)
        ok &= !segpt_lt<D, true>(qf, qt, u, uu, c, dadd(cs.oc(D, m), r), cs.t2_im(i, m));
    }
    return ok;
}

// collide(qi_from, qi_to, qj_from, qj_to, i, j): point.jl:9-12
template <int D>
__device__ __forceinline__ bool point_collide_pair(const ColSeg &cs, const double *qif,
                                                   const double *qit, const double *qjf,
                                                   const double *qjt, int i, int j) {
    const double thr = cs.thr(i, j);
    if (cs.has_tab) return segseg_lt<D, true>(qif, qit, qjf, qjt, thr, cs.t2_ij(i, j));
    return segseg_lt<D, false>(qif, qit, qjf, qjt, thr, 0.0);
}

// collide(q_i, q_j, i, j): point.jl:14-16
template <int D>
__device__ __forceinline__ bool point_collide_static(const ColSeg &cs, const double *qi,
                                                     const double *qj, int i, int j) {
    double v[D];
#pragma unroll
    for (int k = 0; k < D; ++k) v[k] = dsub(qi[k], qj[k]);
    const double thr = cs.thr(i, j);
    if (cs.has_tab) return norm_lt<D, true>(v, sumsq<D>(v), thr, cs.t2_ij(i, j));
    return norm_lt<D, false>(v, sumsq<D>(v), thr, 0.0);
}

// one term of collide(Q, q_to, i): dist(q_from, q_to, Q[j].q) < rads[i] + rads[j]  point.jl:30
template <int D>
__device__ __forceinline__ bool point_collide_move_vs_static(const ColSeg &cs, const double *qf,
                                                             const double *qt, const double *u,
                                                             double uu, const double *qj, int i,
                                                             int j) {
    const double thr = cs.thr(i, j);
    if (cs.has_tab) return segpt_lt<D, true>(qf, qt, u, uu, qj, thr, cs.t2_ij(i, j));
    return segpt_lt<D, false>(qf, qt, u, uu, qj, thr, 0.0);
}

}  // namespace mrmp
