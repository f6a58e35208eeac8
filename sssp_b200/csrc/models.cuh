// models.cuh -- compile-time model policies used by the roadmap / expansion kernels.
// A policy exposes the reference's per-model contract (src/MRMP.jl:40-54): state
// dimension, dist, connect(q,i), connect(q_from,q_to,i), sampler scaling.
#pragma once
#include "arm_model.cuh"
#include "chain_model.cuh"
#include "point_model.cuh"

namespace mrmp {

template <int D>
struct PointModel {
    static constexpr int SD = D;
    // StatePoint2D(rand(2)...) / StatePoint3D(rand(3)...): samples are used as drawn
    static __device__ __forceinline__ void scale_sample(double *) {}
    static __device__ __forceinline__ bool connect_point(const CtxView &, const ConSeg &cs,
                                                         const double *q, int i) {
        return point_connect_point<D>(cs, q, i);
    }
    static __device__ __forceinline__ bool connect_edge(const CtxView &, const ConSeg &cs,
                                                        const double *qf, const double *qt, int i) {
        return point_connect_edge<D>(cs, qf, qt, i);
    }
    // dist(q_from, q_to) <= rad  (point2d.jl:12-14, point3d.jl:13-15; prm.jl:207, sssp.jl:88)
    // r2 = sq_le_threshold(rad) = max{ s : RN(sqrt(s)) <= rad }
    static __device__ __forceinline__ bool dist_le(const CtxView &, const double *qa,
                                                   const double *qb, double rad, double r2) {
        double v[D];
#pragma unroll
        for (int k = 0; k < D; ++k) v[k] = dsub(qa[k], qb[k]);
        const double s = sumsq<D>(v);
        if (s == 0.0) return norm_slow<D>(v) <= rad;
        return s <= r2;
    }
    // s with: dist(qa, qb) <= rad  <=>  s <= sq_le_threshold(rad), for rad >= kGateRadMin
    // (s == 0 then means a distance below 1e-153, see dist_le)
    static __device__ __forceinline__ double dist_sq(const double *qa, const double *qb) {
        double v[D];
#pragma unroll
        for (int k = 0; k < D; ++k) v[k] = dsub(qa[k], qb[k]);
        return sumsq<D>(v);
    }
    // dist(qa, qb) as a value
    static __device__ __forceinline__ double dist(const CtxView &, const double *qa,
                                                  const double *qb) {
        double v[D];
#pragma unroll
        for (int k = 0; k < D; ++k) v[k] = dsub(qa[k], qb[k]);
        return norm_val<D>(v);
    }
    // dist(qa, qb) > thr (false for NaN, like `minimum(...) > thr`, sssp.jl:250);
    // r2 = sq_le_threshold(thr)
    static __device__ __forceinline__ bool dist_gt(const CtxView &, const double *qa,
                                                   const double *qb, double thr, double r2) {
        double v[D];
#pragma unroll
        for (int k = 0; k < D; ++k) v[k] = dsub(qa[k], qb[k]);
        const double s = sumsq<D>(v);
        if (s == 0.0) return norm_slow<D>(v) > thr;
        return s > r2;
    }
    // get_mid_status: point2d.jl:8-10, point3d.jl:9-11
    static __device__ __forceinline__ void mid(const double *p, const double *q, double *out) {
#pragma unroll
        for (int k = 0; k < D; ++k) out[k] = ddiv(dadd(p[k], q[k]), 2.0);
    }
};

// StateArm22 (src/models/arm22.jl): thread-per-check bodies for the roadmap kernels
struct ArmModel {
    static constexpr int SD = 2;
    // StateArm22((rand(2) * 2pi)...)  arm22.jl:214-218
    static __device__ __forceinline__ void scale_sample(double *q) {
        q[0] = dmul(q[0], 6.283185307179586);
        q[1] = dmul(q[1], 6.283185307179586);
    }
    static __device__ __forceinline__ ArmObs obs(const CtxView &cv, const ConSeg &cs) {
        ArmObs o;
        o.s = cs.s;
        o.obs = cs.obs;
        o.t2 = cs.t2;
        o.mp = cs.mp;
        o.n_obs = cs.n_obs;
        (void)cv;
        return o;
    }
    static __device__ __forceinline__ bool connect_point(const CtxView &cv, const ConSeg &cs,
                                                         const double *q, int i) {
        const int base = cv.base_off - cv.con_off;
        return arm_connect_point(arm_ctx(cv), obs(cv, cs), q, cs.s[base + 2 * i],
                                 cs.s[base + 2 * i + 1], cs.rad(i));
    }
    static __device__ __forceinline__ bool connect_edge(const CtxView &cv, const ConSeg &cs,
                                                        const double *qf, const double *qt, int i) {
        const int base = cv.base_off - cv.con_off;
        return arm_connect_edge(arm_ctx(cv), obs(cv, cs), qf, qt, cs.s[base + 2 * i],
                                cs.s[base + 2 * i + 1], cs.rad(i));
    }
    // dist(q_from, q_to) <= rad with dist of arm22.jl:15-20
    static __device__ __forceinline__ bool dist_le(const CtxView &, const double *qa,
                                                   const double *qb, double rad, double r2) {
        double v[2];
        arm_dist_vec(qa, qb, v);
        const double s = sumsq<2>(v);
        if (s == 0.0) return norm_slow<2>(v) <= rad;
        return s <= r2;
    }
    static __device__ __forceinline__ double dist_sq(const double *qa, const double *qb) {
        double v[2];
        arm_dist_vec(qa, qb, v);
        return sumsq<2>(v);
    }
    static __device__ __forceinline__ double dist(const CtxView &, const double *qa,
                                                  const double *qb) {
        return arm_dist(qa, qb);
    }
    static __device__ __forceinline__ bool dist_gt(const CtxView &, const double *qa,
                                                   const double *qb, double thr, double r2) {
        double v[2];
        arm_dist_vec(qa, qb, v);
        const double s = sumsq<2>(v);
        if (s == 0.0) return norm_slow<2>(v) > thr;
        return s > r2;
    }
    // get_mid_status: arm22.jl:8-13
    static __device__ __forceinline__ void mid(const double *p, const double *q, double *out) {
        arm_mid_status(p, q, out);
    }
};

// line2d / snake2d / arm33 / capsule3d: thread-per-check bodies for the roadmap kernels
template <int MODEL>
struct ChainModel {
    using C = Chain<MODEL>;
    static constexpr int SD = C::SD;
    // positions rand(), angles rand() * 2pi (line2d.jl:165-169, snake2d.jl:261-265,
    // arm33.jl:288-292, capsule3d.jl:216-220)
    static __device__ __forceinline__ void scale_sample(double *q) {
#pragma unroll
        for (int k = C::NPOS; k < SD; ++k) q[k] = dmul(q[k], 6.283185307179586);
    }
    static __device__ __forceinline__ ChainPar par(const CtxView &cv) {
        ChainPar p;
        p.sr.ok = cv.step_rat_ok;
        p.sr.num = cv.step_rat_num;
        p.sr.den = cv.step_rat_den;
        p.step = cv.step_dist;
        p.safety = cv.safety_dist;
        p.safety_t2 = cv.safety_t2;
        p.max_dist = cv.max_dist;
        return p;
    }
    static __device__ __forceinline__ ChainGeom geom(const CtxView &cv, const ConSeg &cs, int i) {
        ChainGeom g;
        g.rad = cs.rad(i);
        g.len = MODEL == kModelCapsule3D ? cs.s[cv.aux_off - cv.con_off + i] : g.rad;
#pragma unroll
        for (int x = 0; x < 3; ++x)
            g.base[x] = (ChainTraits<MODEL>::kHasBase && x < C::W)
                            ? cs.s[cv.base_off - cv.con_off + C::W * i + x]
                            : 0.0;
        return g;
    }
    static __device__ __forceinline__ ChainObs obs(const ConSeg &cs, int i) {
        ChainObs o;
        o.s = cs.s;
        o.obs = cs.obs;
        o.t2 = cs.t2 + (ChainTraits<MODEL>::kObsAddsRad ? i * cs.mp : 0);
        o.mp = cs.mp;
        o.n_obs = cs.n_obs;
        return o;
    }
    static __device__ __forceinline__ bool connect_point(const CtxView &cv, const ConSeg &cs,
                                                         const double *q, int i) {
        return chain_connect_point<MODEL>(par(cv), obs(cs, i), geom(cv, cs, i), q);
    }
    static __device__ __forceinline__ bool connect_edge(const CtxView &cv, const ConSeg &cs,
                                                        const double *qf, const double *qt, int i) {
        return chain_connect_edge<MODEL>(par(cv), obs(cs, i), geom(cv, cs, i), qf, qt);
    }
    static __device__ __forceinline__ bool dist_le(const CtxView &, const double *qa,
                                                   const double *qb, double rad, double r2) {
        double v[SD];
        C::dist_vec(qa, qb, v);
        const double s = sumsq<SD>(v);
        if (s == 0.0) return norm_slow<SD>(v) <= rad;
        return s <= r2;
    }
    static __device__ __forceinline__ bool dist_gt(const CtxView &, const double *qa,
                                                   const double *qb, double thr, double r2) {
        double v[SD];
        C::dist_vec(qa, qb, v);
        const double s = sumsq<SD>(v);
        if (s == 0.0) return norm_slow<SD>(v) > thr;
        return s > r2;
    }
    static __device__ __forceinline__ double dist_sq(const double *qa, const double *qb) {
        double v[SD];
        C::dist_vec(qa, qb, v);
        return sumsq<SD>(v);
    }
    static __device__ __forceinline__ double dist(const CtxView &, const double *qa,
                                                  const double *qb) {
        return C::dist(qa, qb);
    }
    static __device__ __forceinline__ void mid(const double *p, const double *q, double *out) {
        C::mid(p, q, out);
    }
};

}  // namespace mrmp
