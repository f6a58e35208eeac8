// models.cuh -- compile-time model policies used by the roadmap / expansion kernels.
// A policy exposes the reference's per-model contract (src/MRMP.jl:40-54): state
// dimension, dist, connect(q,i), connect(q_from,q_to,i), sampler scaling.
#pragma once
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
    // dist(qa, qb) as a value
    static __device__ __forceinline__ double dist(const CtxView &, const double *qa,
                                                  const double *qb) {
        double v[D];
#pragma unroll
        for (int k = 0; k < D; ++k) v[k] = dsub(qa[k], qb[k]);
        return norm_val<D>(v);
    }
};

}  // namespace mrmp
