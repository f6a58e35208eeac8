// chain_model.cuh -- the remaining discretised models of the reference as instances of one
// template: line2d, snake2d, arm33, capsule3d (src/models/line2d.jl, snake2d.jl, arm33.jl,
// capsule3d.jl; reference card: SURVEY appendix B).
//
// A body is a chain of straight links between P joint points in a W-dimensional workspace.
// A move is discretised like arm22's (arm_model.cuh: ArmSched, Julia's 0:step:D): at every
// step positions are interpolated as (1-e)*from + e*to and angles as
// from + e*diff_angles(to, from) (line2d.jl:53-55), the joint points are rebuilt and tested.
// Everything is __host__ __device__ (tests/hostprobe runs it on the CPU against the checker).
#pragma once
#include "arm_model.cuh"

namespace mrmp {

constexpr int kModelLine2D = 3, kModelSnake2D = 4, kModelArm33 = 5, kModelCapsule3D = 6;
constexpr int kChainMaxPoints = 5;

template <int MODEL>
struct ChainTraits;
template <>
struct ChainTraits<kModelLine2D> {  // line2d.jl:3-7: (x, y, theta), one link of length rads[i]
    static constexpr int SD = 3, W = 2, P = 2, NPOS = 2;
    static constexpr bool kObsAddsRad = false, kPairRadSum = false, kHasBase = false;
};
template <>
struct ChainTraits<kModelSnake2D> {  // snake2d.jl:5-12: (x, y, theta1..4), four links
    static constexpr int SD = 6, W = 2, P = 5, NPOS = 2;
    static constexpr bool kObsAddsRad = false, kPairRadSum = false, kHasBase = false;
};
template <>
struct ChainTraits<kModelArm33> {  // arm33.jl:3-10: (theta, phi) x 3, fixed 3-D base
    static constexpr int SD = 6, W = 3, P = 4, NPOS = 0;
    static constexpr bool kObsAddsRad = false, kPairRadSum = false, kHasBase = true;
};
template <>
struct ChainTraits<kModelCapsule3D> {  // capsule3d.jl:5-12: (x, y, z, roll, pitch, yaw)
    static constexpr int SD = 6, W = 3, P = 2, NPOS = 3;
    static constexpr bool kObsAddsRad = true, kPairRadSum = true, kHasBase = false;
};

struct ChainGeom {   // per-agent constants
    double base[3];  // arm33: positions[i]
    double len;      // link length: rads[i]; capsule3d: axises[i]
    double rad;      // rads[i] (capsule radius / bounds margin of line2d's point form)
};

struct ChainObs {  // obstacle SoA + exact-compare table row of this agent
    const double *s;
    int obs, t2, mp, n_obs;
    MRMP_HD double oc(int k, int m) const { return s[obs + k * mp + m]; }
    MRMP_HD double T2(int m) const { return s[t2 + m]; }
};

struct ChainPar {  // what a check needs from the ctx
    StepRat sr;
    double step, safety, safety_t2, max_dist;
};

template <int MODEL>
struct Chain {
    using T = ChainTraits<MODEL>;
    static constexpr int SD = T::SD, W = T::W, P = T::P, L = T::P - 1, NPOS = T::NPOS;

    // joint points of the posture with (interpolated) state s
    static MRMP_HD void points(const double *s, const ChainGeom &g, double *pts) {
        if constexpr (MODEL == kModelLine2D) {  // line2d.jl:32-33 / :52-57
            double sn, cs;
            sincos_rn(s[2], sn, cs);
            pts[0] = s[0];
            pts[1] = s[1];
            pts[2] = dadd(pts[0], dmul(g.len, cs));
            pts[3] = dadd(pts[1], dmul(g.len, sn));
        } else if constexpr (MODEL == kModelSnake2D) {  // snake2d.jl:36-44 / :105-107
            pts[0] = s[0];
            pts[1] = s[1];
#pragma unroll
            for (int k = 1; k < 5; ++k) {
                double sn, cs;
                sincos_rn(s[1 + k], sn, cs);
                pts[2 * k] = dadd(pts[2 * k - 2], dmul(g.len, cs));
                pts[2 * k + 1] = dadd(pts[2 * k - 1], dmul(g.len, sn));
            }
        } else if constexpr (MODEL == kModelArm33) {  // arm33.jl:34-40 / :83-96
#pragma unroll
            for (int x = 0; x < 3; ++x) pts[x] = g.base[x];
#pragma unroll
            for (int k = 1; k < 4; ++k) {
                double st, ct, sp, cp;
                sincos_rn(s[2 * (k - 1)], st, ct);
                sincos_rn(s[2 * (k - 1) + 1], sp, cp);
                pts[3 * k] = dadd(pts[3 * k - 3], dmul(g.len, dmul(st, cp)));
                pts[3 * k + 1] = dadd(pts[3 * k - 2], dmul(g.len, dmul(st, sp)));
                pts[3 * k + 2] = dadd(pts[3 * k - 1], dmul(g.len, ct));
            }
        } else {  // capsule3d.jl:38-59 / :89-93: tip = root + R[:,1] * axes
            double sph, cph, sth, cth;
            sincos_rn(s[3], sph, cph);
            sincos_rn(s[4], sth, cth);
#pragma unroll
            for (int x = 0; x < 3; ++x) pts[x] = s[x];
            pts[3] = dadd(pts[0], dmul(dmul(cph, cth), g.len));
            pts[4] = dadd(pts[1], dmul(dmul(sph, cth), g.len));
            pts[5] = dadd(pts[2], dmul(-sth, g.len));
        }
    }

    // components of dist(q1, q2): raw position differences, angles through diff_angles / pi
    static MRMP_HD void dist_vec(const double *q1, const double *q2, double *v) {
#pragma unroll
        for (int k = 0; k < NPOS; ++k) v[k] = dsub(q1[k], q2[k]);
#pragma unroll
        for (int k = NPOS; k < SD; ++k) v[k] = ddiv(diff_angles(q1[k], q2[k]), kPi64);
    }
    static MRMP_HD double dist(const double *q1, const double *q2) {
        double v[SD];
        dist_vec(q1, q2, v);
        return norm_val<SD>(v);
    }
    static MRMP_HD void mid(const double *p, const double *q, double *out) {
#pragma unroll
        for (int k = 0; k < NPOS; ++k) out[k] = ddiv(dadd(p[k], q[k]), 2.0);
#pragma unroll
        for (int k = NPOS; k < SD; ++k) out[k] = dadd(ddiv(diff_angles(q[k], p[k]), 2.0), p[k]);
    }

    // bounds margin: capsule3d always rads[i] (capsule3d.jl:74,96); line2d rads[i] in the
    // point form (line2d.jl:35) but none in the edge loop (line2d.jl:60); others none
    static MRMP_HD double margin(const ChainGeom &g, bool point_form) {
        if constexpr (MODEL == kModelCapsule3D) return g.rad;
        if constexpr (MODEL == kModelLine2D) return point_form ? g.rad : 0.0;
        return 0.0;
    }

    static MRMP_HD bool self_hit(const double *p, double safety, double safety_t2) {
        if constexpr (MODEL == kModelSnake2D) {  // check_self_collision_snake2d, snake2d.jl:46-59
            bool hit = segseg_lt<2, true>(p, p + 2, p + 4, p + 6, safety, safety_t2);
            hit |= segseg_lt<2, true>(p, p + 2, p + 6, p + 8, safety, safety_t2);
            hit |= segseg_lt<2, true>(p + 2, p + 4, p + 6, p + 8, safety, safety_t2);
            double u[2] = {dsub(p[4], p[6]), dsub(p[5], p[7])};
            hit |= segpt_lt<2, true>(p + 4, p + 6, u, dotd<2>(u, u), p + 8, safety, safety_t2);
            return hit;
        } else if constexpr (MODEL == kModelArm33) {  // arm33.jl:64,109
            return segseg_lt<3, true>(p, p + 3, p + 6, p + 9, safety, safety_t2);
        } else {
            return false;
        }
    }

    // posture tests of connect; true when free
    static MRMP_HD bool posture_free(const ChainPar &cp, const ChainObs &ob, const ChainGeom &g,
                                     const double *pts, bool point_form) {
        const double lo = margin(g, point_form), hi = dsub(1.0, lo);
        bool bad = false;
#pragma unroll
        for (int k = 0; k < P * W; ++k) bad |= (pts[k] < lo) | (hi < pts[k]);
#pragma unroll
        for (int k = 1; k < P; ++k) {
            const double *a = pts + (k - 1) * W, *b = pts + k * W;
            double u[W];
#pragma unroll
            for (int x = 0; x < W; ++x) u[x] = dsub(a[x], b[x]);
            const double uu = dotd<W>(u, u);
            for (int m = 0; m < ob.n_obs; ++m) {
                double o[W];
#pragma unroll
                for (int x = 0; x < W; ++x) o[x] = ob.oc(x, m);
                const double thr = T::kObsAddsRad ? dadd(ob.oc(W, m), g.rad) : ob.oc(W, m);
                bad |= segpt_lt<W, true>(a, b, u, uu, o, thr, ob.T2(m));
            }
        }
        bad |= self_hit(pts, cp.safety, cp.safety_t2);
        return !bad;
    }

    // link pair lp = (link of i) * L + (link of j) closer than thr?
    static MRMP_HD bool link_pair_hit(int lp, const double *pi, const double *pj, double thr,
                                      double T2) {
        const int ki = lp / L, kj = lp - ki * L;
        return segseg_lt<W, true>(pi + ki * W, pi + (ki + 1) * W, pj + kj * W, pj + (kj + 1) * W, thr,
                                  T2);
    }
    static MRMP_HD bool links_hit(const double *pi, const double *pj, double thr, double T2) {
        bool hit = false;
#pragma unroll
        for (int lp = 0; lp < L * L; ++lp) hit |= link_pair_hit(lp, pi, pj, thr, T2);
        return hit;
    }
};

// One agent's move.
template <int MODEL>
struct ChainMove {
    using C = Chain<MODEL>;
    double from[C::SD], tgt[C::SD];  // tgt: `to` for positions, diff_angles(to, from) for angles
    ChainGeom g;
    ArmSched sch;
    MRMP_HD void state_at(long long k, double *s) const {
        const double e = sch.e(k);
        const double f = dsub(1.0, e);
#pragma unroll
        for (int c = 0; c < C::NPOS; ++c) s[c] = dadd(dmul(f, from[c]), dmul(e, tgt[c]));
#pragma unroll
        for (int c = C::NPOS; c < C::SD; ++c) s[c] = dadd(from[c], dmul(e, tgt[c]));
    }
    MRMP_HD void posture(long long k, double *pts) const {
        double s[C::SD];
        state_at(k, s);
        C::points(s, g, pts);
    }
};

template <int MODEL>
MRMP_HD ChainMove<MODEL> chain_move(const ChainPar &cp, const double *qf, const double *qt,
                                    const ChainGeom &g) {
    using C = Chain<MODEL>;
    ChainMove<MODEL> m;
#pragma unroll
    for (int c = 0; c < C::SD; ++c) m.from[c] = qf[c];
#pragma unroll
    for (int c = 0; c < C::NPOS; ++c) m.tgt[c] = qt[c];
#pragma unroll
    for (int c = C::NPOS; c < C::SD; ++c) m.tgt[c] = diff_angles(qt[c], qf[c]);
    m.g = g;
    m.sch = arm_schedule(cp.sr, cp.step, C::dist(qf, qt));
    return m;
}

// ---- whole-check bodies, one thread per check (roadmap kernels, hostprobe) -------------------
template <int MODEL>
MRMP_HD bool chain_connect_point(const ChainPar &cp, const ChainObs &ob, const ChainGeom &g,
                                 const double *q) {
    double pts[kChainMaxPoints * 3];
    Chain<MODEL>::points(q, g, pts);
    return Chain<MODEL>::posture_free(cp, ob, g, pts, true);
}

template <int MODEL>
MRMP_HD bool chain_connect_edge(const ChainPar &cp, const ChainObs &ob, const ChainGeom &g,
                                const double *qf, const double *qt) {
    const ChainMove<MODEL> mv = chain_move<MODEL>(cp, qf, qt, g);
    if (cp.max_dist == cp.max_dist && mv.sch.D > cp.max_dist) return false;
    const long long n = mv.sch.n();
    for (long long k = 0; k < n; ++k) {
        double pts[kChainMaxPoints * 3];
        mv.posture(k, pts);
        if (!Chain<MODEL>::posture_free(cp, ob, g, pts, false)) return false;
    }
    return true;
}

template <int MODEL>
MRMP_HD bool chain_collide_static(const ChainGeom &gi, const ChainGeom &gj, const double *qi,
                                  const double *qj, double thr, double T2) {
    double pi[kChainMaxPoints * 3], pj[kChainMaxPoints * 3];
    Chain<MODEL>::points(qi, gi, pi);
    Chain<MODEL>::points(qj, gj, pj);
    return Chain<MODEL>::links_hit(pi, pj, thr, T2);
}

template <int MODEL>
MRMP_HD bool chain_collide_pair_serial(const ChainPar &cp, const ChainGeom &gi, const ChainGeom &gj,
                                       const double *qif, const double *qit, const double *qjf,
                                       const double *qjt, double thr, double T2) {
    const ChainMove<MODEL> mi = chain_move<MODEL>(cp, qif, qit, gi);
    const ChainMove<MODEL> mj = chain_move<MODEL>(cp, qjf, qjt, gj);
    const long long ni = mi.sch.n(), nj = mj.sch.n();
    for (long long s = 0; s < ni; ++s) {
        double pi[kChainMaxPoints * 3];
        mi.posture(s, pi);
        for (long long t = 0; t < nj; ++t) {
            double pj[kChainMaxPoints * 3];
            mj.posture(t, pj);
            if (Chain<MODEL>::links_hit(pi, pj, thr, T2)) return true;
        }
    }
    return false;
}

}  // namespace mrmp
