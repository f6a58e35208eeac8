// arm_model.cuh -- bodies of the StateArm22 closures (src/models/arm22.jl) for sm_100a.
//
// A state is (theta1, theta2); the arm of agent i has its base at positions[i] and two links
// of length rads[i] (arm22.jl:22-27).  Moves are discretised: the reference walks
//     for e in vcat(collect(0:step_dist:D) / D, 1.0)            (arm22.jl:65,117,125,180)
// with D = dist(q_from, q_to) in normalised angle space (arm22.jl:15-20), rebuilds the joints
// at  theta_from + e * diff_angles(theta_to, theta_from)  and tests every posture.
//
// Everything here is __host__ __device__ for the same reason as geom.cuh (tests/hostprobe).
#pragma once
#include "mrmp_internal.cuh"
#include "trig.cuh"

namespace mrmp {

// ---- Julia's  0:step:D  ---------------------------------------------------------------------
// Base.rat (twiceprecision.jl) for Float64: continued-fraction expansion with numerator and
// denominator bounded by 2^24 (the narrow type is Float32).
MRMP_HD void jl_rat(double x, long long &num, long long &den) {
    const double m = 16777216.0;
    double y = x;
    long long a = 1, d = 1, b = 0, c = 0;
    while (fabs(y) <= m) {
        const long long f = (long long)trunc(y);
        y = dsub(y, (double)f);
        const long long na = f * a + c, nb = f * b + d;
        c = a;
        a = na;
        d = b;
        b = nb;
        const long long aa = a < 0 ? -a : a, ab = b < 0 ? -b : b;
        if ((aa > ab ? aa : ab) > 16777216) {
            num = c;
            den = d;
            return;
        }
        if (ddiv((double)a, (double)b) == x) {
            num = a;
            den = b;
            return;
        }
        y = ddiv(1.0, y);
    }
    num = a;
    den = b;
}

// D-independent half of the rational lift of (0.0 : step : D): filled once per ctx on the host
struct StepRat {
    int ok;              // step has a short exact rational and the lift is admissible
    long long num, den;  // step == num / den
};

MRMP_HD StepRat step_rat(double step) {
    StepRat r{0, 0, 0};
    long long sn, sd;
    jl_rat(step, sn, sd);
    if (sd == 0 || ddiv((double)sn, (double)sd) != step) return r;
    const long long den = sd;  // lcm(1, sd)
    if (!(fabs(dmul(step, (double)den)) <= 9007199254740992.0)) return r;
    const long long n = (long long)rint(dmul(step, (double)den));
    if (n == 0) return r;
    r.ok = 1;
    r.num = n;
    r.den = den;
    return r;
}

MRMP_HD bool between(double a, double x, double b) {
    return (a <= x && x <= b) || (b <= x && x <= a);
}

// The interpolation schedule of one move: e(k) for k = 0 .. n-1, where the last one is 1.0.
struct ArmSched {
    long long len;  // length of 0:step:D
    int rational;   // elements are RN(k*num/den) (rational lift) instead of fl(k*step)
    long long num, den;
    double step, D;
    MRMP_HD long long n() const { return len + 1; }
    MRMP_HD double e(long long k) const {
        if (k >= len) return 1.0;
        const double r = rational ? ddiv((double)(k * num), (double)den) : dmul((double)k, step);
        return ddiv(r, D);
    }
};

MRMP_HD ArmSched arm_schedule(const StepRat &sr, double step, double D) {
    ArmSched s;
    s.len = 0;
    s.rational = 0;
    s.num = sr.num;
    s.den = sr.den;
    s.step = step;
    s.D = D;
    if (!(D == D)) return s;  // NaN move: only e = 1 (the reference's behaviour is undefined here)
    if (sr.ok) {
        long long en, ed;
        jl_rat(D, en, ed);
        if (ed != 0 && ddiv((double)en, (double)ed) == D) {
            long long l = (sr.den * en) / (sr.num * ed) + 1;  // div(): truncation
            if (l < 0) l = 0;
            if (between(0.0, dmul((double)(l - 1), step), dadd(D, ddiv(step, 2.0))) &&
                !between(0.0, dmul((double)l, step), D)) {
                s.rational = 1;
                s.len = l;
                return s;
            }
        }
    }
    const double lf = ddiv(D, step);
    if (lf < 0) {
        s.len = 0;
    } else if (lf == 0) {
        s.len = 1;
    } else {
        long long l = (long long)rint(lf) + 1;
        const double stop2 = dmul((double)(l - 1), step);
        if (0.0 < D && D < stop2) l -= 1;
        s.len = l;
    }
    return s;
}

// ---- per-model data of the ctx blob --------------------------------------------------------
struct ArmSeg {  // accessor over the (staged) connect or collide segment
    const double *s;
    int rads, base;
    MRMP_HD ArmSeg(const double *seg, int rads_rel, int base_rel) : s(seg), rads(rads_rel), base(base_rel) {}
    MRMP_HD double rad(int i) const { return s[rads + i]; }
    MRMP_HD double bx(int i) const { return s[base + 2 * i]; }
    MRMP_HD double by(int i) const { return s[base + 2 * i + 1]; }
};

// dist(q1, q2) = norm([diff_angles(q1.t1,q2.t1)/pi, diff_angles(q1.t2,q2.t2)/pi])  arm22.jl:15-20
constexpr double kPi64 = 3.141592653589793;
MRMP_HD void arm_dist_vec(const double *q1, const double *q2, double *v) {
    v[0] = ddiv(diff_angles(q1[0], q2[0]), kPi64);
    v[1] = ddiv(diff_angles(q1[1], q2[1]), kPi64);
}
MRMP_HD double arm_dist(const double *q1, const double *q2) {
    double v[2];
    arm_dist_vec(q1, q2, v);
    return norm_val<2>(v);
}

// joints b, c of the posture (t1, t2): arm22.jl:24-25 / :69-70 / :122-123
MRMP_HD void arm_joints(double t1, double t2, double ax, double ay, double rad, double *b,
                        double *c) {
    double s1, c1, s2, c2;
    sincos_rn(t1, s1, c1);
    sincos_rn(t2, s2, c2);
    b[0] = dadd(ax, dmul(rad, c1));
    b[1] = dadd(ay, dmul(rad, s1));
    c[0] = dadd(b[0], dmul(rad, c2));
    c[1] = dadd(b[1], dmul(rad, s2));
}

// One agent's move: everything the step loop needs.
struct ArmMove {
    double t1, t2, dt1, dt2;  // q_from and the signed shortest turns (arm22.jl:60-61)
    double ax, ay, rad;
    ArmSched sch;
    MRMP_HD void posture(long long k, double *b, double *c) const {  // arm22.jl:66-70
        const double e = sch.e(k);
        arm_joints(dadd(t1, dmul(e, dt1)), dadd(t2, dmul(e, dt2)), ax, ay, rad, b, c);
    }
};

MRMP_HD ArmMove arm_move(const StepRat &sr, double step, const double *qf, const double *qt,
                         double ax, double ay, double rad) {
    ArmMove m;
    m.t1 = qf[0];
    m.t2 = qf[1];
    m.dt1 = diff_angles(qt[0], qf[0]);
    m.dt2 = diff_angles(qt[1], qf[1]);
    m.ax = ax;
    m.ay = ay;
    m.rad = rad;
    m.sch = arm_schedule(sr, step, arm_dist(qf, qt));
    return m;
}

// ---- posture tests -------------------------------------------------------------------------
// Obstacle view for the arm: centres and radii SoA, exact-compare table of  d < o.r
struct ArmObs {
    const double *s;
    int obs, t2, mp, n_obs;
    MRMP_HD double ox(int m) const { return s[obs + m]; }
    MRMP_HD double oy(int m) const { return s[obs + mp + m]; }
    MRMP_HD double orad(int m) const { return s[obs + 2 * mp + m]; }
    MRMP_HD double T2(int m) const { return s[t2 + m]; }
};

// arm22.jl:43-51 / :73-82: true when the posture (a, b, c) is free
MRMP_HD bool arm_posture_free(const ArmObs &ob, const double *a, const double *b, const double *c,
                              double safety, double safety_t2) {
    bool bad = false;
#pragma unroll
    for (int k = 0; k < 2; ++k)
        bad |= (a[k] < 0.0) | (1.0 < a[k]) | (b[k] < 0.0) | (1.0 < b[k]) | (c[k] < 0.0) |
               (1.0 < c[k]);
    double u1[2] = {dsub(a[0], b[0]), dsub(a[1], b[1])};
    double u2[2] = {dsub(b[0], c[0]), dsub(b[1], c[1])};
    const double uu1 = dotd<2>(u1, u1), uu2 = dotd<2>(u2, u2);
    for (int m = 0; m < ob.n_obs; ++m) {
        const double o[2] = {ob.ox(m), ob.oy(m)};
        bad |= segpt_lt<2, true>(a, b, u1, uu1, o, ob.orad(m), ob.T2(m));  // dist(a,b,o) < o.r
        bad |= segpt_lt<2, true>(b, c, u2, uu2, o, ob.orad(m), ob.T2(m));  // dist(b,c,o) < o.r
    }
    bad |= segpt_lt<2, true>(a, b, u1, uu1, c, safety, safety_t2);  // self collision :51,:82
    return !bad;
}

// one of the four link-link tests of arm22.jl:134-142; lp = 2*(link of i) + (link of j)
MRMP_HD bool arm_link_pair_hit(int lp, const double *ai, const double *bi, const double *ci,
                               const double *aj, const double *bj, const double *cj, double safety,
                               double safety_t2) {
    const double *p11 = (lp & 2) ? bi : ai, *p12 = (lp & 2) ? ci : bi;
    const double *p21 = (lp & 1) ? bj : aj, *p22 = (lp & 1) ? cj : bj;
    return segseg_lt<2, true>(p11, p12, p21, p22, safety, safety_t2);
}

MRMP_HD bool arm_links_hit(const double *ai, const double *bi, const double *ci, const double *aj,
                           const double *bj, const double *cj, double safety, double safety_t2) {
    bool hit = false;
#pragma unroll
    for (int lp = 0; lp < 4; ++lp)
        hit |= arm_link_pair_hit(lp, ai, bi, ci, aj, bj, cj, safety, safety_t2);
    return hit;
}

// ---- whole-check bodies, one thread per check (used by the roadmap kernels and hostprobe) ---
struct ArmCtx {  // what a check needs from the ctx
    StepRat sr;
    double step, safety, safety_t2, max_dist;
};

MRMP_HD ArmCtx arm_ctx(const CtxView &cv) {
    ArmCtx a;
    a.sr.ok = cv.step_rat_ok;
    a.sr.num = cv.step_rat_num;
    a.sr.den = cv.step_rat_den;
    a.step = cv.step_dist;
    a.safety = cv.safety_dist;
    a.safety_t2 = cv.safety_t2;
    a.max_dist = cv.max_dist;
    return a;
}

// connect(q, i): arm22.jl:40-54
MRMP_HD bool arm_connect_point(const ArmCtx &ac, const ArmObs &ob, const double *q, double ax,
                               double ay, double rad) {
    const double a[2] = {ax, ay};
    double b[2], c[2];
    arm_joints(q[0], q[1], ax, ay, rad, b, c);
    return arm_posture_free(ob, a, b, c, ac.safety, ac.safety_t2);
}

// connect(q_from, q_to, i): arm22.jl:56-86
MRMP_HD bool arm_connect_edge(const ArmCtx &ac, const ArmObs &ob, const double *qf,
                              const double *qt, double ax, double ay, double rad) {
    const ArmMove mv = arm_move(ac.sr, ac.step, qf, qt, ax, ay, rad);
    if (ac.max_dist == ac.max_dist && mv.sch.D > ac.max_dist) return false;  // :58
    const double a[2] = {ax, ay};
    const long long n = mv.sch.n();
    for (long long k = 0; k < n; ++k) {
        double b[2], c[2];
        mv.posture(k, b, c);
        if (!arm_posture_free(ob, a, b, c, ac.safety, ac.safety_t2)) return false;
    }
    return true;
}

// collide(q_i, q_j, i, j): arm22.jl:148-162
MRMP_HD bool arm_collide_static(const ArmCtx &ac, const double *qi, const double *qj, double aix,
                                double aiy, double ri, double ajx, double ajy, double rj) {
    const double ai[2] = {aix, aiy}, aj[2] = {ajx, ajy};
    double bi[2], ci[2], bj[2], cj[2];
    arm_joints(qi[0], qi[1], aix, aiy, ri, bi, ci);
    arm_joints(qj[0], qj[1], ajx, ajy, rj, bj, cj);
    return arm_links_hit(ai, bi, ci, aj, bj, cj, ac.safety, ac.safety_t2);
}

// collide(qi_from, qi_to, qj_from, qj_to, i, j): arm22.jl:100-146 -- plain double loop (the
// warp-cooperative kernel in kernels_arm.cu is the fast path; this is its serial statement)
MRMP_HD bool arm_collide_pair_serial(const ArmCtx &ac, const double *qif, const double *qit,
                                     const double *qjf, const double *qjt, double aix, double aiy,
                                     double ri, double ajx, double ajy, double rj) {
    const ArmMove mi = arm_move(ac.sr, ac.step, qif, qit, aix, aiy, ri);
    const ArmMove mj = arm_move(ac.sr, ac.step, qjf, qjt, ajx, ajy, rj);
    const double ai[2] = {aix, aiy}, aj[2] = {ajx, ajy};
    const long long ni = mi.sch.n(), nj = mj.sch.n();
    for (long long s = 0; s < ni; ++s) {
        double bi[2], ci[2];
        mi.posture(s, bi, ci);
        for (long long t = 0; t < nj; ++t) {
            double bj[2], cj[2];
            mj.posture(t, bj, cj);
            if (arm_links_hit(ai, bi, ci, aj, bj, cj, ac.safety, ac.safety_t2)) return true;
        }
    }
    return false;
}

// get_mid_status: arm22.jl:8-13
MRMP_HD void arm_mid_status(const double *p, const double *q, double *out) {
    out[0] = dadd(ddiv(diff_angles(q[0], p[0]), 2.0), p[0]);
    out[1] = dadd(ddiv(diff_angles(q[1], p[1]), 2.0), p[1]);
}

}  // namespace mrmp
