// probe.cu -- TEST INFRASTRUCTURE: runs the device-side arithmetic headers of
// sssp_b200/csrc (geom.cuh, trig.cuh, arm_model.cuh are __host__ __device__) on the CPU so
// that `-m "not gpu"` tests can compare exactly the source the kernels are built from with the
// oracle, bit for bit, in the GPU-less build container.  Not part of libmrmp_b200.so.
#include <stdint.h>

#include "../../sssp_b200/csrc/chain_model.cuh"

using namespace mrmp;

extern "C" {

void probe_sincos(double x, double *s, double *c) { sincos_rn(x, *s, *c); }
double probe_atan2(double y, double x) { return atan2_rn(y, x); }
double probe_diff_angles(double a, double b) { return diff_angles(a, b); }

int probe_segseg_lt(const double *p11, const double *p12, const double *p21, const double *p22,
                    int d, double thr, double T2) {
    return d == 2 ? segseg_lt<2, true>(p11, p12, p21, p22, thr, T2)
                  : segseg_lt<3, true>(p11, p12, p21, p22, thr, T2);
}
int probe_segseg_lt_sqrt(const double *p11, const double *p12, const double *p21,
                         const double *p22, int d, double thr) {
    return d == 2 ? segseg_lt<2, false>(p11, p12, p21, p22, thr, 0.0)
                  : segseg_lt<3, false>(p11, p12, p21, p22, thr, 0.0);
}
int probe_segpt_lt(const double *a, const double *b, const double *c, int d, double thr,
                   double T2) {
    double u[3];
    for (int k = 0; k < d; ++k) u[k] = mrmp::dsub(a[k], b[k]);
    const int v = d == 2 ? segpt_lt<2, true>(a, b, u, dotd<2>(u, u), c, thr, T2)
                         : segpt_lt<3, true>(a, b, u, dotd<3>(u, u), c, thr, T2);
    // the branch-free form the SIMT callers use must give the same verdict: -1 flags a mismatch
    const int w = d == 2 ? segpt_lt_simt<2, true>(a, b, u, dotd<2>(u, u), c, thr, T2)
                         : segpt_lt_simt<3, true>(a, b, u, dotd<3>(u, u), c, thr, T2);
    return v == w ? v : -1;
}

// batch forms (thr / T2 looked up per pair from N x N tables, like the kernels do)
void probe_segseg_batch(int64_t n, int d, const double *a, const double *b, const double *c,
                        const double *e, const int32_t *ai, const int32_t *aj, int N,
                        const double *thr, const double *T2, uint8_t *out) {
    for (int64_t k = 0; k < n; ++k) {
        const double t = thr[ai[k] * N + aj[k]], t2 = T2[ai[k] * N + aj[k]];
        out[k] = d == 2 ? segseg_lt<2, true>(a + 2 * k, b + 2 * k, c + 2 * k, e + 2 * k, t, t2)
                        : segseg_lt<3, true>(a + 3 * k, b + 3 * k, c + 3 * k, e + 3 * k, t, t2);
    }
}
void probe_segpt_batch(int64_t n, int d, const double *a, const double *b, const double *c,
                       const int32_t *ai, const int32_t *aj, int N, const double *thr,
                       const double *T2, uint8_t *out) {
    for (int64_t k = 0; k < n; ++k) {
        const double t = thr[ai[k] * N + aj[k]], t2 = T2[ai[k] * N + aj[k]];
        double u[3];
        for (int x = 0; x < d; ++x) u[x] = mrmp::dsub(a[d * k + x], b[d * k + x]);
        out[k] = d == 2 ? segpt_lt<2, true>(a + 2 * k, b + 2 * k, u, dotd<2>(u, u), c + 2 * k, t, t2)
                        : segpt_lt<3, true>(a + 3 * k, b + 3 * k, u, dotd<3>(u, u), c + 3 * k, t, t2);
        const uint8_t w =
            d == 2 ? segpt_lt_simt<2, true>(a + 2 * k, b + 2 * k, u, dotd<2>(u, u), c + 2 * k, t, t2)
                   : segpt_lt_simt<3, true>(a + 3 * k, b + 3 * k, u, dotd<3>(u, u), c + 3 * k, t, t2);
        if (w != out[k]) out[k] = 255;  // branch-free form disagrees: poison the verdict
    }
}

// verdict filter of the conflict-table kernel (geom.cuh: segseg_filter): out[k] = 1 / 0 / -1
// (undecided); exact[k] = segseg_lt.  thr per pair.
void probe_segseg_filter_batch(int64_t n, int d, const double *a, const double *b, const double *c,
                               const double *e, const double *thr, const double *T2, int8_t *out,
                               uint8_t *exact) {
    for (int64_t k = 0; k < n; ++k) {
        if (d == 2) {
            const double i1 = seg_filter_inv<2>(a + 2 * k, b + 2 * k),
                         i2 = seg_filter_inv<2>(c + 2 * k, e + 2 * k);
            out[k] = (int8_t)segseg_filter<2>(a + 2 * k, b + 2 * k, c + 2 * k, e + 2 * k, i1, i2, thr[k]);
            exact[k] = segseg_lt<2, true>(a + 2 * k, b + 2 * k, c + 2 * k, e + 2 * k, thr[k], T2[k]);
        } else {
            const double i1 = seg_filter_inv<3>(a + 3 * k, b + 3 * k),
                         i2 = seg_filter_inv<3>(c + 3 * k, e + 3 * k);
            out[k] = (int8_t)segseg_filter<3>(a + 3 * k, b + 3 * k, c + 3 * k, e + 3 * k, i1, i2, thr[k]);
            exact[k] = segseg_lt<3, true>(a + 3 * k, b + 3 * k, c + 3 * k, e + 3 * k, thr[k], T2[k]);
        }
    }
}

// FP32 broad phase of the conflict-table kernel (geom.cuh: bound32_far) exactly as the kernel
// feeds it: out[k] = 1 when the pair is rejected as "certainly no collision"
void probe_bound32_batch(int64_t n, int d, const double *a, const double *b, const double *c,
                         const double *e, const double *ri, const double *rj, uint8_t *out) {
    const double pad = 9.5367431640625e-07;
    for (int64_t k = 0; k < n; ++k) {
        if (d == 2) {
            const double *pa = a + 2 * k, *pb = b + 2 * k, *pc = c + 2 * k, *pe = e + 2 * k;
            const RowB32 r = bound32_row<2>(pa, pb, ri[k] + pad, seg_filter_inv<2>(pa, pb) >= 0.0);
            const bool cin = seg_filter_inv<2>(pc, pe) >= 0.0;
            const float ch = cin ? f32_up(half_len_up<2>(pc, pe) + rj[k]) : INFINITY;
            out[k] = bound32_far<2>(r, (float)((pc[0] + pe[0]) * 0.5), (float)((pc[1] + pe[1]) * 0.5), 0.0f, ch);
        } else {
            const double *pa = a + 3 * k, *pb = b + 3 * k, *pc = c + 3 * k, *pe = e + 3 * k;
            const RowB32 r = bound32_row<3>(pa, pb, ri[k] + pad, seg_filter_inv<3>(pa, pb) >= 0.0);
            const bool cin = seg_filter_inv<3>(pc, pe) >= 0.0;
            const float ch = cin ? f32_up(half_len_up<3>(pc, pe) + rj[k]) : INFINITY;
            out[k] = bound32_far<3>(r, (float)((pc[0] + pe[0]) * 0.5), (float)((pc[1] + pe[1]) * 0.5),
                                    (float)((pc[2] + pe[2]) * 0.5), ch);
        }
    }
}

// vcat(collect(0:step:D)/D, 1.0): writes min(n, cap) values, returns n
int64_t probe_step_schedule(double step, double D, double *out, int64_t cap) {
    const ArmSched s = arm_schedule(step_rat(step), step, D);
    for (int64_t k = 0; k < s.n() && k < cap; ++k) out[k] = s.e(k);
    return s.n();
}

struct ProbeArm {
    int n_obs;
    const double *ox, *oy, *orad, *t2;  // obstacle SoA + exact-compare table of d < o.r
    double step, safety, safety_t2, max_dist;
};

static void unpack(const ProbeArm *p, double *buf, ArmObs &ob, ArmCtx &ac) {
    // SoA block in the layout ArmObs expects: x[mp] y[mp] r[mp] | t2[mp]
    const int mp = p->n_obs > 0 ? p->n_obs : 1;
    for (int m = 0; m < p->n_obs; ++m) {
        buf[m] = p->ox[m];
        buf[mp + m] = p->oy[m];
        buf[2 * mp + m] = p->orad[m];
        buf[3 * mp + m] = p->t2[m];
    }
    ob.s = buf;
    ob.obs = 0;
    ob.t2 = 3 * mp;
    ob.mp = mp;
    ob.n_obs = p->n_obs;
    ac.sr = step_rat(p->step);
    ac.step = p->step;
    ac.safety = p->safety;
    ac.safety_t2 = p->safety_t2;
    ac.max_dist = p->max_dist;
}

int probe_arm_connect_point(const ProbeArm *p, const double *q, double ax, double ay, double rad) {
    double buf[4 * 64];
    if (p->n_obs > 64) return -1;
    ArmObs ob;
    ArmCtx ac;
    unpack(p, buf, ob, ac);
    return arm_connect_point(ac, ob, q, ax, ay, rad);
}
int probe_arm_connect_edge(const ProbeArm *p, const double *qf, const double *qt, double ax,
                           double ay, double rad) {
    double buf[4 * 64];
    if (p->n_obs > 64) return -1;
    ArmObs ob;
    ArmCtx ac;
    unpack(p, buf, ob, ac);
    return arm_connect_edge(ac, ob, qf, qt, ax, ay, rad);
}
int probe_arm_collide_static(const ProbeArm *p, const double *qi, const double *qj, double aix,
                             double aiy, double ri, double ajx, double ajy, double rj) {
    double buf[4 * 64];
    ArmObs ob;
    ArmCtx ac;
    ProbeArm q = *p;
    q.n_obs = 0;
    unpack(&q, buf, ob, ac);
    return arm_collide_static(ac, qi, qj, aix, aiy, ri, ajx, ajy, rj);
}
int probe_arm_collide_pair(const ProbeArm *p, const double *qif, const double *qit,
                           const double *qjf, const double *qjt, double aix, double aiy, double ri,
                           double ajx, double ajy, double rj) {
    double buf[4 * 64];
    ArmObs ob;
    ArmCtx ac;
    ProbeArm q = *p;
    q.n_obs = 0;
    unpack(&q, buf, ob, ac);
    return arm_collide_pair_serial(ac, qif, qit, qjf, qjt, aix, aiy, ri, ajx, ajy, rj);
}
double probe_arm_dist(const double *q1, const double *q2) { return arm_dist(q1, q2); }
void probe_arm_mid_status(const double *p, const double *q, double *out) {
    arm_mid_status(p, q, out);
}
}  // extern "C"

// ---- line2d / snake2d / arm33 / capsule3d (chain_model.cuh) ------------------------------------
struct ProbeChain {
    int model, n_obs;
    const double *obs_soa;  // [(W+1)][n_obs]: coordinates then radius
    const double *t2;       // [n_obs] exact-compare table of this agent's obstacle threshold
    double step, safety, safety_t2, max_dist;
};

template <int MODEL>
static int chain_dispatch(int what, const ProbeChain *p, const double *g1, const double *g2,
                          const double *a, const double *b, const double *c, const double *d,
                          double thr, double T2, double *out) {
    ChainPar cp;
    cp.sr = step_rat(p->step);
    cp.step = p->step;
    cp.safety = p->safety;
    cp.safety_t2 = p->safety_t2;
    cp.max_dist = p->max_dist;
    ChainObs ob;
    ob.s = p->obs_soa;
    ob.obs = 0;
    ob.mp = p->n_obs > 0 ? p->n_obs : 1;
    ob.n_obs = p->n_obs;
    ob.t2 = 0;
    ChainObs obt = ob;  // T2 lives in a separate array: point the accessor at it
    ChainGeom gi, gj;
    for (int x = 0; x < 3; ++x) {
        gi.base[x] = g1[x];
        gj.base[x] = g2 ? g2[x] : 0.0;
    }
    gi.len = g1[3];
    gi.rad = g1[4];
    gj.len = g2 ? g2[3] : 0.0;
    gj.rad = g2 ? g2[4] : 0.0;
    // ChainObs reads coordinates and T2 through one base pointer: pack both into one buffer
    double buf[5 * 64];
    const int mp = ob.mp;
    for (int k = 0; k <= Chain<MODEL>::W; ++k)
        for (int m = 0; m < p->n_obs; ++m) buf[k * mp + m] = p->obs_soa[k * p->n_obs + m];
    for (int m = 0; m < p->n_obs; ++m) buf[(Chain<MODEL>::W + 1) * mp + m] = p->t2[m];
    obt.s = buf;
    obt.t2 = (Chain<MODEL>::W + 1) * mp;
    switch (what) {
    case 0: return chain_connect_point<MODEL>(cp, obt, gi, a);
    case 1: return chain_connect_edge<MODEL>(cp, obt, gi, a, b);
    case 2: return chain_collide_static<MODEL>(gi, gj, a, b, thr, T2);
    case 3: return chain_collide_pair_serial<MODEL>(cp, gi, gj, a, b, c, d, thr, T2);
    case 4: *out = Chain<MODEL>::dist(a, b); return 0;
    default: Chain<MODEL>::mid(a, b, out); return 0;
    }
}

extern "C" {
// what: 0 connect(q,i) [a], 1 connect(qf,qt,i) [a,b], 2 collide static [a,b], 3 collide 6-arity
// [a,b,c,d], 4 dist [a,b] -> out[0], 5 mid status [a,b] -> out[SD].  g = {base[3], len, rad}.
int probe_chain(int what, const ProbeChain *p, const double *g1, const double *g2, const double *a,
                const double *b, const double *c, const double *d, double thr, double T2,
                double *out) {
    if (p->n_obs > 64) return -1;
    switch (p->model) {
    case kModelLine2D: return chain_dispatch<kModelLine2D>(what, p, g1, g2, a, b, c, d, thr, T2, out);
    case kModelSnake2D: return chain_dispatch<kModelSnake2D>(what, p, g1, g2, a, b, c, d, thr, T2, out);
    case kModelArm33: return chain_dispatch<kModelArm33>(what, p, g1, g2, a, b, c, d, thr, T2, out);
    case kModelCapsule3D: return chain_dispatch<kModelCapsule3D>(what, p, g1, g2, a, b, c, d, thr, T2, out);
    default: return -2;
    }
}
}
