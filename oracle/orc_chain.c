/*
 * orc_chain.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * The remaining discretised models of the reference: line2d, snake2d, arm33, capsule3d
 * (src/models/line2d.jl, snake2d.jl, arm33.jl, capsule3d.jl).  All four follow the same
 * pattern as arm22: a body made of straight links between joint points, moves discretised by
 *     for e in vcat(collect(0:step_dist:D) / D, 1.0)
 * with D = dist(q_from, q_to) (positions raw, angles through diff_angles / pi), positions
 * interpolated as (1 - e) * from + e * to and angles as from + e * diff_angles(to, from).
 * The functions below restate, per model, how the joint points are built and which tests a
 * posture / a pair of postures must pass; the step loops are shared.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "mrmp_oracle.h"

#define MAXP 5         /* joint points of the longest body (snake2d) */
#define CH_MAX_STEPS 8192

static const double PI64 = 3.141592653589793; /* Float64(pi) */

typedef struct {
    int sd;   /* state dimension */
    int w;    /* workspace dimension */
    int p;    /* joint points per posture (links = p - 1) */
    int npos; /* leading positional state components */
} chain_desc;

static chain_desc desc_of(int model) {
    switch (model) {
    case ORC_LINE2D: return (chain_desc){3, 2, 2, 2};    /* line2d.jl:3-7 */
    case ORC_SNAKE2D: return (chain_desc){6, 2, 5, 2};   /* snake2d.jl:5-12 */
    case ORC_ARM33: return (chain_desc){6, 3, 4, 0};     /* arm33.jl:3-10 */
    default: return (chain_desc){6, 3, 2, 3};            /* capsule3d.jl:5-12 */
    }
}

int orc_chain_state_dim(int model) { return desc_of(model).sd; }
int orc_chain_workspace_dim(int model) { return desc_of(model).w; }

/* dist(q1, q2): line2d.jl:17-20, snake2d.jl:25-34, arm33.jl:23-32, capsule3d.jl:25-35 */
double orc_chain_dist(const orc_ctx *c, const double *q1, const double *q2) {
    chain_desc m = desc_of(c->model);
    double v[6];
    for (int k = 0; k < m.npos; ++k) v[k] = q1[k] - q2[k];
    for (int k = m.npos; k < m.sd; ++k) v[k] = orc_diff_angles(q1[k], q2[k]) / PI64;
    return orc_norm(v, m.sd);
}

/* get_mid_status: line2d.jl:9-15, snake2d.jl:14-23, arm33.jl:12-21, capsule3d.jl:14-23 */
void orc_chain_mid_status(const orc_ctx *c, const double *p, const double *q, double *out) {
    chain_desc m = desc_of(c->model);
    for (int k = 0; k < m.npos; ++k) out[k] = (p[k] + q[k]) / 2;
    for (int k = m.npos; k < m.sd; ++k) out[k] = orc_diff_angles(q[k], p[k]) / 2 + p[k];
}

/* joint points of the posture with (already interpolated) state s */
static void chain_points(const orc_ctx *c, int i, const double *s, double P[MAXP][3]) {
    const double r = c->rads[i];
    switch (c->model) {
    case ORC_LINE2D: /* line2d.jl:32-33 / :52-57: root a, tip b = a + r [cos t, sin t] */
        P[0][0] = s[0];
        P[0][1] = s[1];
        P[1][0] = P[0][0] + r * orc_cos(s[2]);
        P[1][1] = P[0][1] + r * orc_sin(s[2]);
        break;
    case ORC_SNAKE2D: /* snake2d.jl:36-44 / :105-107 */
        P[0][0] = s[0];
        P[0][1] = s[1];
        for (int k = 1; k < 5; ++k) {
            P[k][0] = P[k - 1][0] + r * orc_cos(s[1 + k]);
            P[k][1] = P[k - 1][1] + r * orc_sin(s[1 + k]);
        }
        break;
    case ORC_ARM33: /* arm33.jl:34-40 / :83-96: spherical angles (theta_k, phi_k) */
        for (int x = 0; x < 3; ++x) P[0][x] = c->base[3 * i + x];
        for (int k = 1; k < 4; ++k) {
            const double th = s[2 * (k - 1)], ph = s[2 * (k - 1) + 1];
            P[k][0] = P[k - 1][0] + r * (orc_sin(th) * orc_cos(ph));
            P[k][1] = P[k - 1][1] + r * (orc_sin(th) * orc_sin(ph));
            P[k][2] = P[k - 1][2] + r * orc_cos(th);
        }
        break;
    default: { /* capsule3d.jl:38-59 / :89-93: tip = root + R * [axes, 0, 0] (first column of R) */
        const double ax = c->aux[i];
        const double cph = orc_cos(s[3]), sph = orc_sin(s[3]);
        const double cth = orc_cos(s[4]), sth = orc_sin(s[4]);
        for (int x = 0; x < 3; ++x) P[0][x] = s[x];
        P[1][0] = P[0][0] + (cph * cth) * ax;
        P[1][1] = P[0][1] + (sph * cth) * ax;
        P[1][2] = P[0][2] + (-sth) * ax;
        break;
    }
    }
}

/* posture tests of connect: bounds, obstacles, self collision.  point_form selects the bounds
 * rule of line2d (line2d.jl:35 uses [r, 1-r] for connect(q,i) but [0,1] in the edge loop :60) */
static int chain_posture_free(const orc_ctx *c, int i, double P[MAXP][3], int point_form) {
    chain_desc m = desc_of(c->model);
    const double r = c->rads[i];
    double lo = 0.0, hi = 1.0;
    if (c->model == ORC_CAPSULE3D || (c->model == ORC_LINE2D && point_form)) { /* capsule3d.jl:74,96 */
        lo = r;
        hi = 1 - r;
    }
    for (int k = 0; k < m.p; ++k)
        for (int x = 0; x < m.w; ++x)
            if (P[k][x] < lo || hi < P[k][x]) return 0;
    for (int k = 1; k < m.p; ++k) /* each link against every obstacle */
        for (int o = 0; o < c->n_obs; ++o) {
            const double *ob = c->obs + (m.w + 1) * o;
            const double thr = c->model == ORC_CAPSULE3D ? ob[m.w] + r : ob[m.w]; /* capsule3d.jl:76 */
            if (orc_dist_sp(P[k - 1], P[k], ob, m.w) < thr) return 0;
        }
    if (c->model == ORC_SNAKE2D) { /* check_self_collision_snake2d, snake2d.jl:46-59 */
        const double sd = c->safety_dist;
        if (orc_dist_ss(P[0], P[1], P[2], P[3], 2) < sd || orc_dist_ss(P[0], P[1], P[3], P[4], 2) < sd ||
            orc_dist_ss(P[1], P[2], P[3], P[4], 2) < sd || orc_dist_sp(P[2], P[3], P[4], 2) < sd)
            return 0;
    } else if (c->model == ORC_ARM33) { /* arm33.jl:64,109 */
        if (orc_dist_ss(P[0], P[1], P[2], P[3], 3) < c->safety_dist) return 0;
    }
    return 1;
}

/* every link of i against every link of j: line2d.jl:110, snake2d.jl:192-194, arm33.jl:206-208,
 * capsule3d.jl:163 (threshold rads[i] + rads[j] for the capsule, safety_dist otherwise) */
static int chain_links_hit(const orc_ctx *c, int i, int j, double Pi[MAXP][3], double Pj[MAXP][3]) {
    chain_desc m = desc_of(c->model);
    const double thr = c->model == ORC_CAPSULE3D ? c->rads[i] + c->rads[j] : c->safety_dist;
    for (int k = 1; k < m.p; ++k)
        for (int l = 1; l < m.p; ++l)
            if (orc_dist_ss(Pi[k - 1], Pi[k], Pj[l - 1], Pj[l], m.w) < thr) return 1;
    return 0;
}

/* one move: schedule + per-component interpolation targets */
typedef struct {
    int n;
    double *es;
    double dt[6];
} chain_move;

static int chain_move_init(const orc_ctx *c, const double *qf, const double *qt, chain_move *mv) {
    chain_desc m = desc_of(c->model);
    const double D = orc_chain_dist(c, qf, qt);
    mv->es = (double *)malloc(sizeof(double) * CH_MAX_STEPS);
    mv->n = orc_step_schedule(c->step_dist, D, mv->es, CH_MAX_STEPS);
    for (int k = m.npos; k < m.sd; ++k) mv->dt[k] = orc_diff_angles(qt[k], qf[k]);
    return mv->n <= CH_MAX_STEPS;
}

static void chain_state_at(const orc_ctx *c, const double *qf, const double *qt, const chain_move *mv,
                           double e, double *s) {
    chain_desc m = desc_of(c->model);
    for (int k = 0; k < m.npos; ++k) s[k] = (1 - e) * qf[k] + e * qt[k]; /* line2d.jl:53 */
    for (int k = m.npos; k < m.sd; ++k) s[k] = qf[k] + e * mv->dt[k];    /* line2d.jl:55 */
}

/* connect(q, i) */
int orc_chain_connect_point(const orc_ctx *c, const double *q, int i) {
    double P[MAXP][3];
    chain_points(c, i, q, P);
    return chain_posture_free(c, i, P, 1);
}

/* connect(q_from, q_to, i): line2d.jl:47-68, snake2d.jl:87-124, arm33.jl:70-113, capsule3d.jl:81-103 */
int orc_chain_connect_edge(const orc_ctx *c, const double *qf, const double *qt, int i) {
    const double D = orc_chain_dist(c, qf, qt);
    if (!isnan(c->max_dist) && D > c->max_dist) return 0;
    chain_move mv;
    int ok = chain_move_init(c, qf, qt, &mv);
    for (int s = 0; ok && s < mv.n; ++s) {
        double st[6], P[MAXP][3];
        chain_state_at(c, qf, qt, &mv, mv.es[s], st);
        chain_points(c, i, st, P);
        if (!chain_posture_free(c, i, P, 0)) ok = 0;
    }
    free(mv.es);
    return ok;
}

/* collide(q_i, q_j, i, j) */
int orc_chain_collide_static(const orc_ctx *c, const double *qi, const double *qj, int i, int j) {
    double Pi[MAXP][3], Pj[MAXP][3];
    chain_points(c, i, qi, Pi);
    chain_points(c, j, qj, Pj);
    return chain_links_hit(c, i, j, Pi, Pj);
}

/* collide(qi_from, qi_to, qj_from, qj_to, i, j) */
int orc_chain_collide_pair(const orc_ctx *c, const double *qif, const double *qit, const double *qjf,
                           const double *qjt, int i, int j) {
    chain_move mi, mj;
    int ok = chain_move_init(c, qif, qit, &mi) & chain_move_init(c, qjf, qjt, &mj);
    int hit = 0;
    for (int s = 0; ok && !hit && s < mi.n; ++s) {
        double si[6], Pi[MAXP][3];
        chain_state_at(c, qif, qit, &mi, mi.es[s], si);
        chain_points(c, i, si, Pi);
        for (int t = 0; !hit && t < mj.n; ++t) {
            double sj[6], Pj[MAXP][3];
            chain_state_at(c, qjf, qjt, &mj, mj.es[t], sj);
            chain_points(c, j, sj, Pj);
            hit = chain_links_hit(c, i, j, Pi, Pj);
        }
    }
    free(mi.es);
    free(mj.es);
    return hit;
}

/* collide(Q, q_to, i): moving agent i against the static postures of the others */
int orc_chain_collide_config_move(const orc_ctx *c, const double *Q, const double *q_to, int i) {
    chain_desc m = desc_of(c->model);
    const int N = c->n_agents;
    const double *qf = Q + m.sd * i;
    chain_move mv;
    int ok = chain_move_init(c, qf, q_to, &mv);
    int hit = 0;
    for (int s = 0; ok && !hit && s < mv.n; ++s) {
        double si[6], Pi[MAXP][3];
        chain_state_at(c, qf, q_to, &mv, mv.es[s], si);
        chain_points(c, i, si, Pi);
        for (int j = 0; !hit && j < N; ++j) {
            if (j == i) continue;
            double Pj[MAXP][3];
            chain_points(c, j, Q + m.sd * j, Pj);
            hit = chain_links_hit(c, i, j, Pi, Pj);
        }
    }
    free(mv.es);
    return hit;
}

/* gen_uniform_sampling: line2d.jl:165-169, snake2d.jl:261-265, arm33.jl:288-292,
 * capsule3d.jl:216-220: positions rand(), angles rand() * 2pi.  u[k] = k-th uniform of the try. */
void orc_chain_scale_sample(const orc_ctx *c, const double *u, double *q) {
    chain_desc m = desc_of(c->model);
    const double TWO_PI = 6.283185307179586;
    for (int k = 0; k < m.npos; ++k) q[k] = u[k];
    for (int k = m.npos; k < m.sd; ++k) q[k] = u[k] * TWO_PI;
}
