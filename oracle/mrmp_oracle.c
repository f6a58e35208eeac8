/*
 * mrmp_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See mrmp_oracle.h for the scope / pinning statement.
 *
 * Build: gcc -O2 -std=c11 -fopenmp -ffp-contract=off -fno-fast-math -fPIC -shared
 * (-DORC_DOT_FMA builds the sensitivity variant in which the 2-3 element `dot`
 * accumulates with fma like an FMA-enabled OpenBLAS ddot tail loop would.)
 */
#include "mrmp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------ */
/* Julia Base / LinearAlgebra semantics that live outside /root/reference    */
/* ------------------------------------------------------------------------ */

/* Julia max(x,y) on Float64 propagates NaN */
static inline double jl_max(double a, double b) {
    if (isnan(a) || isnan(b)) return NAN;
    return a > b ? a : b;
}
/* Julia min(x,y) on Float64 propagates NaN (used by `minimum`) */
static inline double jl_min(double a, double b) {
    if (isnan(a) || isnan(b)) return NAN;
    return a < b ? a : b;
}

/* LinearAlgebra.norm(::Vector{Float64}) for length < 32 -> generic_norm2:
 * maxabs early-outs, plain left-to-right sum of squares, scaled variant only when
 * maxabs^2 under/overflows. */
double orc_norm(const double *v, int d) {
    double maxabs = fabs(v[0]);
    for (int k = 1; k < d; ++k) maxabs = jl_max(maxabs, fabs(v[k]));
    if (maxabs == 0.0 || isinf(maxabs)) return maxabs;
    double mm = maxabs * maxabs;
    if (isfinite((double)d * maxabs * maxabs) && mm != 0.0) {
        double s = v[0] * v[0];
        for (int k = 1; k < d; ++k) s += v[k] * v[k];
        return sqrt(s);
    } else {
        double t = fabs(v[0]) / maxabs;
        double s = t * t;
        for (int k = 1; k < d; ++k) {
            t = fabs(v[k]) / maxabs;
            s += t * t;
        }
        return maxabs * sqrt(s);
    }
}

/* LinearAlgebra.dot(::Vector{Float64},::Vector{Float64}) -> BLAS ddot; for 2-3
 * elements: x1*y1 + x2*y2 (+ x3*y3), accumulated left to right. */
double orc_dot(const double *a, const double *b, int d) {
#ifdef ORC_DOT_FMA
    double s = 0.0;
    for (int k = 0; k < d; ++k) s = fma(a[k], b[k], s);
    return s;
#else
    double s = a[0] * b[0];
    for (int k = 1; k < d; ++k) s += a[k] * b[k];
    return s;
#endif
}

int orc_dot_is_fma(void) {
#ifdef ORC_DOT_FMA
    return 1;
#else
    return 0;
#endif
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------ */
/* geometry: src/utils.jl:17-99                                              */
/* ------------------------------------------------------------------------ */

/* utils.jl:17-19  dist(a,b) = norm(a - b) */
double orc_dist_pp(const double *a, const double *b, int d) {
    double v[3];
    for (int k = 0; k < d; ++k) v[k] = a[k] - b[k];
    return orc_norm(v, d);
}

/* utils.jl:44-57  segment [a,b] to point c */
double orc_dist_sp(const double *a, const double *b, const double *c, int d) {
    double u[3], bc[3], ac[3], p[3];
    for (int k = 0; k < d; ++k) {
        u[k] = a[k] - b[k];
        bc[k] = b[k] - c[k];
        ac[k] = a[k] - c[k];
    }
    double df0 = orc_dot(u, bc, d); /* :46 */
    double df1 = orc_dot(u, ac, d); /* :47 */
    double e = 1.0;                 /* :49 */
    if (df0 * df1 < 0) {            /* :50 */
        e = -orc_dot(u, bc, d) / orc_dot(u, u, d); /* :51 */
    } else if (df0 >= 0) {          /* :52 */
        e = 0.0;
    }
    for (int k = 0; k < d; ++k) p[k] = (e * a[k] + (1 - e) * b[k]) - c[k]; /* :56 */
    return orc_norm(p, d);
}

/* utils.jl:59-94  segment [p11,p12] to segment [p21,p22] */
double orc_dist_ss(const double *p11, const double *p12, const double *p21, const double *p22,
                   int d) {
    double v1[3], v2[3], w[3];
    for (int k = 0; k < d; ++k) {
        v1[k] = p12[k] - p11[k]; /* :68 */
        v2[k] = p22[k] - p21[k]; /* :69 */
        w[k] = p21[k] - p11[k];
    }
    double D1 = orc_dot(w, v1, d);  /* :70 */
    double D2 = orc_dot(w, v2, d);  /* :71 */
    double Dv = orc_dot(v1, v2, d); /* :72 */
    double V1 = orc_dot(v1, v1, d); /* :73 */
    double V2 = orc_dot(v2, v2, d); /* :74 */
    double D = V1 * V2 - Dv * Dv;   /* :75 */
    if (D > 0) {
        double t1 = (D1 * V2 - D2 * Dv) / D; /* :77 */
        double t2 = (D1 * Dv - D2 * V1) / D; /* :78 */
        if (0 <= t1 && t1 <= 1 && 0 <= t2 && t2 <= 1) {
            double q[3];
            for (int k = 0; k < d; ++k)
                q[k] = (p11[k] + t1 * v1[k]) - (p21[k] + t2 * v2[k]); /* :80-82 */
            return orc_norm(q, d);
        }
    }
    /* :86-93 minimum([...]) */
    double m = orc_dist_sp(p11, p12, p21, d);
    m = jl_min(m, orc_dist_sp(p11, p12, p22, d));
    m = jl_min(m, orc_dist_sp(p21, p22, p11, d));
    m = jl_min(m, orc_dist_sp(p21, p22, p12, d));
    return m;
}

/* utils.jl:97-99 */
double orc_diff_angles(double t1, double t2) {
    double dt = t1 - t2;
    return orc_atan2(orc_sin(dt), orc_cos(dt));
}

/* ------------------------------------------------------------------------ */
/* context                                                                   */
/* ------------------------------------------------------------------------ */

orc_ctx *orc_ctx_create2(int model, int n_agents, const double *rads, const double *base,
                         const double *aux, int n_obs, const double *obs_aos, double step_dist,
                         double safety_dist, double max_dist) {
    if (model < 0 || model > ORC_CAPSULE3D || n_agents <= 0 || n_obs < 0) return NULL;
    orc_ctx *c = (orc_ctx *)calloc(1, sizeof(orc_ctx));
    c->model = model;
    if (model >= ORC_LINE2D) {
        c->d = orc_chain_state_dim(model);
        c->wd = orc_chain_workspace_dim(model);
    } else {
        c->d = model == ORC_POINT3D ? 3 : 2;
        c->wd = model == ORC_POINT3D ? 3 : 2;
    }
    c->n_agents = n_agents;
    c->rads = (double *)malloc(sizeof(double) * n_agents);
    memcpy(c->rads, rads, sizeof(double) * n_agents);
    if (model == ORC_ARM22 || model == ORC_ARM33) {
        if (!base) {
            free(c->rads);
            free(c);
            return NULL;
        }
        c->base = (double *)malloc(sizeof(double) * c->wd * n_agents);
        memcpy(c->base, base, sizeof(double) * c->wd * n_agents);
    }
    if (model == ORC_CAPSULE3D) {
        if (!aux) {
            free(c->rads);
            free(c);
            return NULL;
        }
        c->aux = (double *)malloc(sizeof(double) * n_agents);
        memcpy(c->aux, aux, sizeof(double) * n_agents);
    }
    c->n_obs = n_obs;
    c->obs = (double *)malloc(sizeof(double) * (c->wd + 1) * (n_obs > 0 ? n_obs : 1));
    if (n_obs > 0) memcpy(c->obs, obs_aos, sizeof(double) * (c->wd + 1) * n_obs);
    c->step_dist = step_dist;
    c->safety_dist = safety_dist;
    c->max_dist = max_dist;
    return c;
}

orc_ctx *orc_ctx_create(int model, int n_agents, const double *rads, const double *base,
                        int n_obs, const double *obs_aos, double step_dist,
                        double safety_dist, double max_dist) {
    return orc_ctx_create2(model, n_agents, rads, base, NULL, n_obs, obs_aos, step_dist,
                           safety_dist, max_dist);
}

void orc_ctx_destroy(orc_ctx *c) {
    if (!c) return;
    free(c->rads);
    free(c->base);
    free(c->aux);
    free(c->obs);
    free(c);
}

/* ------------------------------------------------------------------------ */
/* arm22 helpers: src/models/arm22.jl                                        */
/* ------------------------------------------------------------------------ */

/* arm22.jl:15-20 */
static double arm22_dist(const double *q1, const double *q2) {
    const double PI = 3.141592653589793; /* Float64(pi) */
    double v[2] = {orc_diff_angles(q1[0], q2[0]) / PI, orc_diff_angles(q1[1], q2[1]) / PI};
    return orc_norm(v, 2);
}

/* arm22.jl:22-27 / :69-70 / :122-123: joints from angles */
static void arm22_joints(double t1, double t2, const double *a, double rad, double *b, double *c) {
    b[0] = a[0] + rad * orc_cos(t1);
    b[1] = a[1] + rad * orc_sin(t1);
    c[0] = b[0] + rad * orc_cos(t2);
    c[1] = b[1] + rad * orc_sin(t2);
}

/* Base.rat (twiceprecision.jl), Float64 -> narrow type Float32, m = 2^24 */
static void jl_rat(double x, int64_t *num, int64_t *den) {
    const double m = 16777216.0;
    double y = x;
    int64_t a = 1, d = 1, b = 0, c = 0;
    while (fabs(y) <= m) {
        int64_t f = (int64_t)trunc(y);
        y -= (double)f;
        int64_t na = f * a + c;
        c = a;
        a = na;
        int64_t nb = f * b + d;
        d = b;
        b = nb;
        int64_t aa = a < 0 ? -a : a, ab = b < 0 ? -b : b;
        if (!((aa > ab ? aa : ab) <= 16777216)) {
            *num = c;
            *den = d;
            return;
        }
        if ((double)a / (double)b == x) {
            *num = a;
            *den = b;
            return;
        }
        y = 1.0 / y;
    }
    *num = a;
    *den = b;
}

static int64_t gcd64(int64_t a, int64_t b) {
    if (a < 0) a = -a;
    if (b < 0) b = -b;
    while (b) {
        int64_t t = a % b;
        a = b;
        b = t;
    }
    return a;
}

static int isbetween(double a, double x, double b) { return (a <= x && x <= b) || (b <= x && x <= a); }

/* vcat(collect(0:step:D) / D, 1.0)  (arm22.jl:65,117,125,180)
 * Julia (:)(start::Float64, step, stop): rational lift when start/step/stop have
 * short exact rationals, else literal fallback (Base/twiceprecision.jl). */
int orc_step_schedule(double step, double D, double *out, int cap) {
    int64_t len = -1;
    int rational = 0;
    int64_t step_n = 0, den = 0;
    int64_t sn, sd;
    jl_rat(step, &sn, &sd);
    if (sd != 0 && (double)sn / (double)sd == step) {
        int64_t en, ed;
        jl_rat(D, &en, &ed);
        if (ed != 0 && (double)en / (double)ed == D) {
            /* start = 0.0 -> (0, 1); den = lcm(1, step_d) */
            den = sd / gcd64(1, sd);
            const double m = 9007199254740992.0;
            if (den != 0 && fabs(0.0 * (double)den) <= m && fabs(step * (double)den) <= m &&
                den % sd == 0) {
                step_n = (int64_t)rint(step * (double)den);
                if (step_n != 0 && ed != 0) {
                    int64_t q = (den * en - ed * 0) / (step_n * ed); /* div: truncation */
                    int64_t l = q + 1;
                    if (l < 0) l = 0;
                    if (isbetween(0.0, 0.0 + (double)(l - 1) * step, D + step / 2) &&
                        !isbetween(0.0, 0.0 + (double)l * step, D)) {
                        rational = 1;
                        len = l;
                    }
                }
            }
        }
    }
    if (!rational) {
        double lf = (D - 0.0) / step;
        if (lf < 0) {
            len = 0;
        } else if (lf == 0) {
            len = 1;
        } else {
            len = (int64_t)rint(lf) + 1;
            double stop2 = 0.0 + (double)(len - 1) * step;
            len -= (0.0 < D && D < stop2) ? 1 : 0;
        }
    }
    int n = 0;
    for (int64_t k = 0; k < len; ++k) {
        double r = rational ? (double)(k * step_n) / (double)den : (double)k * step;
        if (n < cap) out[n] = r / D;
        ++n;
    }
    if (n < cap) out[n] = 1.0;
    ++n;
    return n;
}

#define ARM_MAX_STEPS 4096

/* arm22.jl:43-51 / :73-82: one posture test; 1 = free */
static int arm22_posture_free(const orc_ctx *c, const double *a, const double *b,
                              const double *cc) {
    const double pts[6] = {a[0], a[1], b[0], b[1], cc[0], cc[1]};
    for (int k = 0; k < 6; ++k)
        if (pts[k] < 0 || 1 < pts[k]) return 0;
    for (int m = 0; m < c->n_obs; ++m) {
        const double *o = c->obs + 3 * m;
        if (orc_dist_sp(a, b, o, 2) < o[2] || orc_dist_sp(b, cc, o, 2) < o[2]) return 0;
    }
    if (orc_dist_sp(a, b, cc, 2) < c->safety_dist) return 0;
    return 1;
}

/* arm22.jl:134-142 / :153-161: four link-link tests */
static int arm22_links_hit(const orc_ctx *c, const double *ai, const double *bi, const double *ci,
                           const double *aj, const double *bj, const double *cj) {
    double s = c->safety_dist;
    return orc_dist_ss(ai, bi, aj, bj, 2) < s || orc_dist_ss(ai, bi, bj, cj, 2) < s ||
           orc_dist_ss(bi, ci, aj, bj, 2) < s || orc_dist_ss(bi, ci, bj, cj, 2) < s;
}

/* ------------------------------------------------------------------------ */
/* state-space helpers                                                       */
/* ------------------------------------------------------------------------ */

double orc_state_dist(const orc_ctx *c, const double *q1, const double *q2) {
    if (c->model >= ORC_LINE2D) return orc_chain_dist(c, q1, q2);
    if (c->model == ORC_ARM22) return arm22_dist(q1, q2);
    return orc_dist_pp(q1, q2, c->d); /* point2d.jl:12-14, point3d.jl:13-15 */
}

void orc_mid_status(const orc_ctx *c, const double *p, const double *q, double *out) {
    if (c->model >= ORC_LINE2D) { orc_chain_mid_status(c, p, q, out); return; };
    if (c->model == ORC_ARM22) { /* arm22.jl:8-13 */
        out[0] = orc_diff_angles(q[0], p[0]) / 2 + p[0];
        out[1] = orc_diff_angles(q[1], p[1]) / 2 + p[1];
    } else { /* point2d.jl:8-10, point3d.jl:9-11 */
        for (int k = 0; k < c->d; ++k) out[k] = (p[k] + q[k]) / 2;
    }
}

/* ------------------------------------------------------------------------ */
/* closure bodies                                                            */
/* ------------------------------------------------------------------------ */

/* connect(q, i): point2d.jl:49-57, point3d.jl:54-64, arm22.jl:40-54 */
int orc_connect_point(const orc_ctx *c, const double *q, int i) {
    if (c->model >= ORC_LINE2D) return orc_chain_connect_point(c, q, i);
    if (c->model == ORC_ARM22) {
        double b[2], cc[2];
        const double *a = c->base + 2 * i;
        arm22_joints(q[0], q[1], a, c->rads[i], b, cc);
        return arm22_posture_free(c, a, b, cc);
    }
    int d = c->d;
    double r = c->rads[i];
    for (int k = 0; k < d; ++k)
        if (q[k] < r || 1 - r < q[k]) return 0;
    for (int m = 0; m < c->n_obs; ++m) {
        const double *o = c->obs + (d + 1) * m;
        if (orc_dist_pp(q, o, d) < o[d] + r) return 0;
    }
    return 1;
}

/* connect(q_from, q_to, i): point2d.jl:59-66, point3d.jl:66-77, arm22.jl:56-86 */
int orc_connect_edge(const orc_ctx *c, const double *qf, const double *qt, int i) {
    if (c->model >= ORC_LINE2D) return orc_chain_connect_edge(c, qf, qt, i);
    if (c->model == ORC_ARM22) {
        double D = arm22_dist(qf, qt);                      /* :57 */
        if (!isnan(c->max_dist) && D > c->max_dist) return 0; /* :58 */
        double dt1 = orc_diff_angles(qt[0], qf[0]);         /* :60 */
        double dt2 = orc_diff_angles(qt[1], qf[1]);         /* :61 */
        const double *a = c->base + 2 * i;
        double es[ARM_MAX_STEPS];
        int n = orc_step_schedule(c->step_dist, D, es, ARM_MAX_STEPS);
        if (n > ARM_MAX_STEPS) return 0;
        for (int s = 0; s < n; ++s) {
            double e = es[s];
            double t1 = qf[0] + e * dt1, t2 = qf[1] + e * dt2; /* :66-67 */
            double b[2], cc[2];
            arm22_joints(t1, t2, a, c->rads[i], b, cc);
            if (!arm22_posture_free(c, a, b, cc)) return 0;
        }
        return 1;
    }
    int d = c->d;
    double r = c->rads[i];
    for (int k = 0; k < d; ++k)
        if (qt[k] < r || 1 - r < qt[k]) return 0; /* q_to only */
    for (int m = 0; m < c->n_obs; ++m) {
        const double *o = c->obs + (d + 1) * m;
        if (orc_dist_sp(qf, qt, o, d) < o[d] + r) return 0;
    }
    return 1;
}

/* collide(qi_from,qi_to,qj_from,qj_to,i,j): point.jl:9-12, arm22.jl:100-146 */
int orc_collide_pair(const orc_ctx *c, const double *qif, const double *qit, const double *qjf,
                     const double *qjt, int i, int j) {
    if (c->model >= ORC_LINE2D) return orc_chain_collide_pair(c, qif, qit, qjf, qjt, i, j);
    if (c->model == ORC_ARM22) {
        double Di = arm22_dist(qif, qit), Dj = arm22_dist(qjf, qjt);
        double dt1i = orc_diff_angles(qit[0], qif[0]), dt2i = orc_diff_angles(qit[1], qif[1]);
        double dt1j = orc_diff_angles(qjt[0], qjf[0]), dt2j = orc_diff_angles(qjt[1], qjf[1]);
        double ei[ARM_MAX_STEPS], ej[ARM_MAX_STEPS];
        int ni = orc_step_schedule(c->step_dist, Di, ei, ARM_MAX_STEPS);
        int nj = orc_step_schedule(c->step_dist, Dj, ej, ARM_MAX_STEPS);
        if (ni > ARM_MAX_STEPS || nj > ARM_MAX_STEPS) return 0;
        const double *ai = c->base + 2 * i, *aj = c->base + 2 * j;
        for (int s = 0; s < ni; ++s) {
            double bi[2], ci[2];
            arm22_joints(qif[0] + ei[s] * dt1i, qif[1] + ei[s] * dt2i, ai, c->rads[i], bi, ci);
            for (int t = 0; t < nj; ++t) {
                double bj[2], cj[2];
                arm22_joints(qjf[0] + ej[t] * dt1j, qjf[1] + ej[t] * dt2j, aj, c->rads[j], bj, cj);
                if (arm22_links_hit(c, ai, bi, ci, aj, bj, cj)) return 1;
            }
        }
        return 0;
    }
    return orc_dist_ss(qif, qit, qjf, qjt, c->d) < c->rads[i] + c->rads[j];
}

/* collide(q_i,q_j,i,j): point.jl:14-16, arm22.jl:148-162 */
int orc_collide_static(const orc_ctx *c, const double *qi, const double *qj, int i, int j) {
    if (c->model >= ORC_LINE2D) return orc_chain_collide_static(c, qi, qj, i, j);
    if (c->model == ORC_ARM22) {
        const double *ai = c->base + 2 * i, *aj = c->base + 2 * j;
        double bi[2], ci[2], bj[2], cj[2];
        arm22_joints(qi[0], qi[1], ai, c->rads[i], bi, ci);
        arm22_joints(qj[0], qj[1], aj, c->rads[j], bj, cj);
        return arm22_links_hit(c, ai, bi, ci, aj, bj, cj);
    }
    return orc_dist_pp(qi, qj, c->d) < c->rads[i] + c->rads[j];
}

/* collide(Q_from,Q_to): point.jl:18-25, arm22.jl:164-171 */
int orc_collide_configs(const orc_ctx *c, const double *Qf, const double *Qt,
                        int64_t *first_pair) {
    int N = c->n_agents, d = c->d;
    for (int i = 0; i < N; ++i)
        for (int j = i + 1; j < N; ++j)
            if (orc_collide_pair(c, Qf + d * i, Qt + d * i, Qf + d * j, Qt + d * j, i, j)) {
                if (first_pair) *first_pair = (int64_t)i * N + j;
                return 1;
            }
    if (first_pair) *first_pair = -1;
    return 0;
}

/* collide(Q, q_to, i): point.jl:27-33, arm22.jl:173-209 */
int orc_collide_config_move(const orc_ctx *c, const double *Q, const double *q_to, int i) {
    if (c->model >= ORC_LINE2D) return orc_chain_collide_config_move(c, Q, q_to, i);
    int N = c->n_agents, d = c->d;
    const double *qf = Q + d * i;
    if (c->model == ORC_ARM22) {
        double Di = arm22_dist(qf, q_to);
        double dt1 = orc_diff_angles(q_to[0], qf[0]), dt2 = orc_diff_angles(q_to[1], qf[1]);
        double es[ARM_MAX_STEPS];
        int n = orc_step_schedule(c->step_dist, Di, es, ARM_MAX_STEPS);
        if (n > ARM_MAX_STEPS) return 0;
        const double *ai = c->base + 2 * i;
        for (int s = 0; s < n; ++s) {
            double bi[2], ci[2];
            arm22_joints(qf[0] + es[s] * dt1, qf[1] + es[s] * dt2, ai, c->rads[i], bi, ci);
            for (int j = 0; j < N; ++j) {
                if (j == i) continue;
                const double *aj = c->base + 2 * j;
                double bj[2], cj[2];
                arm22_joints(Q[2 * j], Q[2 * j + 1], aj, c->rads[j], bj, cj);
                if (arm22_links_hit(c, ai, bi, ci, aj, bj, cj)) return 1;
            }
        }
        return 0;
    }
    for (int j = 0; j < N; ++j) {
        if (j == i) continue;
        if (orc_dist_sp(qf, q_to, Q + d * j, d) < c->rads[i] + c->rads[j]) return 1;
    }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* batched drivers                                                           */
/* ------------------------------------------------------------------------ */

static int pick_threads(int nthreads) {
#ifdef _OPENMP
    return nthreads > 0 ? nthreads : omp_get_max_threads();
#else
    (void)nthreads;
    return 1;
#endif
}

void orc_connect_points_batch(const orc_ctx *c, int64_t n, const int32_t *agent, const double *q,
                              uint8_t *out, int nthreads) {
    int d = c->d, nt = pick_threads(nthreads);
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int64_t k = 0; k < n; ++k) out[k] = (uint8_t)orc_connect_point(c, q + d * k, agent[k]);
}

void orc_connect_edges_batch(const orc_ctx *c, int64_t n, const int32_t *agent, const double *qf,
                             const double *qt, uint8_t *out, int nthreads) {
    int d = c->d, nt = pick_threads(nthreads);
#pragma omp parallel for num_threads(nt) schedule(dynamic, 4096)
    for (int64_t k = 0; k < n; ++k)
        out[k] = (uint8_t)orc_connect_edge(c, qf + d * k, qt + d * k, agent[k]);
}

void orc_collide_pairs_batch(const orc_ctx *c, int64_t n, const int32_t *ai, const int32_t *aj,
                             const double *qif, const double *qit, const double *qjf,
                             const double *qjt, uint8_t *out, int nthreads) {
    int d = c->d, nt = pick_threads(nthreads);
    int chunk = c->model == ORC_ARM22 ? 16 : 4096;
#pragma omp parallel for num_threads(nt) schedule(dynamic, chunk)
    for (int64_t k = 0; k < n; ++k)
        out[k] = (uint8_t)orc_collide_pair(c, qif + d * k, qit + d * k, qjf + d * k, qjt + d * k,
                                           ai[k], aj[k]);
}

void orc_collide_static_pairs_batch(const orc_ctx *c, int64_t n, const int32_t *ai,
                                    const int32_t *aj, const double *qi, const double *qj,
                                    uint8_t *out, int nthreads) {
    int d = c->d, nt = pick_threads(nthreads);
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int64_t k = 0; k < n; ++k)
        out[k] = (uint8_t)orc_collide_static(c, qi + d * k, qj + d * k, ai[k], aj[k]);
}

void orc_collide_configs_batch(const orc_ctx *c, int64_t n_cfg, const double *Qf,
                               const double *Qt, uint8_t *out, int64_t *first_pair,
                               int nthreads) {
    int64_t stride = (int64_t)c->d * c->n_agents;
    int nt = pick_threads(nthreads);
#pragma omp parallel for num_threads(nt) schedule(dynamic, 16)
    for (int64_t k = 0; k < n_cfg; ++k) {
        int64_t fp;
        out[k] = (uint8_t)orc_collide_configs(c, Qf + stride * k, Qt + stride * k, &fp);
        if (first_pair) first_pair[k] = fp;
    }
}

void orc_collide_config_moves_batch(const orc_ctx *c, int agent_i, const double *Q,
                                    int64_t n_cand, const double *q_to, uint8_t *out,
                                    int nthreads) {
    int d = c->d, nt = pick_threads(nthreads);
#pragma omp parallel for num_threads(nt) schedule(dynamic, 64)
    for (int64_t k = 0; k < n_cand; ++k)
        out[k] = (uint8_t)orc_collide_config_move(c, Q, q_to + d * k, agent_i);
}

void orc_collide_edges_indexed_batch(const orc_ctx *c, int64_t n, const int32_t *ai,
                                     const int32_t *aj, const int32_t *vif, const int32_t *vit,
                                     const int32_t *vjf, const int32_t *vjt,
                                     const double *nodes, int64_t nv_stride, uint8_t *out,
                                     int nthreads) {
    int d = c->d, nt = pick_threads(nthreads);
    int chunk = c->model == ORC_ARM22 ? 16 : 4096;
#pragma omp parallel for num_threads(nt) schedule(dynamic, chunk)
    for (int64_t k = 0; k < n; ++k) {
        const double *ni = nodes + (int64_t)ai[k] * nv_stride * d;
        const double *nj = nodes + (int64_t)aj[k] * nv_stride * d;
        out[k] = (uint8_t)orc_collide_pair(c, ni + (int64_t)d * vif[k], ni + (int64_t)d * vit[k],
                                           nj + (int64_t)d * vjf[k], nj + (int64_t)d * vjt[k],
                                           ai[k], aj[k]);
    }
}

/* Cross product of the 6-arity collide (point.jl:9-12 / arm22.jl:100-146): every row edge
 * against every column edge, same-agent pairs excluded; the loop shape of PP's invalid()
 * (pp.jl:179-189) and CBS's conflict count (cbs.jl:336-347).  group_size == 0: full bit
 * matrix [n_rows][ceil(n_cols/32)]; > 0: OR over consecutive column groups. */
void orc_collide_cross(const orc_ctx *c, int64_t n_rows, const int32_t *ra, const double *rf,
                       const double *rt, int64_t n_cols, const int32_t *ca, const double *cf,
                       const double *ct, int group_size, uint32_t *out, int nthreads) {
    int d = c->d, nt = pick_threads(nthreads);
    int64_t bits = group_size > 0 ? (n_cols + group_size - 1) / group_size : n_cols;
    int64_t wpr = (bits + 31) / 32;
#pragma omp parallel for num_threads(nt) schedule(dynamic, 8)
    for (int64_t r = 0; r < n_rows; ++r) {
        uint32_t *o = out + r * wpr;
        for (int64_t w = 0; w < wpr; ++w) o[w] = 0;
        for (int64_t k = 0; k < n_cols; ++k) {
            if (ra[r] == ca[k]) continue;
            if (orc_collide_pair(c, rf + d * r, rt + d * r, cf + d * k, ct + d * k, ra[r], ca[k])) {
                int64_t b = group_size > 0 ? k / group_size : k;
                o[b >> 5] |= 1u << (b & 31);
            }
        }
    }
}

/* Reductions of the same cross product that the reference's callers compute one collide call at
 * a time (mode 2 = count, 3 = first; out int32 [n_rows][n_groups]):
 *  - CBS soft conflict count, cbs.jl:336-347 (cbs != 0):  per candidate edge of agent i (row) and
 *    time step (group), count(j -> collide(q_j_to, q_i_to, j, i) || collide(q_j_from, q_j_to,
 *    q_i_from, q_i_to, j, i)) over the other agents' path edges (columns of that group) -- the
 *    COLUMN agent is the first argument of both calls, as written at :345-346.
 *  - TPG type-2 dependency scan, smoother.jl:78-106 (keys): per action_self (row) and other
 *    agent j (group): the FIRST action_other of j, in list order, with a.t > action_self.t
 *    (col_key > row_key) that collides (:86-103, `break` at the first hit).
 * group of column k = col_group[k] if given, else k / group_size. */
void orc_collide_cross_reduce(const orc_ctx *c, int64_t n_rows, const int32_t *ra, const double *rf,
                              const double *rt, const int32_t *row_key, int64_t n_cols,
                              const int32_t *ca, const double *cf, const double *ct,
                              const int32_t *col_group, const int32_t *col_key, int group_size,
                              int n_groups, int mode, int cbs, int32_t *out, int nthreads) {
    int d = c->d, nt = pick_threads(nthreads);
#pragma omp parallel for num_threads(nt) schedule(dynamic, 8)
    for (int64_t r = 0; r < n_rows; ++r) {
        int32_t *o = out + r * (int64_t)n_groups;
        for (int g = 0; g < n_groups; ++g) o[g] = mode == 3 ? -1 : 0;
        for (int64_t k = 0; k < n_cols; ++k) {
            if (ra[r] == ca[k]) continue;                       /* j != i */
            if (row_key && !(col_key[k] > row_key[r])) continue; /* smoother.jl:88 a.t > t */
            int64_t g = col_group ? col_group[k] : k / group_size;
            if (mode == 3 && o[g] >= 0) continue;               /* `break` after the first hit */
            int hit;
            if (cbs)
                hit = orc_collide_static(c, ct + d * k, rt + d * r, ca[k], ra[r]) ||
                      orc_collide_pair(c, cf + d * k, ct + d * k, rf + d * r, rt + d * r, ca[k], ra[r]);
            else
                hit = orc_collide_pair(c, rf + d * r, rt + d * r, cf + d * k, ct + d * k, ra[r], ca[k]);
            if (!hit) continue;
            if (mode == 3)
                o[g] = (int32_t)k;
            else
                o[g] += 1;
        }
    }
}

/* ------------------------------------------------------------------------ */
/* sampler: counter-based Philox4x32-10 (Salmon et al., SC'11)               */
/* ------------------------------------------------------------------------ */

void orc_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                    uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

static inline double u01_53(uint32_t hi, uint32_t lo) {
    /* 53-bit uniform in [0,1): 27 bits of hi, 26 bits of lo */
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

/* gen_uniform_sampling: StatePoint2D(rand(2)...) point2d.jl:71-75;
 * StatePoint3D(rand(3)...) point3d.jl:82-86; StateArm22((rand(2)*2pi)...) arm22.jl:214-218.
 * Try t of (seed, agent) -> counter (t_lo, t_hi, agent, block), key (seed_lo, seed_hi). */
void orc_sample_state(const orc_ctx *c, uint64_t seed, int agent, uint64_t t, double *q) {
    uint32_t r[4];
    if (c->model >= ORC_LINE2D) { /* uniform k of a try = pair (k % 2) of Philox block k / 2 */
        double u[6];
        for (int b = 0; 2 * b < c->d; ++b) {
            orc_philox4x32((uint32_t)t, (uint32_t)(t >> 32), (uint32_t)agent, (uint32_t)b,
                           (uint32_t)seed, (uint32_t)(seed >> 32), r);
            u[2 * b] = u01_53(r[0], r[1]);
            if (2 * b + 1 < c->d) u[2 * b + 1] = u01_53(r[2], r[3]);
        }
        orc_chain_scale_sample(c, u, q);
        return;
    }
    orc_philox4x32((uint32_t)t, (uint32_t)(t >> 32), (uint32_t)agent, 0u, (uint32_t)seed,
                   (uint32_t)(seed >> 32), r);
    q[0] = u01_53(r[0], r[1]);
    q[1] = u01_53(r[2], r[3]);
    if (c->model == ORC_POINT3D) {
        orc_philox4x32((uint32_t)t, (uint32_t)(t >> 32), (uint32_t)agent, 1u, (uint32_t)seed,
                       (uint32_t)(seed >> 32), r);
        q[2] = u01_53(r[0], r[1]);
    } else if (c->model == ORC_ARM22) {
        const double TWO_PI = 6.283185307179586; /* 2pi in Float64 */
        q[0] = q[0] * TWO_PI;
        q[1] = q[1] * TWO_PI;
    }
}

/* prm.jl:193-199 */
int64_t orc_sample_free(const orc_ctx *c, int agent, uint64_t seed, int64_t n_wanted,
                        double *q_out, int64_t max_tries) {
    int d = c->d;
    int64_t got = 0, t = 0;
    while (got < n_wanted && t < max_tries) {
        double q[6];
        orc_sample_state(c, seed, agent, (uint64_t)t, q);
        ++t;
        if (orc_connect_point(c, q, agent)) {
            memcpy(q_out + d * got, q, sizeof(double) * d);
            ++got;
        }
    }
    return got == n_wanted ? t : -1;
}

/* ------------------------------------------------------------------------ */
/* PRM! neighbour pass: src/solvers/prm.jl:202-210                           */
/* ------------------------------------------------------------------------ */

int64_t orc_prm_build_agent(const orc_ctx *c, int agent, int nv, double rad, const double *nodes,
                            int32_t *row_ptr, int32_t *col_idx, int64_t cap) {
    int d = c->d;
    int64_t nnz = 0;
    for (int j = 0; j < nv; ++j) {
        row_ptr[j] = (int32_t)nnz;
        for (int k = 0; k < nv; ++k) {
            if (j == k) continue; /* :204 */
            const double *qf = nodes + d * j, *qt = nodes + d * k;
            if ((isnan(rad) || orc_state_dist(c, qf, qt) <= rad) && /* :207 */
                orc_connect_edge(c, qf, qt, agent)) {
                if (nnz < cap) col_idx[nnz] = k;
                ++nnz;
            }
        }
    }
    row_ptr[nv] = (int32_t)nnz;
    return nnz;
}

/* prm.jl:251-276 (map over agents) */
void orc_prm_build(const orc_ctx *c, int agent0, int n, int nv, const double *rad,
                   const double *nodes, int32_t *row_ptr, int32_t *col_idx, int64_t cap,
                   int64_t *nnz_out, int nthreads) {
    int d = c->d, nt = pick_threads(nthreads);
#pragma omp parallel for num_threads(nt) schedule(dynamic, 1)
    for (int a = 0; a < n; ++a)
        nnz_out[a] = orc_prm_build_agent(c, agent0 + a, nv, rad[a], nodes + (int64_t)a * nv * d,
                                         row_ptr + (int64_t)a * (nv + 1),
                                         col_idx + (int64_t)a * cap, cap);
}

/* ------------------------------------------------------------------------ */
/* SSSP helpers: src/solvers/sssp.jl:87-89, 269-291                          */
/* ------------------------------------------------------------------------ */

int orc_conn(const orc_ctx *c, int agent, const double *qf, const double *qt, double epsilon) {
    return (isnan(epsilon) || orc_state_dist(c, qf, qt) <= epsilon) &&
           orc_connect_edge(c, qf, qt, agent);
}

void orc_steering(const orc_ctx *c, int agent, const double *q_rand, const double *q_near,
                  int depth, double epsilon, double *q_out) {
    int d = c->d;
    double qh[6], ql[6], q[6];
    memcpy(qh, q_rand, sizeof(double) * d);
    if (!orc_conn(c, agent, q_near, qh, epsilon)) { /* :277 */
        memcpy(ql, q_near, sizeof(double) * d);
        for (int s = 0; s < depth; ++s) {
            orc_mid_status(c, ql, qh, q); /* :281 */
            if (orc_conn(c, agent, q_near, q, epsilon))
                memcpy(ql, q, sizeof(double) * d);
            else
                memcpy(qh, q, sizeof(double) * d);
        }
        memcpy(qh, ql, sizeof(double) * d); /* :288 */
    }
    memcpy(q_out, qh, sizeof(double) * d);
}

/* ------------------------------------------------------------------------ */
/* Dijkstra distance table: src/solvers/utils.jl:146-179                     */
/* ------------------------------------------------------------------------ */

void orc_distance_table(const orc_ctx *c, int nv, const double *nodes, const int32_t *row_ptr,
                        const int32_t *col_idx, int goal, double *table) {
    /* O(nv^2) array Dijkstra: final values are independent of tie-breaking */
    int d = c->d;
    uint8_t *done = (uint8_t *)calloc((size_t)nv, 1);
    for (int v = 0; v < nv; ++v) table[v] = INFINITY; /* typemax(Float64) == Inf, :152 */
    table[goal] = 0;
    for (;;) {
        int v = -1;
        double best = INFINITY;
        for (int u = 0; u < nv; ++u)
            if (!done[u] && table[u] < best) {
                best = table[u];
                v = u;
            }
        if (v < 0) break;
        done[v] = 1;
        for (int e = row_ptr[v]; e < row_ptr[v + 1]; ++e) {
            int u = col_idx[e];
            double g = orc_state_dist(c, nodes + d * v, nodes + d * u) + table[v]; /* :168 */
            if (g < table[u]) table[u] = g;
        }
    }
    free(done);
}

/* ------------------------------------------------------------------------ */
/* one SSSP node expansion: expand! (sssp.jl:235-267) followed by the        */
/* successor filter collide(S.Q, p.q, i) (sssp.jl:196-223), as sequential C  */
/* ------------------------------------------------------------------------ */
/* nodes[cap][d] holds nv vertices of agent `agent` and receives the accepted ones.
 * q_rand[K][d] are the samples in draw order.  out_bits/in_bits [K][words]: row r describes
 * accepted vertex r (id nv+r): out bit t <=> conn(q_new, q_t) for t < id, in bit t <=>
 * conn(q_t, q_new) for t <= id.  Q[N][d]: states of all agents in the search node (Q[agent]
 * = nodes[v_from]); succ_ids = old_nbrs, new neighbours of v_from (ascending), v_from;
 * succ_collide[k] = collide(Q, q_succ, agent) or collide(Q, Q') when slow != 0.
 * Returns the number of accepted vertices; *n_succ receives the successor count. */
int orc_sssp_expand(const orc_ctx *c, int agent, double *nodes, int nv, int v_from, int K,
                    const double *q_rand, int depth, double epsilon, double thr, int words,
                    uint32_t *out_bits, uint32_t *in_bits, const double *Q, int n_old,
                    const int32_t *old_nbrs, int slow, int32_t *succ_ids, uint8_t *succ_collide,
                    int *n_succ) {
    const int d = c->d, N = c->n_agents;
    const double *q_near = nodes + (size_t)d * v_from;
    int n = nv;
    memset(out_bits, 0, sizeof(uint32_t) * (size_t)K * words);
    memset(in_bits, 0, sizeof(uint32_t) * (size_t)K * words);
    for (int k = 0; k < K; ++k) {
        double q_new[6];
        orc_steering(c, agent, q_rand + (size_t)d * k, q_near, depth, epsilon, q_new); /* :246 */
        double mn = INFINITY;                                                           /* :250 */
        int has_nan = 0;
        for (int v = 0; v < n; ++v) {
            double dd = orc_state_dist(c, nodes + (size_t)d * v, q_new);
            if (isnan(dd)) has_nan = 1;
            if (dd < mn) mn = dd;
        }
        if (has_nan || !(mn > thr)) continue;
        const int r = n - nv, id = n;
        memcpy(nodes + (size_t)d * id, q_new, sizeof(double) * d);
        for (int t = 0; t < id; ++t)                                                    /* :252-257 */
            if (orc_conn(c, agent, q_new, nodes + (size_t)d * t, epsilon))
                out_bits[(size_t)r * words + (t >> 5)] |= 1u << (t & 31);
        ++n;                                                                            /* :258 */
        for (int t = 0; t < n; ++t)                                                     /* :259-262 */
            if (orc_conn(c, agent, nodes + (size_t)d * t, q_new, epsilon))
                in_bits[(size_t)r * words + (t >> 5)] |= 1u << (t & 31);
    }
    int ns = 0;
    for (int k = 0; k < n_old; ++k) succ_ids[ns++] = old_nbrs[k];
    for (int r = 0; r < n - nv; ++r)
        if ((in_bits[(size_t)r * words + (v_from >> 5)] >> (v_from & 31)) & 1u) succ_ids[ns++] = nv + r;
    succ_ids[ns++] = v_from;
    double *Qt = slow ? (double *)malloc(sizeof(double) * d * N) : NULL;
    for (int k = 0; k < ns; ++k) {
        const double *qp = nodes + (size_t)d * succ_ids[k];
        if (!slow) {
            succ_collide[k] = (uint8_t)orc_collide_config_move(c, Q, qp, agent);       /* :206 */
        } else {
            int64_t fp;
            memcpy(Qt, Q, sizeof(double) * d * N);
            memcpy(Qt + (size_t)d * agent, qp, sizeof(double) * d);
            succ_collide[k] = (uint8_t)orc_collide_configs(c, Q, Qt, &fp);             /* :207 */
        }
    }
    free(Qt);
    *n_succ = ns;
    return n - nv;
}
