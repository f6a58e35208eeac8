/*
 * mrmp_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's (Kei18/sssp, MRMP.jl) connect /
 * collide / roadmap hot path, IEEE-754 binary64, no FMA contraction
 * (compile with -ffp-contract=off -fno-fast-math).  Every function cites the
 * reference file:line it follows (paths relative to /root/reference).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (sssp_b200/)
 * never links, imports or calls it.
 *
 * PARITY PINNING: the reference cannot be executed here (no Julia runtime in
 * the image, none on the GPU box).  The reference's own tests hold exactly ONE
 * known-answer vector for this path (test/dist.jl:4-9, segment-segment == 0);
 * the oracle is pinned against it (tests/test_oracle_kat.py).  Everything else
 * (verdicts, edge sets, arm22 trig, float-range schedule) is "parity unpinned":
 * the reference SOURCE is the specification, restated here and cross-checked by
 * an independent literal Python transliteration (oracle/pyref.py).
 */
#ifndef MRMP_ORACLE_H
#define MRMP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_POINT2D = 0, ORC_POINT3D = 1, ORC_ARM22 = 2, ORC_LINE2D = 3, ORC_SNAKE2D = 4,
       ORC_ARM33 = 5, ORC_CAPSULE3D = 6 };

typedef struct orc_ctx {
    int model;          /* ORC_* */
    int d;              /* state dimension: 2, 3, 2, 3, 6, 6, 6 */
    int wd;             /* workspace dimension (obstacle centre dim): 2, 3, 2, 2, 2, 3, 3 */
    int n_agents;
    double *rads;       /* [n_agents]  agent radius (points) / link length (arm22) */
    double *base;       /* [n_agents*wd] arm22 / arm33 base positions, else NULL */
    double *aux;        /* [n_agents] capsule3d axis lengths (axises), else NULL */
    int n_obs;
    double *obs;        /* AoS [n_obs*(wd+1)]: x,y,(z),r  (src/obstacle.jl:7-19) */
    double step_dist;   /* src/MRMP.jl:65 */
    double safety_dist; /* src/MRMP.jl:66 */
    double max_dist;    /* NaN == nothing */
} orc_ctx;

orc_ctx *orc_ctx_create(int model, int n_agents, const double *rads, const double *base,
                        int n_obs, const double *obs_aos, double step_dist,
                        double safety_dist, double max_dist);
/* same with the second per-agent parameter of capsule3d (axises, capsule3d.jl:61-68) */
orc_ctx *orc_ctx_create2(int model, int n_agents, const double *rads, const double *base,
                         const double *aux, int n_obs, const double *obs_aos, double step_dist,
                         double safety_dist, double max_dist);
void orc_ctx_destroy(orc_ctx *c);

/* geometry: src/utils.jl:17-99 */
double orc_norm(const double *v, int d);
double orc_dot(const double *a, const double *b, int d);
double orc_dist_pp(const double *a, const double *b, int d);
double orc_dist_sp(const double *a, const double *b, const double *c, int d);
double orc_dist_ss(const double *p11, const double *p12, const double *p21,
                   const double *p22, int d);
double orc_diff_angles(double t1, double t2);
/* state-space dist(q1,q2): point2d.jl:12-14, point3d.jl:13-15, arm22.jl:15-20 */
double orc_state_dist(const orc_ctx *c, const double *q1, const double *q2);
/* get_mid_status: point2d.jl:8-10, point3d.jl:9-11, arm22.jl:8-13 */
void orc_mid_status(const orc_ctx *c, const double *p, const double *q, double *out);

/* trig used by arm22 (oracle/orc_trig.c: msun-style, see header there) */
double orc_sin(double x);
double orc_cos(double x);
double orc_atan2(double y, double x);

/* Julia float range 0:step:D, then ./D, then ++[1.0]  (arm22.jl:65).  Writes the
 * e-values into out (capacity cap), returns the count (>= 2 incl. trailing 1.0). */
int orc_step_schedule(double step, double D, double *out, int cap);

/* closure bodies */
int orc_connect_point(const orc_ctx *c, const double *q, int i);
int orc_connect_edge(const orc_ctx *c, const double *qf, const double *qt, int i);
int orc_collide_pair(const orc_ctx *c, const double *qif, const double *qit,
                     const double *qjf, const double *qjt, int i, int j);
int orc_collide_static(const orc_ctx *c, const double *qi, const double *qj, int i, int j);
/* collide(Q_from,Q_to): returns verdict; *first_pair = i*N+j of the first hit or -1 */
int orc_collide_configs(const orc_ctx *c, const double *Qf, const double *Qt,
                        int64_t *first_pair);
int orc_collide_config_move(const orc_ctx *c, const double *Q, const double *q_to, int i);

/* batched drivers (OpenMP over checks; nthreads<=0 -> omp default) */
void orc_connect_points_batch(const orc_ctx *c, int64_t n, const int32_t *agent,
                              const double *q, uint8_t *out, int nthreads);
void orc_connect_edges_batch(const orc_ctx *c, int64_t n, const int32_t *agent,
                             const double *qf, const double *qt, uint8_t *out, int nthreads);
void orc_collide_pairs_batch(const orc_ctx *c, int64_t n, const int32_t *ai, const int32_t *aj,
                             const double *qif, const double *qit, const double *qjf,
                             const double *qjt, uint8_t *out, int nthreads);
void orc_collide_static_pairs_batch(const orc_ctx *c, int64_t n, const int32_t *ai,
                                    const int32_t *aj, const double *qi, const double *qj,
                                    uint8_t *out, int nthreads);
void orc_collide_configs_batch(const orc_ctx *c, int64_t n_cfg, const double *Qf,
                               const double *Qt, uint8_t *out, int64_t *first_pair,
                               int nthreads);
void orc_collide_config_moves_batch(const orc_ctx *c, int agent_i, const double *Q,
                                    int64_t n_cand, const double *q_to, uint8_t *out,
                                    int nthreads);
/* roadmap-indexed collide: checks reference (agent, vertex ids) of node tables
 * nodes[agent*nv_stride + v] (d doubles each) */
void orc_collide_edges_indexed_batch(const orc_ctx *c, int64_t n, const int32_t *ai,
                                     const int32_t *aj, const int32_t *vif, const int32_t *vit,
                                     const int32_t *vjf, const int32_t *vjt,
                                     const double *nodes, int64_t nv_stride, uint8_t *out,
                                     int nthreads);

void orc_collide_cross(const orc_ctx *c, int64_t n_rows, const int32_t *ra, const double *rf,
                       const double *rt, int64_t n_cols, const int32_t *ca, const double *cf,
                       const double *ct, int group_size, uint32_t *out, int nthreads);
/* count / first reductions of the cross product: cbs.jl:336-347, smoother.jl:78-106 */
void orc_collide_cross_reduce(const orc_ctx *c, int64_t n_rows, const int32_t *ra, const double *rf,
                              const double *rt, const int32_t *row_key, int64_t n_cols,
                              const int32_t *ca, const double *cf, const double *ct,
                              const int32_t *col_group, const int32_t *col_key, int group_size,
                              int n_groups, int mode, int cbs, int32_t *out, int nthreads);

/* counter-based sampler (Philox4x32-10), replaces Julia's task-local RNG:
 * gen_uniform_sampling, point2d.jl:71-75, point3d.jl:82-86, arm22.jl:214-218 */
void orc_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                    uint32_t k1, uint32_t out[4]);
void orc_sample_state(const orc_ctx *c, uint64_t seed, int agent, uint64_t t, double *q);
/* PRM! sampling loop prm.jl:193-199: first n_wanted free samples in try order */
int64_t orc_sample_free(const orc_ctx *c, int agent, uint64_t seed, int64_t n_wanted,
                        double *q_out, int64_t max_tries);

/* PRM! neighbour pass prm.jl:202-210 for one agent; rad NaN == nothing.
 * row_ptr[nv+1], col_idx capacity cap; returns nnz (may exceed cap: then truncated) */
int64_t orc_prm_build_agent(const orc_ctx *c, int agent, int nv, double rad,
                            const double *nodes, int32_t *row_ptr, int32_t *col_idx,
                            int64_t cap);
/* PRMs prm.jl:251-276: agents agent0..agent0+n-1, nodes [n][nv][d], row_ptr [n][nv+1],
 * col_idx [n][cap]; nnz_out [n] */
void orc_prm_build(const orc_ctx *c, int agent0, int n, int nv, const double *rad,
                   const double *nodes, int32_t *row_ptr, int32_t *col_idx, int64_t cap,
                   int64_t *nnz_out, int nthreads);

/* steering sssp.jl:269-291 with conn sssp.jl:87-89 (epsilon NaN == nothing) */
void orc_steering(const orc_ctx *c, int agent, const double *q_rand, const double *q_near,
                  int depth, double epsilon, double *q_out);
int orc_conn(const orc_ctx *c, int agent, const double *qf, const double *qt, double epsilon);

/* Dijkstra distance table solvers/utils.jl:146-179 over CSR (0-based ids) from `goal` */
void orc_distance_table(const orc_ctx *c, int nv, const double *nodes, const int32_t *row_ptr,
                        const int32_t *col_idx, int goal, double *table);

/* one SSSP node expansion (expand! sssp.jl:235-267 + successor filter :196-223); see the
 * definition for the argument layout */
int orc_sssp_expand(const orc_ctx *c, int agent, double *nodes, int nv, int v_from, int K,
                    const double *q_rand, int depth, double epsilon, double thr, int words,
                    uint32_t *out_bits, uint32_t *in_bits, const double *Q, int n_old,
                    const int32_t *old_nbrs, int slow, int32_t *succ_ids, uint8_t *succ_collide,
                    int *n_succ);

/* orc_chain.c: line2d / snake2d / arm33 / capsule3d (dispatched to by the functions above) */
int orc_chain_state_dim(int model);
int orc_chain_workspace_dim(int model);
double orc_chain_dist(const orc_ctx *c, const double *q1, const double *q2);
void orc_chain_mid_status(const orc_ctx *c, const double *p, const double *q, double *out);
int orc_chain_connect_point(const orc_ctx *c, const double *q, int i);
int orc_chain_connect_edge(const orc_ctx *c, const double *qf, const double *qt, int i);
int orc_chain_collide_static(const orc_ctx *c, const double *qi, const double *qj, int i, int j);
int orc_chain_collide_pair(const orc_ctx *c, const double *qif, const double *qit,
                           const double *qjf, const double *qjt, int i, int j);
int orc_chain_collide_config_move(const orc_ctx *c, const double *Q, const double *q_to, int i);
void orc_chain_scale_sample(const orc_ctx *c, const double *u, double *q);

int orc_dot_is_fma(void);
int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
