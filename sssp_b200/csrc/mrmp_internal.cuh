// mrmp_internal.cuh -- shared between the kernel translation units and abi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "geom.cuh"

namespace mrmp {

// Device-side view of one problem instance.  `blob` is one contiguous array of doubles in
// HBM; offsets are in doubles.  Layout (each section padded to a multiple of 2 doubles so
// that every section start is 16-byte aligned for the 1-D TMA bulk copy):
//   [CONNECT SEGMENT]  obs SoA: x[mp] y[mp] (z[mp]) r[mp] | rads[np] | base[wd*np] | aux[np] |
//                      t2_obs[N or 1][mp] | g_obs[N or 1][mp]
//   [COLLIDE SEGMENT]  rads[np] | base[wd*np] | aux[np] | t2_pair[N][N] | g_pair[N][N]
//                      (pair tables when the pair threshold is rads[i] + rads[j])
// t2_* = sq_lt_threshold(thr) (sqrt-free exact compare), g_* = broad-phase bound; the
// thresholds themselves (o.r + rads[i], rads[i] + rads[j]) are recomputed on the fly where
// the rare slow path needs them.
struct CtxView {
    const double *blob;
    int model, d, n_agents, n_obs;  // d = state dimension
    int wd;                         // workspace dimension (obstacle centres, joint points)
    int mp, np;          // padded obstacle / agent counts
    // connect segment
    int con_off, con_len;  // segment [con_off, con_off+con_len) is staged to smem
    int obs_off;           // SoA base: coordinate k of obstacle m at obs_off + k*mp + m; radius at k = wd
    int rads_off, base_off;     // base: wd doubles per agent (arm22, arm33)
    int aux_off;                // capsule3d axis lengths (axises[i])
    int t2_obs_off, g_obs_off;  // [agent*mp + m]
    // collide segment
    int col_off, col_len;
    int crads_off, cbase_off, caux_off;
    int has_pair_tab;
    int t2_pair_off, g_pair_off;  // [i*N + j]
    double step_dist, safety_dist, max_dist;
    double safety_t2;  // sq_lt_threshold(safety_dist)
    // D-independent half of Julia's rational lift of 0:step_dist:D (arm_model.cuh: step_rat)
    int step_rat_ok;
    long long step_rat_num, step_rat_den;
};

constexpr int kMaxStageBytes = 96 * 1024;  // larger segments are read from global (L1/L2)

// ---- 1-D TMA bulk copy global -> shared with mbarrier completion (UBLKCP in SASS) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Stage `n_doubles` (multiple of 2) from gsrc (16-byte aligned) into smem_dst.  All threads
// of the CTA must call it; returns after the data is visible to every thread.
__device__ __forceinline__ void tma_stage(double *smem_dst, const double *gsrc, int n_doubles,
                                          uint64_t *bar) {
    const uint32_t bytes = static_cast<uint32_t>(n_doubles) * 8u;
    const uint32_t bar_a = smem_u32(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a),
                     "r"(bytes)
                     : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                "r"(smem_u32(smem_dst)),
            "l"(gsrc), "r"(bytes), "r"(bar_a)
            : "memory");
    }
    // every thread waits on phase 0 of the barrier
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar_a)
        : "memory");
}

// Returns a pointer through which segment-relative offsets can be read: the staged smem copy
// when dyn_smem_bytes > 0, else the blob in global memory.  smem layout: [bar(16B) | data].
__device__ __forceinline__ const double *stage_segment(const CtxView &cv, int seg_off, int seg_len,
                                                       unsigned char *smem_raw, bool staged) {
    if (!staged) return cv.blob + seg_off;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    double *dst = reinterpret_cast<double *>(smem_raw + 16);
    tma_stage(dst, cv.blob + seg_off, seg_len, bar);
    return dst;
}

// ---- launch bookkeeping ------------------------------------------------------------
struct LaunchCfg {
    int sm_count;
    cudaStream_t stream;
    unsigned long long *stats;  // ctx counters [4] (may be NULL): [0] cross-kernel survivors,
                                // [1] arm link pairs that ran the exact test
};

void count_launch(int n = 1);

// one cross-product request (abi.cu: cross_run_dev): output mode and how columns map to slots
struct CrossReq {
    int mode;                  // kCrossBits / kCrossOr / kCrossCount / kCrossFirst
    int group_size;            // > 0: group = column / group_size (unless col_group)
    int n_groups;              // count / first: output slots per row
    const int32_t *col_group;  // device arrays or NULL
    const int32_t *col_key;
    const int32_t *row_key;    // only pairs with col_key[c] > row_key[r] are tested
    int cbs;                   // cbs.jl:343-346: vertex test OR edge test, column agent first
};
constexpr int kCrossBits = 0;   // full bit matrix (hits OR-ed into a zeroed output)
constexpr int kCrossOr = 1;     // OR over column groups, one bit per (row, group)
constexpr int kCrossCount = 2;  // int32 count per (row, group)
constexpr int kCrossFirst = 3;  // int32 smallest colliding column index per (row, group), -1 = none

// kernels_point.cu
cudaError_t launch_connect_points(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                  const int32_t *agent, const double *q, uint8_t *out);
cudaError_t launch_connect_edges(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                 const int32_t *agent, const double *qf, const double *qt,
                                 uint8_t *out);
cudaError_t launch_collide_pairs(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                 const int32_t *ai, const int32_t *aj, const double *qif,
                                 const double *qit, const double *qjf, const double *qjt,
                                 uint8_t *out);
cudaError_t launch_collide_static_pairs(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                        const int32_t *ai, const int32_t *aj, const double *qi,
                                        const double *qj, uint8_t *out);
cudaError_t launch_collide_configs(const CtxView &cv, const LaunchCfg &lc, int64_t n_cfg,
                                   const double *Qf, const double *Qt, uint8_t *out,
                                   int64_t *first_pair);
cudaError_t launch_collide_config_moves(const CtxView &cv, const LaunchCfg &lc, int agent_i,
                                        const double *Q, int64_t n_cand, const double *q_to,
                                        uint8_t *out);
cudaError_t launch_collide_edges_indexed(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                         const int32_t *ai, const int32_t *aj,
                                         const int32_t *vif, const int32_t *vit,
                                         const int32_t *vjf, const int32_t *vjt,
                                         const double *nodes, int agent0, int n_local,
                                         int nv_stride, uint8_t *out);

// kernels_arm.cu (StateArm22; the launchers above forward here when cv.model == 2)
cudaError_t launch_arm_connect_points(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                      const int32_t *agent, const double *q, uint8_t *out);
cudaError_t launch_arm_connect_edges(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                     const int32_t *agent, const double *qf, const double *qt,
                                     uint8_t *out);
cudaError_t launch_arm_collide_static_pairs(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                            const int32_t *ai, const int32_t *aj,
                                            const double *qi, const double *qj, uint8_t *out);
cudaError_t launch_arm_collide_pairs(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                     const int32_t *ai, const int32_t *aj, const double *qif,
                                     const double *qit, const double *qjf, const double *qjt,
                                     uint8_t *out, unsigned long long *stats);
cudaError_t launch_arm_collide_edges_indexed(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                             const int32_t *ai, const int32_t *aj,
                                             const int32_t *vif, const int32_t *vit,
                                             const int32_t *vjf, const int32_t *vjt,
                                             const double *nodes, int agent0, int nv_stride,
                                             uint8_t *out, unsigned long long *stats);
cudaError_t launch_arm_collide_cross(const CtxView &cv, const LaunchCfg &lc, int64_t n_rows,
                                     const int32_t *row_agent, const double *row_from,
                                     const double *row_to, int64_t n_cols,
                                     const int32_t *col_agent, const double *col_from,
                                     const double *col_to, const int32_t *col_vf,
                                     const int32_t *col_vt, const double *nodes, int agent0,
                                     int nv_stride, const CrossReq &rq, int out_stride,
                                     uint32_t *out, unsigned long long *stats);
// first_scratch: n_cfg words of device scratch
cudaError_t launch_arm_collide_configs(const CtxView &cv, const LaunchCfg &lc, int64_t n_cfg,
                                       const double *Qf, const double *Qt, uint8_t *out,
                                       int64_t *first_pair, unsigned long long *first_scratch,
                                       unsigned long long *stats);
cudaError_t launch_arm_collide_config_moves(const CtxView &cv, const LaunchCfg &lc, int agent_i,
                                            const double *Q, int64_t n_cand, const double *q_to,
                                            uint8_t *out, unsigned long long *stats);

// kernels_chain.cu (line2d / snake2d / arm33 / capsule3d; forwarded to when cv.model >= 3)
cudaError_t launch_chain_connect_points(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                      const int32_t *agent, const double *q, uint8_t *out);
cudaError_t launch_chain_connect_edges(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                     const int32_t *agent, const double *qf, const double *qt,
                                     uint8_t *out);
cudaError_t launch_chain_collide_static_pairs(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                            const int32_t *ai, const int32_t *aj,
                                            const double *qi, const double *qj, uint8_t *out);
cudaError_t launch_chain_collide_pairs(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                     const int32_t *ai, const int32_t *aj, const double *qif,
                                     const double *qit, const double *qjf, const double *qjt,
                                     uint8_t *out, unsigned long long *stats);
cudaError_t launch_chain_collide_edges_indexed(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                             const int32_t *ai, const int32_t *aj,
                                             const int32_t *vif, const int32_t *vit,
                                             const int32_t *vjf, const int32_t *vjt,
                                             const double *nodes, int agent0, int nv_stride,
                                             uint8_t *out, unsigned long long *stats);
cudaError_t launch_chain_collide_cross(const CtxView &cv, const LaunchCfg &lc, int64_t n_rows,
                                     const int32_t *row_agent, const double *row_from,
                                     const double *row_to, int64_t n_cols,
                                     const int32_t *col_agent, const double *col_from,
                                     const double *col_to, const int32_t *col_vf,
                                     const int32_t *col_vt, const double *nodes, int agent0,
                                     int nv_stride, const CrossReq &rq, int out_stride,
                                     uint32_t *out, unsigned long long *stats);
// first_scratch: n_cfg words of device scratch
cudaError_t launch_chain_collide_configs(const CtxView &cv, const LaunchCfg &lc, int64_t n_cfg,
                                       const double *Qf, const double *Qt, uint8_t *out,
                                       int64_t *first_pair, unsigned long long *first_scratch,
                                       unsigned long long *stats);
cudaError_t launch_chain_collide_config_moves(const CtxView &cv, const LaunchCfg &lc, int agent_i,
                                            const double *Q, int64_t n_cand, const double *q_to,
                                            uint8_t *out, unsigned long long *stats);

// kernels_prm.cu
// adjacency bitmap [n_agents*nv rows][words] + row counts, exclusive scan, CSR emit
cudaError_t launch_prm_adjacency(const CtxView &cv, const LaunchCfg &lc, int agent0, int n_agents,
                                 int nv, const double *rad_le /*[n] R2 or NaN*/,
                                 const double *rad /*[n]*/, const double *nodes,
                                 uint32_t *bitmap, int words, int32_t *row_cnt,
                                 const double *rad_host = nullptr, const double *r2_host = nullptr);
cudaError_t launch_row_scan(const LaunchCfg &lc, int64_t n_rows, const int32_t *row_cnt,
                            int32_t *row_ptr, int64_t *nnz_total, long long *chunk_ws,
                            int64_t chunk_ws_len, int64_t *nnz_mirror = nullptr);
cudaError_t launch_prm_emit(const LaunchCfg &lc, int64_t n_rows, int nv, const uint32_t *bitmap,
                            int words, const int32_t *row_ptr, int32_t *col_idx, int64_t col_cap);
// uniform-grid spatial hash path (point models, large nv): bin -> adjacency -> gather
cudaError_t launch_grid_bin(const CtxView &cv, const LaunchCfg &lc, int n_agents, int nv, int g,
                            int cells, const double *nodes, int32_t *cell_start, int32_t *cursor,
                            int32_t *ids);
cudaError_t launch_prm_grid_adjacency(const CtxView &cv, const LaunchCfg &lc, int agent0,
                                      int n_agents, int nv, int g, int cells,
                                      const int32_t *cell_start, const int32_t *ids,
                                      const double *rad_r2, const double *rad, const double *nodes,
                                      int32_t *row_cnt, int64_t *row_off, int32_t *tmp,
                                      int64_t tmp_cap, unsigned long long *tmp_cursor);
cudaError_t launch_grid_gather_rows(const LaunchCfg &lc, int64_t n_rows, const int32_t *row_ptr,
                                    const int64_t *row_off, const int32_t *tmp, int32_t *col_idx,
                                    int64_t col_cap);
// CSR shards of a multi-GPU build: slot = int32 cnt[rows_cap] | int32 col[cap_nnz]
cudaError_t launch_csr_pack_shard(const LaunchCfg &lc, int64_t n_rows, int64_t rows_cap,
                                  const int32_t *row_ptr, const int32_t *col_idx, int32_t *slot);
struct CsrPushDst {  // destinations of one shard push (own buffer and peer mappings)
    int n;
    int32_t *slot[8];
};
cudaError_t launch_csr_pack_push(const LaunchCfg &lc, int64_t n_rows, int64_t rows_cap,
                                 int64_t cap_nnz, const int32_t *row_ptr, const int32_t *col_idx,
                                 const CsrPushDst &dst);
cudaError_t launch_csr_shard_counts(const LaunchCfg &lc, int64_t n_rows, int64_t rows_cap,
                                    int64_t slot_stride, const int32_t *slots, int32_t *row_cnt);
cudaError_t launch_csr_stitch(const LaunchCfg &lc, int64_t n_rows, int64_t rows_cap,
                              int64_t slot_stride, const int32_t *slots, const int32_t *row_ptr,
                              int32_t *col_idx, int64_t col_cap);
cudaError_t launch_sample_states(const CtxView &cv, const LaunchCfg &lc, int agent, uint64_t seed,
                                 uint64_t t0, int64_t n, double *q_out);
// per agent: nodes[a][0]=init, [1]=goal, then the first nv-2 free samples in try order
cudaError_t launch_sample_free(const CtxView &cv, const LaunchCfg &lc, int agent0, int n_agents,
                               int n_fixed, int nv, int nv_stride, uint64_t seed, double *nodes,
                               int64_t *tries_out, int64_t max_tries,
                               const double *fixed_src = nullptr);


// kernels_sssp.cu
constexpr int kMaxExpand = 64;  // samples per expansion call (num_vertex_expansion)
struct SsspView {               // growing per-agent node tables of one SSSP instance
    double *nodes;              // [N][cap][d]
    int32_t *nv;                // [N] current vertex counts
    int cap;
};
struct SsspExpand {             // scalar arguments of one expansion
    int agent, v_from, K, depth;
    double eps, eps_r2;         // epsilon (NaN = nothing), sq_le_threshold(epsilon)
    double thr, thr_r2;         // min_dist_thread, sq_le_threshold(min_dist_thread)
    unsigned long long seed, t0;
    int host_samples;           // q_rand uploaded by the caller instead of (seed, t0 + k)
    int words;                  // bitmap words per new vertex: ceil((nv_before + K) / 32)
    int n_old;                  // old neighbours of v_from
    int slow_collide;           // no_fast_collision_check: collide(S.Q, Q) instead of (S.Q, q, i)
    int seq;                    // fused path: written to hdr[3] after every result has landed in
                                // mapped host memory (the host polls it instead of a stream wait)
};
struct SsspWs {                 // device workspace of one expansion (laid out by abi_sssp.cu)
    const double *q_rand;       // [K][d]
    const int32_t *Q_ids;       // [N] vertex of every agent in the search node S
    const int32_t *old_nbrs;    // [n_old]
    int32_t *hdr;               // [0] n_new, [1] n_succ, [2] nv_before
    double *q_new;              // [K][d] accepted vertices, in acceptance order
    uint32_t *out_bits;         // [K][words]; in_bits = out_bits + K*words follows directly
    uint32_t *in_bits;
    int32_t *succ_ids;          // [n_old + K + 1]
    uint8_t *succ_out;          // [n_old + K + 1] collide verdict per successor
    double *Qbuf;               // [N][d]
    double *cand;               // [n_old + K + 1][d]
    double *cfg_from, *cfg_to;  // slow_collide: [n_old + K + 1][N][d]
    int64_t *cfg_first;
};
cudaError_t launch_sssp_expand(const CtxView &cv, const LaunchCfg &lc, const SsspView &view,
                               const SsspExpand &a, const SsspWs &ws, int nv_before);
// small inputs of one expansion travel in the kernel parameter space (no host-memory reads
// on the kernel's critical path)
constexpr int kFusedMaxAgents = 128, kFusedMaxOld = 128, kFusedMaxK = 32;
struct SsspFusedIn {
    int32_t q_ids[kFusedMaxAgents];
    int32_t old_nbrs[kFusedMaxOld];
    double q_rand[kFusedMaxK * 3];
};
size_t sssp_fused_smem(const CtxView &cv, const SsspExpand &a);
cudaError_t launch_sssp_expand_fused(const CtxView &cv, const LaunchCfg &lc, const SsspView &view,
                                     const SsspExpand &a, const SsspFusedIn &in, int32_t *h_hdr,
                                     double *h_qnew, uint32_t *h_bits, int32_t *h_succ,
                                     uint8_t *h_verdict);
// v_off[n_graphs+1] vertex offsets; row_ptr over all vertices, col_idx local to each graph;
// w = scratch for one double per edge; table[total vertices]
cudaError_t launch_distance_tables(const CtxView &cv, const LaunchCfg &lc, int n_graphs,
                                   const int32_t *v_off, const int32_t *row_ptr,
                                   const int32_t *col_idx, const double *nodes,
                                   const int32_t *goal, double stay_penalty, double *w, double *table);
cudaError_t launch_state_op(const CtxView &cv, const LaunchCfg &lc, int op, int64_t n,
                            const double *q1, const double *q2, double *out);
cudaError_t launch_rrt_extend(const CtxView &cv, const LaunchCfg &lc, const SsspExpand &a, int n_v,
                              const double *V, const double *q_rand, int from_start,
                              int32_t *out_flag, int32_t *out_near, double *out_q);
cudaError_t launch_nearest(const CtxView &cv, const LaunchCfg &lc, int n_v, const double *V,
                           const double *q, int reversed, int32_t *idx_out, double *dist_out);

// kernels_cross.cu
constexpr int kCrossCols = 32;        // columns per tile
constexpr int kCrossMaxGroups = 256;  // OR mode: groups per launch
struct CrossCols {              // how the columns map to output slots
    int group_size;             // > 0: group = column / group_size (unless col_group)
    const int32_t *col_group;   // explicit group id per column (device) or NULL
    const int32_t *col_key;     // per-column key (device) or NULL
    int64_t orig_base;          // index of the first column of this launch in the caller's order
};
struct CrossOut {
    int mode;
    int words_per_row;        // OR mode: words written per row
    int64_t out_stride;       // 32-bit words (bits / or) or ints (count / first) per row
    uint32_t *out;
    const int32_t *row_key;   // NULL, or: only pairs with col_key > row_key[row] are tested
    int cbs;                  // cbs.jl:343-346 pair test (vertex OR edge, column first)
    int may_split;            // `out` is zeroed memory of this GPU: a warp tile's columns may be
                              // divided over several work items that OR / add into it
    const long long *n_rows_dev;  // NULL, or the row count on the device (<= the host bound
                                  // passed as n_rows): a table launched behind an asynchronous build
};
int cross_split(const LaunchCfg &lc, int64_t n_rows, int n_tiles, int may_split);
size_t cross_tile_bytes(int d);
size_t cross_hdr_bytes(int d);
size_t cross_row_bytes(int d);
// rows -> Morton-sorted row records; cells = 4096 ints of device scratch
cudaError_t launch_cross_sort_rows(const CtxView &cv, const LaunchCfg &lc, int64_t n_rows,
                                   const int32_t *agent, const double *qf, const double *qt,
                                   int32_t *cells, void *rows_out);
// columns (explicit states, or ids vf/vt into `nodes`) -> Morton-sorted tiles + headers
cudaError_t launch_cross_prep_cols(const CtxView &cv, const LaunchCfg &lc, int64_t n_cols,
                                   const int32_t *agent, const double *qf, const double *qt,
                                   const int32_t *vf, const int32_t *vt, const double *nodes,
                                   int agent0, int nv_stride, const CrossCols &cc, int32_t *cells,
                                   int32_t *order, void *tiles, void *hdrs);
cudaError_t launch_cross_sort_csr_rows(const CtxView &cv, const LaunchCfg &lc, int64_t n_rows_csr,
                                       int nv, int agent0, const int32_t *row_ptr,
                                       const int32_t *col_idx, const double *nodes, int32_t *cells,
                                       void *rows_out, int64_t e_cap);
cudaError_t launch_csr_expand_edges(const CtxView &cv, const LaunchCfg &lc, int64_t n_rows_csr,
                                    int nv, int agent0, const int32_t *row_ptr,
                                    const int32_t *col_idx, const double *nodes, int32_t *e_agent,
                                    double *e_from, double *e_to);
cudaError_t launch_cross_collide(const CtxView &cv, const LaunchCfg &lc, int64_t n_rows,
                                 const void *rows, int64_t n_cols, const void *tiles,
                                 const void *hdrs, const CrossOut &co, unsigned long long *stats,
                                 unsigned int *work);
cudaError_t launch_fill_i32(const LaunchCfg &lc, int32_t *out, int64_t n, int32_t v);
cudaError_t launch_cross_first_finalize(const LaunchCfg &lc, int32_t *out, int64_t n);

}  // namespace mrmp
