// kernels_point.cu -- explicit-list connect / collide kernels for the point models.
//
// One check per thread, grid-stride, persistent grid sized in multiples of the SM count.
// Obstacles, radii and the exact-compare threshold tables are staged into shared memory by
// one 1-D TMA bulk copy per CTA (mrmp_internal.cuh: tma_stage).  States stream from HBM as
// the reference's AoS layout (double2 vector loads for 2-D states).
#include "point_model.cuh"

namespace mrmp {

extern __shared__ __align__(16) unsigned char smem_raw[];

constexpr int kThreads = 256;

static inline int seg_smem_bytes(int seg_len) {
    const int bytes = seg_len * 8 + 16;
    return bytes <= kMaxStageBytes ? bytes : 0;
}

static inline int grid_for(const LaunchCfg &lc, int64_t n, int smem_bytes, int threads = kThreads) {
    int per_sm = 2048 / threads;
    if (smem_bytes > 0) {
        int by_smem = (220 * 1024) / smem_bytes;
        if (by_smem < per_sm) per_sm = by_smem < 1 ? 1 : by_smem;
    }
    int64_t want = (n + threads - 1) / threads;
    int64_t cap = (int64_t)lc.sm_count * per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

// ---- connect(q, i) ---------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kThreads)
k_connect_points(CtxView cv, int64_t n, const int32_t *__restrict__ agent,
                 const double *__restrict__ q, uint8_t *__restrict__ out, int staged) {
    const ConSeg cs(cv, stage_segment(cv, cv.con_off, cv.con_len, smem_raw, staged));
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
         k += (int64_t)gridDim.x * blockDim.x) {
        double p[D];
        load_state<D>(q, k, p);
        out[k] = point_connect_point<D>(cs, p, __ldg(agent + k)) ? 1 : 0;
    }
}

// ---- connect(q_from, q_to, i) -----------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kThreads)
k_connect_edges(CtxView cv, int64_t n, const int32_t *__restrict__ agent,
                const double *__restrict__ qf, const double *__restrict__ qt,
                uint8_t *__restrict__ out, int staged) {
    const ConSeg cs(cv, stage_segment(cv, cv.con_off, cv.con_len, smem_raw, staged));
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
         k += (int64_t)gridDim.x * blockDim.x) {
        double a[D], b[D];
        load_state<D>(qf, k, a);
        load_state<D>(qt, k, b);
        out[k] = point_connect_edge<D>(cs, a, b, __ldg(agent + k)) ? 1 : 0;
    }
}

// ---- collide(qi_from, qi_to, qj_from, qj_to, i, j) ---------------------------------------
template <int D>
__global__ void __launch_bounds__(kThreads)
k_collide_pairs(CtxView cv, int64_t n, const int32_t *__restrict__ ai,
                const int32_t *__restrict__ aj, const double *__restrict__ qif,
                const double *__restrict__ qit, const double *__restrict__ qjf,
                const double *__restrict__ qjt, uint8_t *__restrict__ out, int staged) {
    const ColSeg cs(cv, stage_segment(cv, cv.col_off, cv.col_len, smem_raw, staged));
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
         k += (int64_t)gridDim.x * blockDim.x) {
        double a[D], b[D], c[D], d[D];
        load_state<D>(qif, k, a);
        load_state<D>(qit, k, b);
        load_state<D>(qjf, k, c);
        load_state<D>(qjt, k, d);
        out[k] = point_collide_pair<D>(cs, a, b, c, d, __ldg(ai + k), __ldg(aj + k)) ? 1 : 0;
    }
}

// ---- collide over roadmap edges referenced by (agent, vertex) ids ---------------------------
template <int D>
__global__ void __launch_bounds__(kThreads)
k_collide_edges_indexed(CtxView cv, int64_t n, const int32_t *__restrict__ ai,
                        const int32_t *__restrict__ aj, const int32_t *__restrict__ vif,
                        const int32_t *__restrict__ vit, const int32_t *__restrict__ vjf,
                        const int32_t *__restrict__ vjt, const double *__restrict__ nodes,
                        int agent0, int nv_stride, uint8_t *__restrict__ out, int staged) {
    const ColSeg cs(cv, stage_segment(cv, cv.col_off, cv.col_len, smem_raw, staged));
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
         k += (int64_t)gridDim.x * blockDim.x) {
        const int i = __ldg(ai + k), j = __ldg(aj + k);
        const int64_t bi = (int64_t)(i - agent0) * nv_stride, bj = (int64_t)(j - agent0) * nv_stride;
        double a[D], b[D], c[D], d[D];
        load_state<D>(nodes, bi + __ldg(vif + k), a);
        load_state<D>(nodes, bi + __ldg(vit + k), b);
        load_state<D>(nodes, bj + __ldg(vjf + k), c);
        load_state<D>(nodes, bj + __ldg(vjt + k), d);
        out[k] = point_collide_pair<D>(cs, a, b, c, d, i, j) ? 1 : 0;
    }
}

// ---- collide(q_i, q_j, i, j) ----------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kThreads)
k_collide_static_pairs(CtxView cv, int64_t n, const int32_t *__restrict__ ai,
                       const int32_t *__restrict__ aj, const double *__restrict__ qi,
                       const double *__restrict__ qj, uint8_t *__restrict__ out, int staged) {
    const ColSeg cs(cv, stage_segment(cv, cv.col_off, cv.col_len, smem_raw, staged));
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
         k += (int64_t)gridDim.x * blockDim.x) {
        double a[D], b[D];
        load_state<D>(qi, k, a);
        load_state<D>(qj, k, b);
        out[k] = point_collide_static<D>(cs, a, b, __ldg(ai + k), __ldg(aj + k)) ? 1 : 0;
    }
}

// ---- collide(Q_from, Q_to): one warp per configuration pair --------------------------------
// Pairs (i < j) are enumerated in the reference's loop order (point.jl:19); lanes take
// consecutive pair ranks, so the first chunk with a hit holds the first hit.
template <int D>
__global__ void __launch_bounds__(kThreads)
k_collide_configs(CtxView cv, int64_t n_cfg, const double *__restrict__ Qf,
                  const double *__restrict__ Qt, uint8_t *__restrict__ out,
                  int64_t *__restrict__ first_pair, int staged) {
    const ColSeg cs(cv, stage_segment(cv, cv.col_off, cv.col_len, smem_raw, staged));
    const int N = cv.n_agents;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t P = (int64_t)N * (N - 1) / 2;
    for (int64_t cfg = warp; cfg < n_cfg; cfg += n_warps) {
        const double *qf = Qf + cfg * N * D, *qt = Qt + cfg * N * D;
        // lane's first pair: rank = lane
        int i = 0, j = 1 + lane;
        while (i < N - 1 && j >= N) {
            j = j - N + i + 2;  // carry into the next row: row i has N-1-i pairs
            ++i;
        }
        int64_t found = -1;
        for (int64_t base = 0; base < P; base += 32) {
            bool hit = false;
            if (base + lane < P) {
                double a[D], b[D], c[D], d[D];
                load_state<D>(qf, i, a);
                load_state<D>(qt, i, b);
                load_state<D>(qf, j, c);
                load_state<D>(qt, j, d);
                hit = point_collide_pair<D>(cs, a, b, c, d, i, j);
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m) {
                const int src = __ffs(m) - 1;
                const int fi = __shfl_sync(0xffffffffu, i, src);
                const int fj = __shfl_sync(0xffffffffu, j, src);
                found = (int64_t)fi * N + fj;
                break;
            }
            // advance this lane by 32 ranks
            j += 32;
            while (i < N - 1 && j >= N) {
                j = j - N + i + 2;
                ++i;
            }
        }
        if (lane == 0) {
            out[cfg] = found >= 0 ? 1 : 0;
            if (first_pair) first_pair[cfg] = found;
        }
    }
}

// ---- collide(Q, q_to, i): thread per candidate q_to ----------------------------------------
template <int D>
__global__ void __launch_bounds__(kThreads)
k_collide_config_moves(CtxView cv, int agent_i, const double *__restrict__ Q, int64_t n_cand,
                       const double *__restrict__ q_to, uint8_t *__restrict__ out, int staged) {
    const ColSeg cs(cv, stage_segment(cv, cv.col_off, cv.col_len, smem_raw, staged));
    const int N = cv.n_agents;
    double qf[D];
    load_state<D>(Q, agent_i, qf);
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n_cand;
         k += (int64_t)gridDim.x * blockDim.x) {
        double qt[D], u[D];
        load_state<D>(q_to, k, qt);
#pragma unroll
        for (int c = 0; c < D; ++c) u[c] = dsub(qf[c], qt[c]);
        const double uu = dotd<D>(u, u);
        bool hit = false;
        for (int j = 0; j < N; ++j) {
            if (j == agent_i) continue;
            double qj[D];
            load_state<D>(Q, j, qj);
            hit |= point_collide_move_vs_static<D>(cs, qf, qt, u, uu, qj, agent_i, j);
        }
        out[k] = hit ? 1 : 0;
    }
}

// ---- launchers ----------------------------------------------------------------------------
#define MRMP_DISPATCH_D(cv, CALL2, CALL3)                       \
    do {                                                        \
        if ((cv).model == 0) {                                  \
            CALL2;                                              \
        } else if ((cv).model == 1) {                           \
            CALL3;                                              \
        } else {                                                \
            return cudaErrorNotSupported;                       \
        }                                                       \
    } while (0)

template <class K>
static cudaError_t set_smem(K kernel, int bytes) {
    if (bytes > 48 * 1024)
        return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    return cudaSuccess;
}

cudaError_t launch_connect_points(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                  const int32_t *agent, const double *q, uint8_t *out) {
    if (n <= 0) return cudaSuccess;
    if (cv.model == 2) return launch_arm_connect_points(cv, lc, n, agent, q, out);
    if (cv.model >= 3) return launch_chain_connect_points(cv, lc, n, agent, q, out);
    const int sm = seg_smem_bytes(cv.con_len);
    const int grid = grid_for(lc, n, sm);
    cudaError_t e;
    MRMP_DISPATCH_D(
        cv,
        (e = set_smem(k_connect_points<2>, sm),
         k_connect_points<2><<<grid, kThreads, sm, lc.stream>>>(cv, n, agent, q, out, sm > 0)),
        (e = set_smem(k_connect_points<3>, sm),
         k_connect_points<3><<<grid, kThreads, sm, lc.stream>>>(cv, n, agent, q, out, sm > 0)));
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_connect_edges(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                 const int32_t *agent, const double *qf, const double *qt,
                                 uint8_t *out) {
    if (n <= 0) return cudaSuccess;
    if (cv.model == 2) return launch_arm_connect_edges(cv, lc, n, agent, qf, qt, out);
    if (cv.model >= 3) return launch_chain_connect_edges(cv, lc, n, agent, qf, qt, out);
    const int sm = seg_smem_bytes(cv.con_len);
    const int grid = grid_for(lc, n, sm);
    cudaError_t e;
    MRMP_DISPATCH_D(
        cv,
        (e = set_smem(k_connect_edges<2>, sm),
         k_connect_edges<2><<<grid, kThreads, sm, lc.stream>>>(cv, n, agent, qf, qt, out, sm > 0)),
        (e = set_smem(k_connect_edges<3>, sm),
         k_connect_edges<3><<<grid, kThreads, sm, lc.stream>>>(cv, n, agent, qf, qt, out, sm > 0)));
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_collide_pairs(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                 const int32_t *ai, const int32_t *aj, const double *qif,
                                 const double *qit, const double *qjf, const double *qjt,
                                 uint8_t *out) {
    if (n <= 0) return cudaSuccess;
    if (cv.model == 2) return launch_arm_collide_pairs(cv, lc, n, ai, aj, qif, qit, qjf, qjt, out, lc.stats ? lc.stats + 1 : nullptr);
    if (cv.model >= 3) return launch_chain_collide_pairs(cv, lc, n, ai, aj, qif, qit, qjf, qjt, out, lc.stats ? lc.stats + 1 : nullptr);
    const int sm = seg_smem_bytes(cv.col_len);
    const int grid = grid_for(lc, n, sm);
    cudaError_t e;
    MRMP_DISPATCH_D(cv,
                    (e = set_smem(k_collide_pairs<2>, sm),
                     k_collide_pairs<2><<<grid, kThreads, sm, lc.stream>>>(
                         cv, n, ai, aj, qif, qit, qjf, qjt, out, sm > 0)),
                    (e = set_smem(k_collide_pairs<3>, sm),
                     k_collide_pairs<3><<<grid, kThreads, sm, lc.stream>>>(
                         cv, n, ai, aj, qif, qit, qjf, qjt, out, sm > 0)));
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_collide_edges_indexed(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                         const int32_t *ai, const int32_t *aj,
                                         const int32_t *vif, const int32_t *vit,
                                         const int32_t *vjf, const int32_t *vjt,
                                         const double *nodes, int agent0, int n_local,
                                         int nv_stride, uint8_t *out) {
    (void)n_local;
    if (n <= 0) return cudaSuccess;
    if (cv.model == 2) return launch_arm_collide_edges_indexed(cv, lc, n, ai, aj, vif, vit, vjf, vjt, nodes, agent0, nv_stride, out, lc.stats ? lc.stats + 1 : nullptr);
    if (cv.model >= 3) return launch_chain_collide_edges_indexed(cv, lc, n, ai, aj, vif, vit, vjf, vjt, nodes, agent0, nv_stride, out, lc.stats ? lc.stats + 1 : nullptr);
    const int sm = seg_smem_bytes(cv.col_len);
    const int grid = grid_for(lc, n, sm);
    cudaError_t e;
    MRMP_DISPATCH_D(cv,
                    (e = set_smem(k_collide_edges_indexed<2>, sm),
                     k_collide_edges_indexed<2><<<grid, kThreads, sm, lc.stream>>>(
                         cv, n, ai, aj, vif, vit, vjf, vjt, nodes, agent0, nv_stride, out, sm > 0)),
                    (e = set_smem(k_collide_edges_indexed<3>, sm),
                     k_collide_edges_indexed<3><<<grid, kThreads, sm, lc.stream>>>(
                         cv, n, ai, aj, vif, vit, vjf, vjt, nodes, agent0, nv_stride, out, sm > 0)));
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_collide_static_pairs(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                        const int32_t *ai, const int32_t *aj, const double *qi,
                                        const double *qj, uint8_t *out) {
    if (n <= 0) return cudaSuccess;
    if (cv.model == 2) return launch_arm_collide_static_pairs(cv, lc, n, ai, aj, qi, qj, out);
    if (cv.model >= 3) return launch_chain_collide_static_pairs(cv, lc, n, ai, aj, qi, qj, out);
    const int sm = seg_smem_bytes(cv.col_len);
    const int grid = grid_for(lc, n, sm);
    cudaError_t e;
    MRMP_DISPATCH_D(cv,
                    (e = set_smem(k_collide_static_pairs<2>, sm),
                     k_collide_static_pairs<2><<<grid, kThreads, sm, lc.stream>>>(
                         cv, n, ai, aj, qi, qj, out, sm > 0)),
                    (e = set_smem(k_collide_static_pairs<3>, sm),
                     k_collide_static_pairs<3><<<grid, kThreads, sm, lc.stream>>>(
                         cv, n, ai, aj, qi, qj, out, sm > 0)));
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_collide_configs(const CtxView &cv, const LaunchCfg &lc, int64_t n_cfg,
                                   const double *Qf, const double *Qt, uint8_t *out,
                                   int64_t *first_pair) {
    if (n_cfg <= 0) return cudaSuccess;
    if (cv.model == 2) return launch_arm_collide_configs(cv, lc, n_cfg, Qf, Qt, out, first_pair, reinterpret_cast<unsigned long long *>(first_pair), lc.stats ? lc.stats + 1 : nullptr);
    if (cv.model >= 3) return launch_chain_collide_configs(cv, lc, n_cfg, Qf, Qt, out, first_pair, reinterpret_cast<unsigned long long *>(first_pair), lc.stats ? lc.stats + 1 : nullptr);
    const int sm = seg_smem_bytes(cv.col_len);
    const int grid = grid_for(lc, n_cfg * 32, sm);
    cudaError_t e;
    MRMP_DISPATCH_D(cv,
                    (e = set_smem(k_collide_configs<2>, sm),
                     k_collide_configs<2><<<grid, kThreads, sm, lc.stream>>>(
                         cv, n_cfg, Qf, Qt, out, first_pair, sm > 0)),
                    (e = set_smem(k_collide_configs<3>, sm),
                     k_collide_configs<3><<<grid, kThreads, sm, lc.stream>>>(
                         cv, n_cfg, Qf, Qt, out, first_pair, sm > 0)));
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_collide_config_moves(const CtxView &cv, const LaunchCfg &lc, int agent_i,
                                        const double *Q, int64_t n_cand, const double *q_to,
                                        uint8_t *out) {
    if (n_cand <= 0) return cudaSuccess;
    if (cv.model == 2) return launch_arm_collide_config_moves(cv, lc, agent_i, Q, n_cand, q_to, out, lc.stats ? lc.stats + 1 : nullptr);
    if (cv.model >= 3) return launch_chain_collide_config_moves(cv, lc, agent_i, Q, n_cand, q_to, out, lc.stats ? lc.stats + 1 : nullptr);
    const int sm = seg_smem_bytes(cv.col_len);
    const int grid = grid_for(lc, n_cand, sm);
    cudaError_t e;
    MRMP_DISPATCH_D(cv,
                    (e = set_smem(k_collide_config_moves<2>, sm),
                     k_collide_config_moves<2><<<grid, kThreads, sm, lc.stream>>>(
                         cv, agent_i, Q, n_cand, q_to, out, sm > 0)),
                    (e = set_smem(k_collide_config_moves<3>, sm),
                     k_collide_config_moves<3><<<grid, kThreads, sm, lc.stream>>>(
                         cv, agent_i, Q, n_cand, q_to, out, sm > 0)));
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

}  // namespace mrmp
