// kernels_sssp.cu -- the batched SSSP node expansion (src/solvers/sssp.jl:181-223, 235-291).
//
// One search-node expansion of the reference issues, one closure call at a time,
//   expand!  (sssp.jl:235-267): K samples -> steering (:269-291, up to depth+1 conn tests each)
//            -> space-filling test against every roadmap vertex (:250) -> for every accepted
//            sample conn(q_new, v.q) and conn(v.q, q_new) against the whole roadmap (:252-262)
//   successor filter (sssp.jl:196-223): collide(S.Q, p.q, i) for every neighbour p of v and v
// Here the whole expansion is one short kernel sequence on the ctx stream over DEVICE-RESIDENT,
// growing per-agent node tables; the host gets back the accepted vertices, their in/out
// adjacency bitmaps (ascending vertex id = the reference's push! order) and the verdicts of the
// successor filter, and keeps running the branchy best-first search itself.
//
// Sequential semantics kept exact: sample k sees the vertices accepted from samples k' < k of
// the same call (sssp.jl:250 reads the roadmap after earlier push!es) -- resolved by a K x K
// closeness matrix and one serial pass; the new vertex's out-edges exclude itself while its
// in-edge pass includes the self loop (sssp.jl:252-262).
#include <cooperative_groups.h>
#include <math_constants.h>

#include "models.cuh"
#include "point_model.cuh"
#include "sampler.cuh"

namespace mrmp {

extern __shared__ __align__(16) unsigned char smem_raw[];

static inline int seg_smem_bytes(int seg_len) {
    const int bytes = seg_len * 8 + 16;
    return bytes <= kMaxStageBytes ? bytes : 0;
}
template <class K>
static cudaError_t set_smem(K kernel, int bytes) {
    if (bytes > 48 * 1024)
        return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    return cudaSuccess;
}

// conn(q_from, q_to) = (isnothing(epsilon) || dist(q_from,q_to) <= epsilon) && connect(...)
// sssp.jl:87-89
template <class M>
__device__ __forceinline__ bool conn(const CtxView &cv, const ConSeg &cs, const SsspExpand &a,
                                     const double *qf, const double *qt) {
    if (a.eps == a.eps && !M::dist_le(cv, qf, qt, a.eps, a.eps_r2)) return false;
    return M::connect_edge(cv, cs, qf, qt, a.agent);
}

// ---- stage 1: sampling, steering, space-filling test, sequential acceptance ------------------
template <class M>
__global__ void __launch_bounds__(256)
k_sssp_steer(CtxView cv, SsspView view, SsspExpand a, SsspWs ws, int staged) {
    constexpr int D = M::SD;
    __shared__ double s_q[kMaxExpand][D];
    __shared__ int s_near[kMaxExpand];
    __shared__ unsigned long long s_close[kMaxExpand];
    __shared__ int s_rank[kMaxExpand];
    const ConSeg cs(cv, stage_segment(cv, cv.con_off, cv.con_len, smem_raw, staged));
    const int tid = threadIdx.x;
    const int nv0 = view.nv[a.agent];
    double *tab = view.nodes + (int64_t)a.agent * view.cap * D;
    if (tid < a.K) {
        double q_near[D], qh[D];
        load_state<D>(tab, a.v_from, q_near);
        if (a.host_samples) {
#pragma unroll
            for (int c = 0; c < D; ++c) qh[c] = ws.q_rand[tid * D + c];
        } else {
            sample_state<M>(a.seed, a.agent, a.t0 + (uint64_t)tid, qh);
        }
        // steering (sssp.jl:269-291)
        if (!conn<M>(cv, cs, a, q_near, qh)) {
            double ql[D], q[D];
#pragma unroll
            for (int c = 0; c < D; ++c) ql[c] = q_near[c];
            for (int s = 0; s < a.depth; ++s) {
                M::mid(ql, qh, q);
                const bool ok = conn<M>(cv, cs, a, q_near, q);
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    if (ok) ql[c] = q[c];
                    else qh[c] = q[c];
                }
            }
#pragma unroll
            for (int c = 0; c < D; ++c) qh[c] = ql[c];
        }
#pragma unroll
        for (int c = 0; c < D; ++c) s_q[tid][c] = qh[c];
        s_near[tid] = 0;
        s_close[tid] = 0ull;
    }
    __syncthreads();
    // minimum(v -> dist(v.q, q_new), roadmap) > min_dist_thread  (sssp.jl:250)
    for (int idx = tid; idx < a.K * nv0; idx += blockDim.x) {
        const int k = idx / nv0, v = idx - k * nv0;
        double qv[D];
        load_state<D>(tab, v, qv);
        if (!M::dist_gt(cv, qv, s_q[k], a.thr, a.thr_r2)) s_near[k] = 1;
    }
    for (int idx = tid; idx < a.K * a.K; idx += blockDim.x) {
        const int kp = idx / a.K, k = idx - kp * a.K;
        if (kp < k && !M::dist_gt(cv, s_q[kp], s_q[k], a.thr, a.thr_r2))
            atomicOr(&s_close[k], 1ull << kp);
    }
    __syncthreads();
    if (tid == 0) {
        unsigned long long acc = 0ull;
        int n = 0;
        for (int k = 0; k < a.K; ++k) {
            if (!s_near[k] && !(s_close[k] & acc)) {
                acc |= 1ull << k;
                s_rank[k] = n++;
            } else {
                s_rank[k] = -1;
            }
        }
        ws.hdr[0] = n;
        ws.hdr[2] = nv0;
        view.nv[a.agent] = nv0 + n;
    }
    __syncthreads();
    if (tid < a.K && s_rank[tid] >= 0) {
        const int r = s_rank[tid];
#pragma unroll
        for (int c = 0; c < D; ++c) {
            tab[(int64_t)(nv0 + r) * D + c] = s_q[tid][c];
            ws.q_new[r * D + c] = s_q[tid][c];
        }
    }
    for (int idx = tid; idx < 2 * a.K * a.words; idx += blockDim.x) ws.out_bits[idx] = 0u;
}

// ---- stage 2: edges of the accepted vertices ------------------------------------------------
// new vertex r (id nv0 + r): out bit t  <=> conn(q_new, q_t), t <  id   (sssp.jl:252-257)
//                            in  bit t  <=> conn(q_t, q_new), t <= id   (sssp.jl:259-262)
template <class M>
__global__ void __launch_bounds__(128)
k_sssp_conn(CtxView cv, SsspView view, SsspExpand a, SsspWs ws, int staged) {
    constexpr int D = M::SD;
    const ConSeg cs(cv, stage_segment(cv, cv.con_off, cv.con_len, smem_raw, staged));
    const int n_new = ws.hdr[0], nv0 = ws.hdr[2];
    const double *tab = view.nodes + (int64_t)a.agent * view.cap * D;
    const int T = nv0 + a.K;  // targets per new vertex (upper bound)
    const int64_t total = (int64_t)n_new * T * 2;
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < total;
         it += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(it / (2 * T));
        const int rem = (int)(it - (int64_t)r * 2 * T);
        const int t = rem >> 1, dir = rem & 1;
        const int id = nv0 + r;
        if (t > id || (t == id && dir == 0)) continue;
        double qu[D], qt[D];
        load_state<D>(tab, id, qu);
        load_state<D>(tab, t, qt);
        const bool ok = dir == 0 ? conn<M>(cv, cs, a, qu, qt) : conn<M>(cv, cs, a, qt, qu);
        if (ok) {
            uint32_t *bits = (dir == 0 ? ws.out_bits : ws.in_bits) + (int64_t)r * a.words;
            atomicOr(bits + (t >> 5), 1u << (t & 31));
        }
    }
}

// ---- stage 3: successor list of v_from and the operands of the collide filter ---------------
// successors = old neighbours (host order), new neighbours (accepted u with an edge v -> u,
// ascending), then v itself  (`for p_id in vcat(v.neighbors, v.id)`, sssp.jl:197)
template <int D>
__global__ void __launch_bounds__(256)
k_sssp_gather(SsspView view, SsspExpand a, SsspWs ws, int n_agents) {
    __shared__ int s_n;
    const int tid = threadIdx.x;
    const int n_new = ws.hdr[0], nv0 = ws.hdr[2];
    const int cap_s = a.n_old + a.K + 1;
    if (tid == 0) {
        int n = a.n_old;
        for (int r = 0; r < n_new; ++r)
            if ((ws.in_bits[(int64_t)r * a.words + (a.v_from >> 5)] >> (a.v_from & 31)) & 1u)
                ws.succ_ids[n++] = nv0 + r;
        ws.succ_ids[n++] = a.v_from;
        s_n = n;
        ws.hdr[1] = n;
    }
    for (int k = tid; k < a.n_old; k += blockDim.x) ws.succ_ids[k] = ws.old_nbrs[k];
    for (int j = tid; j < n_agents; j += blockDim.x) {
        const double *tj = view.nodes + ((int64_t)j * view.cap + ws.Q_ids[j]) * D;
#pragma unroll
        for (int c = 0; c < D; ++c) ws.Qbuf[j * D + c] = tj[c];
    }
    __syncthreads();
    const int n = s_n;
    const double *tab = view.nodes + (int64_t)a.agent * view.cap * D;
    for (int k = tid; k < cap_s; k += blockDim.x) {
        const int id = k < n ? ws.succ_ids[k] : a.v_from;  // padding: the stay move
        if (k >= n) ws.succ_ids[k] = -1;
#pragma unroll
        for (int c = 0; c < D; ++c) ws.cand[k * D + c] = tab[(int64_t)id * D + c];
    }
    if (a.slow_collide) {
        // operands of collide(S.Q, Q) (sssp.jl:207): Q = S.Q with agent i moved to the successor
        for (int idx = tid; idx < cap_s * n_agents; idx += blockDim.x) {
            const int k = idx / n_agents, j = idx - k * n_agents;
            const int id = k < n ? ws.succ_ids[k] : a.v_from;
            const double *src = j == a.agent
                                    ? tab + (int64_t)(id < 0 ? a.v_from : id) * D
                                    : view.nodes + ((int64_t)j * view.cap + ws.Q_ids[j]) * D;
            const double *from = view.nodes + ((int64_t)j * view.cap + ws.Q_ids[j]) * D;
#pragma unroll
            for (int c = 0; c < D; ++c) {
                ws.cfg_to[(int64_t)idx * D + c] = src[c];
                ws.cfg_from[(int64_t)idx * D + c] = from[c];
            }
        }
    }
}

// ---- single-launch form for search-sized roadmaps (point models) ------------------------------
// A popped search node of the reference's configs touches a roadmap of 10^1..10^3 vertices:
// the whole expansion is ~10^4 primitive tests, i.e. a few microseconds of GPU work, so the
// call is bound by launch and copy latency, not throughput.  This kernel runs every stage in
// ONE launch of one thread-block CLUSTER of 8 CTAs x 1024 threads: the leader CTA keeps the
// state of the expansion in its shared memory and runs the serial stages; the 2*K'*nv edge
// tests -- the bulk of the work -- are spread over all 8 CTAs, which read the accepted
// vertices from, and OR their verdict bits into, the leader's shared memory through
// distributed shared memory (cluster.map_shared_rank).  Inputs are read from and results
// written to MAPPED PINNED host memory directly (no memcpy nodes around the kernel), so a call
// is one launch + one stream synchronisation.  Steering runs one sample per warp with the
// obstacles of each conn test spread over the lanes.  The multi-kernel path above takes
// over for large roadmaps and for the discretised models.

// conn(q_from, q_to) of a point model with the obstacle loop spread over the lanes of a warp
// (point2d.jl:59-66 / point3d.jl:66-77 + the epsilon gate of sssp.jl:87-89); warp-uniform result
template <class M>
__device__ __forceinline__ bool warp_conn_point(const CtxView &cv, const ConSeg &cs,
                                                const SsspExpand &a, const double *qf,
                                                const double *qt, int lane) {
    constexpr int D = M::SD;
    if (a.eps == a.eps && !M::dist_le(cv, qf, qt, a.eps, a.eps_r2)) return false;
    const double r = cs.rad(a.agent);
    bool ok = !out_of_field<D>(qt, r);
    double u[D];
#pragma unroll
    for (int k = 0; k < D; ++k) u[k] = dsub(qf[k], qt[k]);
    const double uu = dotd<D>(u, u);
    for (int m = lane; m < cs.n_obs; m += 32) {
        double c[D];
#pragma unroll
        for (int k = 0; k < D; ++k) c[k] = cs.oc(k, m);
        ok &= !segpt_lt<D, true>(qf, qt, u, uu, c, dadd(cs.oc(D, m), r), cs.t2_im(a.agent, m));
    }
    return __all_sync(0xffffffffu, ok);
}
constexpr int kFusedThreads = 1024;
constexpr int kFusedCluster = 8;  // CTAs of the thread-block cluster that share the edge tests
namespace cg = cooperative_groups;

template <class M>
__global__ void __launch_bounds__(kFusedThreads, 1)
k_sssp_expand_fused(CtxView cv, SsspView view, SsspExpand a, const __grid_constant__ SsspFusedIn in,
                    int32_t *__restrict__ h_hdr,
                    double *__restrict__ h_qnew, uint32_t *__restrict__ h_bits,
                    int32_t *__restrict__ h_succ, uint8_t *__restrict__ h_verdict) {
    constexpr int D = M::SD;
    __shared__ double s_q[kMaxExpand][D];
    __shared__ int s_near[kMaxExpand];
    __shared__ unsigned long long s_close[kMaxExpand];
    __shared__ int s_rank[kMaxExpand];
    __shared__ int s_nnew, s_nsucc, s_nv0;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned crank = cluster.block_rank(), csize = cluster.num_blocks();
    const bool leader = crank == 0;
    const int tid = threadIdx.x, N = cv.n_agents;
    const int cap_s = a.n_old + a.K + 1;
    // dynamic smem: Qids[N] | succ[cap_s] | flag[cap_s] | bits[2*K*words] | Q[N][D]
    int32_t *s_qids = reinterpret_cast<int32_t *>(smem_raw);
    int32_t *s_succ = s_qids + N;
    int32_t *s_flag = s_succ + cap_s;
    uint32_t *s_bits = reinterpret_cast<uint32_t *>(s_flag + cap_s);
    double *s_Q = reinterpret_cast<double *>(smem_raw + ((4 * (size_t)(N + 2 * cap_s + 2 * a.K * a.words) + 15) & ~(size_t)15));
    const ConSeg cs(cv, cv.blob + cv.con_off);
    const ColSeg ks(cv, cv.blob + cv.col_off);
    double *tab = view.nodes + (int64_t)a.agent * view.cap * D;
    const int lane = tid & 31, warp = tid >> 5;
    int nv0 = 0, n_new = 0;
    if (leader) {
        nv0 = view.nv[a.agent];
        // ---- inputs from the kernel parameters: Q_ids[N], old_nbrs[n_old]
        for (int t = tid; t < N; t += blockDim.x) s_qids[t] = in.q_ids[t];
        for (int t = tid; t < a.n_old; t += blockDim.x) s_succ[t] = in.old_nbrs[t];
        for (int t = tid; t < cap_s; t += blockDim.x) s_flag[t] = 0;
        for (int t = tid; t < 2 * a.K * a.words; t += blockDim.x) s_bits[t] = 0u;
        // ---- steering (sssp.jl:269-291): sample k on warp k, obstacles over the lanes
        if (warp < a.K) {
            const int k = warp;
            double q_near[D], qh[D];
            load_state<D>(tab, a.v_from, q_near);
            if (a.host_samples) {
#pragma unroll
                for (int c = 0; c < D; ++c) qh[c] = in.q_rand[k * D + c];
            } else {
                sample_state<M>(a.seed, a.agent, a.t0 + (uint64_t)k, qh);
            }
            if (!warp_conn_point<M>(cv, cs, a, q_near, qh, lane)) {
                double ql[D], q[D];
#pragma unroll
                for (int c = 0; c < D; ++c) ql[c] = q_near[c];
                for (int st = 0; st < a.depth; ++st) {
                    M::mid(ql, qh, q);
                    const bool ok = warp_conn_point<M>(cv, cs, a, q_near, q, lane);
#pragma unroll
                    for (int c = 0; c < D; ++c) {
                        if (ok) ql[c] = q[c];
                        else qh[c] = q[c];
                    }
                }
#pragma unroll
                for (int c = 0; c < D; ++c) qh[c] = ql[c];
            }
            if (lane == 0) {
#pragma unroll
                for (int c = 0; c < D; ++c) s_q[k][c] = qh[c];
                s_near[k] = 0;
                s_close[k] = 0ull;
            }
        }
        __syncthreads();
        for (int j = tid; j < N; j += blockDim.x) {
            const double *tj = view.nodes + ((int64_t)j * view.cap + s_qids[j]) * D;
#pragma unroll
            for (int c = 0; c < D; ++c) s_Q[j * D + c] = tj[c];
        }
        // ---- space-filling test (sssp.jl:250) and the K x K closeness matrix
        for (int idx = tid; idx < a.K * nv0; idx += blockDim.x) {
            const int k = idx / nv0, v = idx - k * nv0;
            double qv[D];
            load_state<D>(tab, v, qv);
            if (!M::dist_gt(cv, qv, s_q[k], a.thr, a.thr_r2)) s_near[k] = 1;
        }
        for (int idx = tid; idx < a.K * a.K; idx += blockDim.x) {
            const int kp = idx / a.K, k = idx - kp * a.K;
            if (kp < k && !M::dist_gt(cv, s_q[kp], s_q[k], a.thr, a.thr_r2))
                atomicOr(&s_close[k], 1ull << kp);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned long long acc = 0ull;
            int n = 0;
            for (int k = 0; k < a.K; ++k) {
                if (!s_near[k] && !(s_close[k] & acc)) {
                    acc |= 1ull << k;
                    s_rank[k] = n++;
                } else {
                    s_rank[k] = -1;
                }
            }
            s_nnew = n;
            s_nv0 = nv0;
            view.nv[a.agent] = nv0 + n;
        }
        __syncthreads();
        // accepted samples: to the node table, the result block, and compacted to the front of
        // s_q in acceptance order for the edge tests
        double mine[D];
        int my_rank = -1;
        if (tid < a.K) {
            my_rank = s_rank[tid];
#pragma unroll
            for (int c = 0; c < D; ++c) mine[c] = s_q[tid][c];
        }
        __syncthreads();
        if (my_rank >= 0) {
#pragma unroll
            for (int c = 0; c < D; ++c) {
                s_q[my_rank][c] = mine[c];
                tab[(int64_t)(nv0 + my_rank) * D + c] = mine[c];
                h_qnew[my_rank * D + c] = mine[c];
            }
        }
    }
    cluster.sync();  // the leader's s_q / s_nnew / s_nv0 are final

    // ---- edges of the accepted vertices (sssp.jl:252-262), spread over the cluster: every CTA
    // reads the accepted vertices from the leader's shared memory and ORs into its bitmaps
    {
        const double(*l_q)[D] = reinterpret_cast<const double(*)[D]>(cluster.map_shared_rank(&s_q[0][0], 0));
        uint32_t *l_bits = cluster.map_shared_rank(s_bits, 0);
        n_new = *cluster.map_shared_rank(&s_nnew, 0);
        nv0 = *cluster.map_shared_rank(&s_nv0, 0);
        const int T = nv0 + a.K;
        const int total = n_new * T * 2;
        for (int it = crank * blockDim.x + tid; it < total; it += csize * blockDim.x) {
            const int r = it / (2 * T);
            const int rem = it - r * 2 * T;
            const int t = rem >> 1, dir = rem & 1;
            const int id = nv0 + r;
            if (t > id || (t == id && dir == 0)) continue;
            double qu[D], qt[D];
#pragma unroll
            for (int c = 0; c < D; ++c) qu[c] = l_q[r][c];
            if (t >= nv0) {
#pragma unroll
                for (int c = 0; c < D; ++c) qt[c] = l_q[t - nv0][c];
            } else {
                load_state<D>(tab, t, qt);  // vertices that existed before this launch
            }
            const bool ok = dir == 0 ? conn<M>(cv, cs, a, qu, qt) : conn<M>(cv, cs, a, qt, qu);
            if (ok)
                atomicOr(l_bits + (dir == 0 ? 0 : a.K * a.words) + r * a.words + (t >> 5),
                         1u << (t & 31));
        }
    }
    cluster.sync();  // all verdict bits have landed in the leader's shared memory
    if (!leader) return;

    // ---- successor list
    if (tid == 0) {
        int n = a.n_old;
        const uint32_t *in_bits = s_bits + a.K * a.words;
        for (int r = 0; r < n_new; ++r)
            if ((in_bits[r * a.words + (a.v_from >> 5)] >> (a.v_from & 31)) & 1u) s_succ[n++] = nv0 + r;
        s_succ[n++] = a.v_from;
        s_nsucc = n;
    }
    __syncthreads();
    const int n_succ = s_nsucc;
    // ---- successor filter
    double qf[D];
#pragma unroll
    for (int c = 0; c < D; ++c) qf[c] = s_Q[a.agent * D + c];
    if (!a.slow_collide) {  // collide(S.Q, p.q, i): point.jl:27-33
        for (int it = tid; it < n_succ * N; it += blockDim.x) {
            const int k = it / N, j = it - k * N;
            if (j == a.agent) continue;
            const int id = s_succ[k];
            double qt[D], u[D];
            if (id >= nv0) {
#pragma unroll
                for (int c = 0; c < D; ++c) qt[c] = s_q[id - nv0][c];
            } else {
                load_state<D>(tab, id, qt);
            }
#pragma unroll
            for (int c = 0; c < D; ++c) u[c] = dsub(qf[c], qt[c]);
            if (point_collide_move_vs_static<D>(ks, qf, qt, u, dotd<D>(u, u), s_Q + j * D, a.agent, j))
                s_flag[k] = 1;
        }
    } else {  // collide(S.Q, Q): all pairs i < j of the 6-arity form (point.jl:18-25)
        const int P = N * (N - 1) / 2;
        for (int it = tid; it < n_succ * P; it += blockDim.x) {
            const int k = it / P;
            int rr = it - k * P, i = 0;
            while (rr >= N - 1 - i) {
                rr -= N - 1 - i;
                ++i;
            }
            const int j = i + 1 + rr;
            const int id = s_succ[k];
            double qt[D];
            if (id >= nv0) {
#pragma unroll
                for (int c = 0; c < D; ++c) qt[c] = s_q[id - nv0][c];
            } else {
                load_state<D>(tab, id, qt);
            }
            const double *fi = s_Q + i * D, *fj = s_Q + j * D;
            const double *ti = i == a.agent ? qt : fi, *tj = j == a.agent ? qt : fj;
            if (point_collide_pair<D>(ks, fi, ti, fj, tj, i, j)) s_flag[k] = 1;
        }
    }
    __syncthreads();
    // ---- results to mapped host memory
    for (int t = tid; t < 2 * a.K * a.words; t += blockDim.x) h_bits[t] = s_bits[t];
    for (int t = tid; t < n_succ; t += blockDim.x) {
        h_succ[t] = s_succ[t];
        h_verdict[t] = (uint8_t)s_flag[t];
    }
    if (tid == 0) {
        h_hdr[0] = n_new;
        h_hdr[1] = n_succ;
        h_hdr[2] = nv0;
    }
    // completion flag for the polling host: after every thread's results are visible system-wide
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        *reinterpret_cast<volatile int32_t *>(h_hdr + 3) = a.seq;
        __threadfence_system();
    }
}

// ---- small per-state helpers for the host search loops --------------------------------------
// op 0: out[k] = dist(q1[k], q2[k]);  op 1: out[k][:] = get_mid_status(q1[k], q2[k])
template <class M>
__global__ void __launch_bounds__(256)
k_state_op(CtxView cv, int op, int64_t n, const double *__restrict__ q1,
           const double *__restrict__ q2, double *__restrict__ out) {
    constexpr int D = M::SD;
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
         k += (int64_t)gridDim.x * blockDim.x) {
        double a[D], b[D];
#pragma unroll
        for (int c = 0; c < D; ++c) {
            a[c] = q1[k * D + c];
            b[c] = q2[k * D + c];
        }
        if (op == 0) {
            out[k] = M::dist(cv, a, b);
        } else {
            double m[D];
            M::mid(a, b, m);
#pragma unroll
            for (int c = 0; c < D; ++c) out[k * D + c] = m[c];
        }
    }
}

// argmin_v dist(V[v], q) (reversed: dist(q, V[v])); first minimum like Julia's argmin
// (sssp.jl:416-417).  One CTA; NaN distances are never selected unless all are NaN (index 0).
template <class M>
__global__ void __launch_bounds__(256)
k_nearest(CtxView cv, int n_v, const double *__restrict__ V, const double *__restrict__ q,
          int reversed, int32_t *__restrict__ idx_out, double *__restrict__ dist_out) {
    constexpr int D = M::SD;
    __shared__ double s_d[256];
    __shared__ int s_i[256];
    double qq[D];
#pragma unroll
    for (int c = 0; c < D; ++c) qq[c] = q[c];
    double best = CUDART_INF;
    int bi = 0x7fffffff;
    for (int v = threadIdx.x; v < n_v; v += blockDim.x) {
        double qv[D];
#pragma unroll
        for (int c = 0; c < D; ++c) qv[c] = V[(int64_t)v * D + c];
        const double dd = reversed ? M::dist(cv, qq, qv) : M::dist(cv, qv, qq);
        if (dd < best || (dd == best && v < bi)) {
            best = dd;
            bi = v;
        }
    }
    s_d[threadIdx.x] = best;
    s_i[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double d2 = s_d[threadIdx.x + o];
            const int i2 = s_i[threadIdx.x + o];
            if (d2 < s_d[threadIdx.x] || (d2 == s_d[threadIdx.x] && i2 < s_i[threadIdx.x])) {
                s_d[threadIdx.x] = d2;
                s_i[threadIdx.x] = i2;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const bool none = s_i[0] == 0x7fffffff;  // every distance NaN (or inf)
        *idx_out = none ? 0 : s_i[0];
        if (dist_out) {
            double q0[D];
#pragma unroll
            for (int c = 0; c < D; ++c) q0[c] = V[c];
            *dist_out = none ? (reversed ? M::dist(cv, qq, q0) : M::dist(cv, q0, qq)) : s_d[0];
        }
    }
}

// ---- distance tables (solvers/utils.jl:146-179), one CTA per roadmap -------------------------
// The reference runs Dijkstra from the goal vertex along the neighbour lists with
// g = dist(v.q, u.q) + table[v].  Rounded addition of a non-negative weight is monotone, so the
// table is the minimum over all paths of the left-to-right accumulated rounded sum -- a value
// that does not depend on the order in which edges are relaxed.  Parallel relaxation of all
// edges to a fixed point (Bellman-Ford with atomicMin on the bit patterns of non-negative
// doubles) therefore yields bit-identical tables.
template <class M>
__global__ void __launch_bounds__(512)
k_distance_tables(CtxView cv, const int32_t *__restrict__ v_off, const int32_t *__restrict__ row_ptr,
                  const int32_t *__restrict__ col_idx, const double *__restrict__ nodes,
                  const int32_t *__restrict__ goal, double stay_penalty, double *__restrict__ w,
                  double *__restrict__ table) {
    constexpr int D = M::SD;
    __shared__ int s_changed;
    const int g = blockIdx.x;
    const int v0 = v_off[g], nv = v_off[g + 1] - v0;
    const int64_t e0 = row_ptr[v0], e1 = row_ptr[v0 + nv];
    const double *tab = nodes + (int64_t)v0 * D;
    unsigned long long *T = reinterpret_cast<unsigned long long *>(table + v0);
    // edge lengths dist(v.q, u.q), one warp per row
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (int v = warp; v < nv; v += n_warps) {
        double qv[D];
        load_state<D>(tab, v, qv);
        for (int64_t e = row_ptr[v0 + v] + lane; e < row_ptr[v0 + v + 1]; e += 32) {
            double qu[D];
            load_state<D>(tab, col_idx[e], qu);
            // gen_g_func (solvers/utils.jl:34-36): dist + (q_from == q_to ? stay_penalty : 0);
            // `==` of two immutable states is egality: every field bit-identical
            bool same = true;
#pragma unroll
            for (int x = 0; x < D; ++x)
                same &= __double_as_longlong(qv[x]) == __double_as_longlong(qu[x]);
            w[e] = dadd(M::dist(cv, qv, qu), same ? stay_penalty : 0.0);
        }
    }
    for (int v = threadIdx.x; v < nv; v += blockDim.x)
        T[v] = v == goal[g] ? 0ull : 0x7ff0000000000000ull;  // typemax(Float64) == Inf
    __syncthreads();
    // non-negative edge costs: the fixed point is reached after at most nv rounds; the cap keeps a
    // corrupted input (negative cost on a cycle) from spinning for ever
    for (int round = 0; round <= nv; ++round) {
        if (threadIdx.x == 0) s_changed = 0;
        __syncthreads();
        bool changed = false;
        for (int v = warp; v < nv; v += n_warps) {
            const double d = __longlong_as_double((long long)T[v]);
            if (d == CUDART_INF) continue;
            for (int64_t e = row_ptr[v0 + v] + lane; e < row_ptr[v0 + v + 1]; e += 32) {
                const double cand = dadd(w[e], d);  // g_func(v.q, u.q) + d   (:168)
                if (cand == cand) {                  // NaN never wins `g < table[u]`
                    const unsigned long long bits = (unsigned long long)__double_as_longlong(cand);
                    if (bits < T[col_idx[e]]) {
                        atomicMin(&T[col_idx[e]], bits);
                        changed = true;
                    }
                }
            }
        }
        if (changed) s_changed = 1;
        __syncthreads();
        if (!s_changed) break;
        __syncthreads();
    }
    (void)e0;
    (void)e1;
}

// ---- extend! of RRT-Connect (sssp.jl:404-440) in one launch -------------------------------------
// nearest vertex of the tree (first minimum, like Julia's argmin), the connection test towards the
// sample and the bisection of `steering_depth` rounds: ~3 + 3*depth closure calls of the reference,
// each of which used to be its own launch + wait.  One CTA; the sequential decisions run on
// thread 0 (the reference's own order of calls).  flag: 0 trapped, 1 advanced, 2 reached.
template <class M>
__global__ void __launch_bounds__(256)
k_rrt_extend(CtxView cv, SsspExpand a, int n_v, const double *__restrict__ V,
             const double *__restrict__ q_rand, int from_start, int32_t *__restrict__ out_flag,
             int32_t *__restrict__ out_near, double *__restrict__ out_q) {
    constexpr int D = M::SD;
    __shared__ double s_d[256];
    __shared__ int s_i[256];
    double qr[D];
#pragma unroll
    for (int c = 0; c < D; ++c) qr[c] = q_rand[c];
    double best = CUDART_INF;
    int bi = 0x7fffffff;
    for (int v = threadIdx.x; v < n_v; v += blockDim.x) {
        double qv[D];
#pragma unroll
        for (int c = 0; c < D; ++c) qv[c] = V[(int64_t)v * D + c];
        // argmin(q -> dist(q, q_rand), V) from the start tree, dist(q_rand, q) from the goal tree
        const double dd = from_start ? M::dist(cv, qv, qr) : M::dist(cv, qr, qv);
        if (dd < best || (dd == best && v < bi)) {
            best = dd;
            bi = v;
        }
    }
    s_d[threadIdx.x] = best;
    s_i[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double d2 = s_d[threadIdx.x + o];
            const int i2 = s_i[threadIdx.x + o];
            if (d2 < s_d[threadIdx.x] || (d2 == s_d[threadIdx.x] && i2 < s_i[threadIdx.x])) {
                s_d[threadIdx.x] = d2;
                s_i[threadIdx.x] = i2;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    const int near = s_i[0] == 0x7fffffff ? 0 : s_i[0];  // all distances NaN: index 0
    const ConSeg cs(cv, cv.blob + cv.con_off);
    double q_near[D], q_l[D], q_h[D], q[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
        q_near[c] = V[(int64_t)near * D + c];
        q_l[c] = q_near[c];
        q_h[c] = qr[c];
    }
    // from_start: connect(q_near, q);  else: connect(q) && connect(q, q_near)   (:424, :428)
    const auto linked = [&](const double *x) {
        if (from_start) return conn<M>(cv, cs, a, q_near, x);
        return M::connect_point(cv, cs, x, a.agent) && conn<M>(cv, cs, a, x, q_near);
    };
    int flag = 2;
    if (!linked(q_h)) {
        flag = 0;
        for (int st = 0; st < a.depth; ++st) {
            if (from_start) M::mid(q_l, q_h, q);
            else M::mid(q_h, q_l, q);
            const bool ok = linked(q);
#pragma unroll
            for (int c = 0; c < D; ++c) {
                if (ok) q_l[c] = q[c];
                else q_h[c] = q[c];
            }
            if (ok) flag = 1;
        }
#pragma unroll
        for (int c = 0; c < D; ++c) q_h[c] = q_l[c];
    }
    *out_flag = flag;
    *out_near = near;
#pragma unroll
    for (int c = 0; c < D; ++c) out_q[c] = q_h[c];
}

// ---- launchers ---------------------------------------------------------------------------------
#define MRMP_MODEL_SWITCH(cv, EXPR)                      \
    switch ((cv).model) {                                \
    case 0: { using M = PointModel<2>; EXPR; } break;    \
    case 1: { using M = PointModel<3>; EXPR; } break;    \
    case 2: { using M = ArmModel; EXPR; } break;         \
    case 3: { using M = ChainModel<3>; EXPR; } break;    \
    case 4: { using M = ChainModel<4>; EXPR; } break;    \
    case 5: { using M = ChainModel<5>; EXPR; } break;    \
    case 6: { using M = ChainModel<6>; EXPR; } break;    \
    default: return cudaErrorNotSupported;               \
    }

cudaError_t launch_sssp_expand(const CtxView &cv, const LaunchCfg &lc, const SsspView &view,
                               const SsspExpand &a, const SsspWs &ws, int nv_before) {
    if (a.K < 1 || a.K > kMaxExpand) return cudaErrorInvalidValue;
    const int sm = seg_smem_bytes(cv.con_len);
    cudaError_t e = cudaSuccess;
    MRMP_MODEL_SWITCH(cv, (e = set_smem(k_sssp_steer<M>, sm),
                           k_sssp_steer<M><<<1, 256, sm, lc.stream>>>(cv, view, a, ws, sm > 0)));
    if (e != cudaSuccess) return e;
    count_launch();
    // conn tests: at most K * 2 * (nv_before + K) items, one per thread
    const int64_t items = (int64_t)a.K * 2 * (nv_before + a.K);
    int64_t want = (items + 127) / 128;
    const int64_t cap = (int64_t)lc.sm_count * 8;
    const int grid = (int)(want < 1 ? 1 : (want < cap ? want : cap));
    MRMP_MODEL_SWITCH(cv, (e = set_smem(k_sssp_conn<M>, sm),
                           k_sssp_conn<M><<<grid, 128, sm, lc.stream>>>(cv, view, a, ws, sm > 0)));
    if (e != cudaSuccess) return e;
    count_launch();
    if (cv.d == 6)
        k_sssp_gather<6><<<1, 256, 0, lc.stream>>>(view, a, ws, cv.n_agents);
    else if (cv.d == 3)
        k_sssp_gather<3><<<1, 256, 0, lc.stream>>>(view, a, ws, cv.n_agents);
    else
        k_sssp_gather<2><<<1, 256, 0, lc.stream>>>(view, a, ws, cv.n_agents);
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int cap_s = a.n_old + a.K + 1;
    if (a.slow_collide)
        return launch_collide_configs(cv, lc, cap_s, ws.cfg_from, ws.cfg_to, ws.succ_out,
                                      ws.cfg_first);
    return launch_collide_config_moves(cv, lc, a.agent, ws.Qbuf, cap_s, ws.cand, ws.succ_out);
}

size_t sssp_fused_smem(const CtxView &cv, const SsspExpand &a) {
    const int cap_s = a.n_old + a.K + 1;
    const size_t ints = 4 * (size_t)(cv.n_agents + 2 * cap_s + 2 * a.K * a.words);
    return ((ints + 15) & ~(size_t)15) + sizeof(double) * cv.d * (size_t)cv.n_agents;
}

// one launch; h_* are device-visible addresses of mapped pinned host memory
cudaError_t launch_sssp_expand_fused(const CtxView &cv, const LaunchCfg &lc, const SsspView &view,
                                     const SsspExpand &a, const SsspFusedIn &in, int32_t *h_hdr,
                                     double *h_qnew, uint32_t *h_bits, int32_t *h_succ,
                                     uint8_t *h_verdict) {
    if (a.K < 1 || a.K > kFusedMaxK || cv.model >= 2 || cv.n_agents > kFusedMaxAgents ||
        a.n_old > kFusedMaxOld)
        return cudaErrorNotSupported;
    const int sm = (int)sssp_fused_smem(cv, a);
    if (sm > 160 * 1024) return cudaErrorNotSupported;
    // one thread-block cluster of kFusedCluster CTAs
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kFusedCluster, 1, 1);
    cfg.blockDim = dim3(kFusedThreads, 1, 1);
    cfg.dynamicSmemBytes = (size_t)sm;
    cfg.stream = lc.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kFusedCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaSuccess;
    if (cv.model == 0) {
        using M = PointModel<2>;
        e = set_smem(k_sssp_expand_fused<M>, sm);
        if (e == cudaSuccess)
            e = cudaLaunchKernelEx(&cfg, k_sssp_expand_fused<M>, cv, view, a, in, h_hdr, h_qnew, h_bits,
                                   h_succ, h_verdict);
    } else {
        using M = PointModel<3>;
        e = set_smem(k_sssp_expand_fused<M>, sm);
        if (e == cudaSuccess)
            e = cudaLaunchKernelEx(&cfg, k_sssp_expand_fused<M>, cv, view, a, in, h_hdr, h_qnew, h_bits,
                                   h_succ, h_verdict);
    }
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_distance_tables(const CtxView &cv, const LaunchCfg &lc, int n_graphs,
                                   const int32_t *v_off, const int32_t *row_ptr,
                                   const int32_t *col_idx, const double *nodes,
                                   const int32_t *goal, double stay_penalty, double *w,
                                   double *table) {
    if (n_graphs <= 0) return cudaSuccess;
    MRMP_MODEL_SWITCH(cv, (k_distance_tables<M><<<n_graphs, 512, 0, lc.stream>>>(
                              cv, v_off, row_ptr, col_idx, nodes, goal, stay_penalty, w, table)));
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_state_op(const CtxView &cv, const LaunchCfg &lc, int op, int64_t n,
                            const double *q1, const double *q2, double *out) {
    if (n <= 0) return cudaSuccess;
    int64_t want = (n + 255) / 256;
    const int64_t cap = (int64_t)lc.sm_count * 8;
    const int grid = (int)(want < cap ? want : cap);
    MRMP_MODEL_SWITCH(cv, (k_state_op<M><<<grid, 256, 0, lc.stream>>>(cv, op, n, q1, q2, out)));
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_rrt_extend(const CtxView &cv, const LaunchCfg &lc, const SsspExpand &a, int n_v,
                              const double *V, const double *q_rand, int from_start,
                              int32_t *out_flag, int32_t *out_near, double *out_q) {
    if (n_v <= 0) return cudaErrorInvalidValue;
    MRMP_MODEL_SWITCH(cv, (k_rrt_extend<M><<<1, 256, 0, lc.stream>>>(cv, a, n_v, V, q_rand, from_start,
                                                                    out_flag, out_near, out_q)));
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_nearest(const CtxView &cv, const LaunchCfg &lc, int n_v, const double *V,
                           const double *q, int reversed, int32_t *idx_out, double *dist_out) {
    if (n_v <= 0) return cudaErrorInvalidValue;
    MRMP_MODEL_SWITCH(cv, (k_nearest<M><<<1, 256, 0, lc.stream>>>(cv, n_v, V, q, reversed, idx_out,
                                                                 dist_out)));
    count_launch();
    return cudaGetLastError();
}

}  // namespace mrmp
