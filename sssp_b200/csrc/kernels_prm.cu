// kernels_prm.cu -- roadmap construction: free-space sampling with ordered compaction,
// all-pairs radius gate + edge connect -> adjacency bitmap -> CSR.
//
// Reference: PRM! (src/solvers/prm.jl:178-213): rejection sampling (:193-199), then for all
// ordered pairs (j,k), j != k, in row-major order:
//   (rad === nothing || dist(q_j,q_k) <= rad) && connect(q_j,q_k)  =>  push!(neighbors_j, k)
// Rows come out ascending in k because the bitmap is expanded in bit order.
#include "models.cuh"
#include "sampler.cuh"

namespace mrmp {

extern __shared__ __align__(16) unsigned char smem_raw[];

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

template <class M>
__global__ void __launch_bounds__(kThreads)
k_sample_states(int agent, uint64_t seed, uint64_t t0, int64_t n, double *__restrict__ q_out) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
         k += (int64_t)gridDim.x * blockDim.x) {
        double q[M::SD];
        sample_state<M>(seed, agent, t0 + (uint64_t)k, q);
#pragma unroll
        for (int c = 0; c < M::SD; ++c) q_out[k * M::SD + c] = q[c];
    }
}

// One CTA per agent.  Round r tests tries [r*T, (r+1)*T); accepted tries are compacted in try
// order (ballot + warp prefix), so the result equals the sequential loop of prm.jl:193-199.
template <class M>
__global__ void __launch_bounds__(kThreads)
k_sample_free(CtxView cv, int agent0, int n_fixed, int nv, int nv_stride, uint64_t seed,
              double *__restrict__ nodes, int64_t *__restrict__ tries_out, int64_t max_tries,
              int staged) {
    __shared__ int warp_cnt[kWarps];
    __shared__ int accepted_sh;
    const ConSeg cs(cv, stage_segment(cv, cv.con_off, cv.con_len, smem_raw, staged));
    const int a = blockIdx.x;
    const int agent = agent0 + a;
    const int wanted = nv - n_fixed;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *out = nodes + ((int64_t)a * nv_stride + n_fixed) * M::SD;
    int accepted = 0;
    int64_t tries = wanted <= 0 ? 0 : -1;
    for (int64_t t0 = 0; accepted < wanted && t0 < max_tries; t0 += kThreads) {
        const int64_t t = t0 + threadIdx.x;
        double q[M::SD];
        sample_state<M>(seed, agent, (uint64_t)t, q);
        const bool ok = t < max_tries && M::connect_point(cv, cs, q, agent);
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int before = accepted;
        for (int w = 0; w < warp; ++w) before += warp_cnt[w];
        const int pos = before + __popc(bal & ((1u << lane) - 1u));
        if (ok && pos < wanted) {
#pragma unroll
            for (int c = 0; c < M::SD; ++c) out[(int64_t)pos * M::SD + c] = q[c];
            if (pos == wanted - 1 && tries_out) tries_out[a] = t + 1;
        }
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < kWarps; ++w) tot += warp_cnt[w];
            accepted_sh = accepted + tot;
        }
        __syncthreads();
        accepted = accepted_sh;
        __syncthreads();
    }
    if (threadIdx.x == 0 && tries_out && (accepted < wanted || wanted <= 0)) tries_out[a] = tries;
}

// ---- adjacency bitmap: one warp per row (agent, j) ------------------------------------------
// Phase A: radius gate for every k (ballot -> one 32-bit word per 32 vertices).
// Phase B: the sparse survivors are compacted into a per-warp queue so that the expensive
//          edge-vs-obstacle test runs on dense lanes; failures clear their bit.
template <class M>
__global__ void __launch_bounds__(kThreads)
k_prm_adjacency(CtxView cv, int agent0, int n_agents, int nv, const double *__restrict__ rad,
                const double *__restrict__ rad_r2, const double *__restrict__ nodes,
                uint32_t *__restrict__ bitmap, int words, int32_t *__restrict__ row_cnt,
                int staged, int seg_bytes) {
    const ConSeg cs(cv, stage_segment(cv, cv.con_off, cv.con_len, smem_raw, staged));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // per-warp scratch behind the staged segment: words bitmap + 64-entry queue
    uint32_t *wbits = reinterpret_cast<uint32_t *>(smem_raw + seg_bytes) + warp * (words + 64);
    int *queue = reinterpret_cast<int *>(wbits + words);
    const int64_t n_rows = (int64_t)n_agents * nv;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n_rows;
         row += n_warps) {
        const int a = (int)(row / nv), j = (int)(row - (int64_t)a * nv);
        const int agent = agent0 + a;
        const double *tab = nodes + (int64_t)a * nv * M::SD;
        const double r = rad[a], r2 = rad_r2[a];
        const bool gate_all = isnan(r);
        double qj[M::SD];
        load_state<M::SD>(tab, j, qj);
        int qlen = 0;  // warp-uniform queue length
        int cnt = 0;
        for (int t = 0; t < words; ++t) {
            const int k = 32 * t + lane;
            bool pass = false;
            double qk[M::SD];
            if (k < nv && k != j) {
                load_state<M::SD>(tab, k, qk);
                pass = gate_all || M::dist_le(cv, qj, qk, r, r2);  // prm.jl:207
            }
            const unsigned bal = __ballot_sync(0xffffffffu, pass);
            if (lane == 0) wbits[t] = bal;
            if (pass) queue[(cnt + __popc(bal & ((1u << lane) - 1u))) & 63] = k;
            qlen += __popc(bal);
            cnt += __popc(bal);
            __syncwarp();
            if (qlen >= 32) {
                // dense flush of 32 queued candidates (oldest first)
                const int head = (cnt - qlen) & 63;  // ring position of the oldest entry
                const int k2 = queue[(head + lane) & 63];
                load_state<M::SD>(tab, k2, qk);
                const bool ok = M::connect_edge(cv, cs, qj, qk, agent);
                if (!ok) atomicAnd(&wbits[k2 >> 5], ~(1u << (k2 & 31)));
                qlen -= 32;
                __syncwarp();
            }
        }
        if (qlen > 0) {
            const int head = (cnt - qlen) & 63;
            if (lane < qlen) {
                const int k2 = queue[(head + lane) & 63];
                double qk[M::SD];
                load_state<M::SD>(tab, k2, qk);
                const bool ok = M::connect_edge(cv, cs, qj, qk, agent);
                if (!ok) atomicAnd(&wbits[k2 >> 5], ~(1u << (k2 & 31)));
            }
        }
        __syncwarp();
        int total = 0;
        for (int t = lane; t < words; t += 32) {
            const uint32_t w = wbits[t];
            bitmap[row * words + t] = w;
            total += __popc(w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
        if (lane == 0) row_cnt[row] = total;
        __syncwarp();
    }
}

// ---- exclusive scan of the row counts (single CTA; n_rows is at most a few 10^6) -----------
__global__ void __launch_bounds__(1024)
k_row_scan(int64_t n_rows, const int32_t *__restrict__ row_cnt, int32_t *__restrict__ row_ptr,
           int64_t *__restrict__ nnz_total) {
    __shared__ long long part[1024];
    const int tid = threadIdx.x;
    const int64_t chunk = (n_rows + 1023) / 1024;
    const int64_t lo = tid * chunk, hi = lo + chunk < n_rows ? lo + chunk : n_rows;
    long long s = 0;
    for (int64_t r = lo; r < hi; ++r) s += row_cnt[r];
    part[tid] = s;
    __syncthreads();
    // Hillis-Steele inclusive scan over 1024 partial sums
    for (int o = 1; o < 1024; o <<= 1) {
        long long v = tid >= o ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    long long base = tid == 0 ? 0 : part[tid - 1];
    for (int64_t r = lo; r < hi; ++r) {
        row_ptr[r] = (int32_t)base;
        base += row_cnt[r];
    }
    if (tid == 1023) {
        row_ptr[n_rows] = (int32_t)part[1023];
        *nnz_total = part[1023];
    }
}

// ---- CSR emit: expand each row's bitmap in bit order ---------------------------------------
__global__ void __launch_bounds__(kThreads)
k_prm_emit(int64_t n_rows, const uint32_t *__restrict__ bitmap, int words,
           const int32_t *__restrict__ row_ptr, int32_t *__restrict__ col_idx, int64_t col_cap) {
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n_rows;
         row += n_warps) {
        int64_t base = row_ptr[row];
        for (int t = 0; t < words; ++t) {
            const uint32_t w = __ldg(bitmap + row * words + t);
            if ((w >> lane) & 1u) {
                const int64_t pos = base + __popc(w & ((1u << lane) - 1u));
                if (pos < col_cap) col_idx[pos] = 32 * t + lane;
            }
            base += __popc(w);
        }
    }
}

// ---- launchers ----------------------------------------------------------------------------
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

#define MRMP_MODEL_SWITCH(cv, EXPR)                      \
    switch ((cv).model) {                                \
    case 0: { using M = PointModel<2>; EXPR; } break;    \
    case 1: { using M = PointModel<3>; EXPR; } break;    \
    case 2: { using M = ArmModel; EXPR; } break;         \
    default: return cudaErrorNotSupported;               \
    }

cudaError_t launch_sample_states(const CtxView &cv, const LaunchCfg &lc, int agent, uint64_t seed,
                                 uint64_t t0, int64_t n, double *q_out) {
    if (n <= 0) return cudaSuccess;
    int64_t want = (n + kThreads - 1) / kThreads;
    int64_t cap = (int64_t)lc.sm_count * 8;
    const int grid = (int)(want < cap ? want : cap);
    MRMP_MODEL_SWITCH(cv, (k_sample_states<M><<<grid, kThreads, 0, lc.stream>>>(agent, seed, t0, n,
                                                                                 q_out)));
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_sample_free(const CtxView &cv, const LaunchCfg &lc, int agent0, int n_agents,
                               int n_fixed, int nv, int nv_stride, uint64_t seed, double *nodes,
                               int64_t *tries_out, int64_t max_tries) {
    if (n_agents <= 0) return cudaSuccess;
    const int sm = seg_smem_bytes(cv.con_len);
    cudaError_t e = cudaSuccess;
    MRMP_MODEL_SWITCH(cv, (e = set_smem(k_sample_free<M>, sm),
                           k_sample_free<M><<<n_agents, kThreads, sm, lc.stream>>>(
                               cv, agent0, n_fixed, nv, nv_stride, seed, nodes, tries_out,
                               max_tries, sm > 0)));
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_prm_adjacency(const CtxView &cv, const LaunchCfg &lc, int agent0, int n_agents,
                                 int nv, const double *rad_r2, const double *rad,
                                 const double *nodes, uint32_t *bitmap, int words,
                                 int32_t *row_cnt) {
    if (n_agents <= 0 || nv <= 0) return cudaSuccess;
    int seg = seg_smem_bytes(cv.con_len);
    const int scratch = kWarps * (words + 64) * 4;
    if (seg + scratch > 200 * 1024) seg = 0;  // very large nv: leave the segment in global
    if (scratch > 200 * 1024) return cudaErrorInvalidValue;
    const int sm = seg + scratch;
    const int64_t n_rows = (int64_t)n_agents * nv;
    int per_sm = (220 * 1024) / (sm > 0 ? sm : 1);
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    int64_t want = (n_rows + kWarps - 1) / kWarps;
    int64_t cap = (int64_t)lc.sm_count * per_sm;
    const int grid = (int)(want < cap ? want : cap);
    cudaError_t e = cudaSuccess;
    MRMP_MODEL_SWITCH(cv, (e = set_smem(k_prm_adjacency<M>, sm),
                           k_prm_adjacency<M><<<grid, kThreads, sm, lc.stream>>>(
                               cv, agent0, n_agents, nv, rad, rad_r2, nodes, bitmap, words,
                               row_cnt, seg > 0, seg)));
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_row_scan(const LaunchCfg &lc, int64_t n_rows, const int32_t *row_cnt,
                            int32_t *row_ptr, int64_t *nnz_total) {
    k_row_scan<<<1, 1024, 0, lc.stream>>>(n_rows, row_cnt, row_ptr, nnz_total);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_prm_emit(const LaunchCfg &lc, int64_t n_rows, int nv, const uint32_t *bitmap,
                            int words, const int32_t *row_ptr, int32_t *col_idx, int64_t col_cap) {
    (void)nv;
    if (n_rows <= 0) return cudaSuccess;
    int64_t want = (n_rows + kWarps - 1) / kWarps;
    int64_t cap = (int64_t)lc.sm_count * 8;
    const int grid = (int)(want < cap ? want : cap);
    k_prm_emit<<<grid, kThreads, 0, lc.stream>>>(n_rows, bitmap, words, row_ptr, col_idx, col_cap);
    count_launch();
    return cudaGetLastError();
}

}  // namespace mrmp
