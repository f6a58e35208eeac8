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

// ---- uniform-grid spatial hash (large roadmaps) ---------------------------------------------
// The all-pairs radius gate above is O(nv^2) per agent: fine for the reference's 10^2..10^3
// vertices, hopeless for 10^4..10^6.  For those the vertices of every agent are binned into a
// uniform grid with cell edge 1/g >= rad (g = floor(1/rad), so a vertex within rad of q_j can
// only sit in the 3^d cells around q_j's cell; coordinates outside the unit cube clamp into
// the border cells, which keeps that property because clamping is monotone).  Binning is a
// one-pass radix (counting) sort by cell id: histogram -> exclusive scan -> scatter.  Each row
// then tests only the vertices of its 3^d cells with the SAME exact gate + connect as the
// brute-force kernel, collects the survivors in a per-warp shared-memory bitmap over vertex
// ids and emits the set bits in ascending order, so the CSR is identical to the reference's
// row-major loop (prm.jl:202-210) whatever order the cells were visited in.
struct GridView {
    int g;                 // cells per dimension
    int cells;             // g^d
    int32_t *cell_start;   // [n_agents][cells + 1]  (counts, then exclusive offsets)
    int32_t *cursor;       // [n_agents][cells]
    int32_t *ids;          // [n_agents][nv] vertex ids grouped by cell
};

template <int D>
__device__ __forceinline__ int cell_coord(double x, int g) {
    const double c = floor(x * (double)g);
    return c < 0.0 ? 0 : (c >= (double)g ? g - 1 : (int)c);  // NaN -> last cell
}
template <int D>
__device__ __forceinline__ int cell_of(const double *q, int g) {
    int id = 0;
#pragma unroll
    for (int k = D - 1; k >= 0; --k) id = id * g + cell_coord<D>(q[k], g);
    return id;
}

template <int D>
__global__ void __launch_bounds__(kThreads)
k_grid_hist(GridView gv, int n_agents, int nv, const double *__restrict__ nodes, int scatter) {
    const int64_t total = (int64_t)n_agents * nv;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int a = (int)(t / nv), v = (int)(t - (int64_t)a * nv);
        double q[D];
        load_state<D>(nodes, t, q);
        const int c = cell_of<D>(q, gv.g);
        if (!scatter) {
            atomicAdd(gv.cell_start + (int64_t)a * (gv.cells + 1) + c, 1);
        } else {
            const int pos = atomicAdd(gv.cursor + (int64_t)a * gv.cells + c, 1);
            gv.ids[(int64_t)a * nv + pos] = v;
        }
    }
}

// per agent: counts -> exclusive offsets (cell_start[cells] = nv), cursor = offsets
__global__ void __launch_bounds__(1024)
k_grid_scan(GridView gv) {
    __shared__ int part[1024];
    int32_t *cs = gv.cell_start + (int64_t)blockIdx.x * (gv.cells + 1);
    int32_t *cur = gv.cursor + (int64_t)blockIdx.x * gv.cells;
    const int tid = threadIdx.x;
    const int chunk = (gv.cells + 1023) / 1024;
    const int lo = tid * chunk, hi = min(lo + chunk, gv.cells);
    int s = 0;
    for (int c = lo; c < hi; ++c) s += cs[c];
    part[tid] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = tid >= o ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    int base = tid == 0 ? 0 : part[tid - 1];
    for (int c = lo; c < hi; ++c) {
        const int n = cs[c];
        cs[c] = base;
        cur[c] = base;
        base += n;
    }
    if (tid == 1023) cs[gv.cells] = part[1023];
}

// warp per row; per-warp smem: bitmap[words] + queue[64]
template <class M>
__global__ void __launch_bounds__(kThreads)
k_prm_grid_adjacency(CtxView cv, GridView gv, int agent0, int n_agents, int nv,
                     const double *__restrict__ rad, const double *__restrict__ rad_r2,
                     const double *__restrict__ nodes, int words, int warps_per_cta,
                     int32_t *__restrict__ row_cnt, int64_t *__restrict__ row_off,
                     int32_t *__restrict__ tmp, int64_t tmp_cap,
                     unsigned long long *__restrict__ tmp_cursor, int staged, int seg_bytes) {
    constexpr int D = M::SD;
    const ConSeg cs(cv, stage_segment(cv, cv.con_off, cv.con_len, smem_raw, staged));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp >= warps_per_cta) return;
    uint32_t *wbits = reinterpret_cast<uint32_t *>(smem_raw + seg_bytes) + (size_t)warp * (words + 64);
    int *queue = reinterpret_cast<int *>(wbits + words);
    for (int t = lane; t < words; t += 32) wbits[t] = 0u;
    __syncwarp();
    const int64_t n_rows = (int64_t)n_agents * nv;
    const int64_t n_warps = (int64_t)gridDim.x * warps_per_cta;
    const int g = gv.g;
    for (int64_t row = (int64_t)blockIdx.x * warps_per_cta + warp; row < n_rows; row += n_warps) {
        const int a = (int)(row / nv), j = (int)(row - (int64_t)a * nv);
        const int agent = agent0 + a;
        const double *tab = nodes + (int64_t)a * nv * D;
        const int32_t *cstart = gv.cell_start + (int64_t)a * (gv.cells + 1);
        const int32_t *ids = gv.ids + (int64_t)a * nv;
        const double r = rad[a], r2 = rad_r2[a];
        double qj[D];
        load_state<D>(tab, j, qj);
        int cj[D];
#pragma unroll
        for (int k = 0; k < D; ++k) cj[k] = cell_coord<D>(qj[k], g);
        int qlen = 0, cnt = 0, lo = words, hi = -1;
        constexpr int NB = D == 2 ? 9 : 27;
        for (int nb = 0; nb < NB; ++nb) {
            int cc[D], rem = nb;
            bool inside = true;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                cc[k] = cj[k] + rem % 3 - 1;
                rem /= 3;
                inside &= cc[k] >= 0 && cc[k] < g;
            }
            if (!inside) continue;
            int cell = 0;
#pragma unroll
            for (int k = D - 1; k >= 0; --k) cell = cell * g + cc[k];
            const int beg = cstart[cell], end = cstart[cell + 1];
            for (int p0 = beg; p0 < end; p0 += 32) {
                const int p = p0 + lane;
                int k = -1;
                bool pass = false;
                double qk[D];
                if (p < end) {
                    k = ids[p];
                    if (k != j) {
                        load_state<D>(tab, k, qk);
                        pass = M::dist_le(cv, qj, qk, r, r2);  // prm.jl:207
                    }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, pass);
                if (pass) queue[(cnt + __popc(bal & ((1u << lane) - 1u))) & 63] = k;
                qlen += __popc(bal);
                cnt += __popc(bal);
                __syncwarp();
                if (qlen >= 32) {
                    const int k2 = queue[((cnt - qlen) + lane) & 63];
                    load_state<D>(tab, k2, qk);
                    if (M::connect_edge(cv, cs, qj, qk, agent)) {
                        atomicOr(&wbits[k2 >> 5], 1u << (k2 & 31));
                        lo = min(lo, k2 >> 5);
                        hi = max(hi, k2 >> 5);
                    }
                    qlen -= 32;
                    __syncwarp();
                }
            }
        }
        if (lane < qlen) {
            const int k2 = queue[((cnt - qlen) + lane) & 63];
            double qk[D];
            load_state<D>(tab, k2, qk);
            if (M::connect_edge(cv, cs, qj, qk, agent)) {
                atomicOr(&wbits[k2 >> 5], 1u << (k2 & 31));
                lo = min(lo, k2 >> 5);
                hi = max(hi, k2 >> 5);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        __syncwarp();
        // count, reserve, emit ascending, clear
        int total = 0;
        for (int t = lo + lane; t <= hi; t += 32) total += __popc(wbits[t]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
        long long off = 0;
        if (lane == 0) {
            off = total ? (long long)atomicAdd(tmp_cursor, (unsigned long long)total) : 0;
            row_cnt[row] = total;
            row_off[row] = off;
        }
        off = __shfl_sync(0xffffffffu, off, 0);
        const bool fits = off + total <= tmp_cap;
        long long base = off;
        for (int t0 = lo; t0 <= hi; t0 += 32) {
            const int t = t0 + lane;
            const uint32_t w = t <= hi ? wbits[t] : 0u;
            if (t <= hi) wbits[t] = 0u;
            int c = __popc(w), incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (fits) {
                long long pos = base + incl - c;
                uint32_t m = w;
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    tmp[pos++] = 32 * t + b;
                }
            }
            base += __shfl_sync(0xffffffffu, incl, 31);
        }
        __syncwarp();
    }
}

// rows were emitted at atomically reserved offsets: copy them into CSR order
__global__ void __launch_bounds__(kThreads)
k_grid_gather_rows(int64_t n_rows, const int32_t *__restrict__ row_ptr,
                   const int64_t *__restrict__ row_off, const int32_t *__restrict__ tmp,
                   int32_t *__restrict__ col_idx, int64_t col_cap) {
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n_rows;
         row += n_warps) {
        const int64_t dst = row_ptr[row], n = row_ptr[row + 1] - dst, src = row_off[row];
        for (int64_t t = lane; t < n; t += 32)
            if (dst + t < col_cap) col_idx[dst + t] = tmp[src + t];
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
    case 3: { using M = ChainModel<3>; EXPR; } break;    \
    case 4: { using M = ChainModel<4>; EXPR; } break;    \
    case 5: { using M = ChainModel<5>; EXPR; } break;    \
    case 6: { using M = ChainModel<6>; EXPR; } break;    \
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

// ---- CSR shards (multi-GPU): pack this rank's rows into a fixed-size slot, stitch all slots ---
// slot layout: int32 cnt[rows_cap] | int32 col[cap_nnz]; rows_cap = cap_agents * nv
__global__ void __launch_bounds__(kThreads)
k_csr_pack_shard(int64_t n_rows, int64_t rows_cap, const int32_t *__restrict__ row_ptr,
                 const int32_t *__restrict__ col_idx, int32_t *__restrict__ slot) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (int64_t r = tid; r < rows_cap; r += stride)
        slot[r] = r < n_rows ? row_ptr[r + 1] - row_ptr[r] : 0;
    const int64_t nnz = n_rows > 0 ? row_ptr[n_rows] : 0;
    for (int64_t e = tid; e < nnz; e += stride) slot[rows_cap + e] = col_idx[e];
}

__global__ void __launch_bounds__(kThreads)
k_csr_shard_counts(int64_t n_rows, int64_t rows_cap, int64_t slot_stride,
                   const int32_t *__restrict__ slots, int32_t *__restrict__ row_cnt) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_rows;
         r += (int64_t)gridDim.x * blockDim.x)
        row_cnt[r] = slots[(r / rows_cap) * slot_stride + r % rows_cap];
}

// row_ptr is the global exclusive scan; a row's offset inside its shard's column block is
// row_ptr[row] - row_ptr[first row of the shard]
__global__ void __launch_bounds__(kThreads)
k_csr_stitch(int64_t n_rows, int64_t rows_cap, int64_t slot_stride,
             const int32_t *__restrict__ slots, const int32_t *__restrict__ row_ptr,
             int32_t *__restrict__ col_idx, int64_t col_cap) {
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n_rows;
         row += n_warps) {
        const int64_t shard = row / rows_cap;
        const int64_t dst = row_ptr[row], n = row_ptr[row + 1] - dst;
        const int32_t *src = slots + shard * slot_stride + rows_cap + (dst - row_ptr[shard * rows_cap]);
        for (int64_t t = lane; t < n; t += 32)
            if (dst + t < col_cap) col_idx[dst + t] = src[t];
    }
}

// ---- grid path launchers -------------------------------------------------------------------
cudaError_t launch_grid_bin(const CtxView &cv, const LaunchCfg &lc, int n_agents, int nv, int g,
                            int cells, const double *nodes, int32_t *cell_start, int32_t *cursor,
                            int32_t *ids) {
    GridView gv{g, cells, cell_start, cursor, ids};
    cudaError_t e = cudaMemsetAsync(cell_start, 0, sizeof(int32_t) * (size_t)n_agents * (cells + 1),
                                    lc.stream);
    if (e != cudaSuccess) return e;
    const int64_t total = (int64_t)n_agents * nv;
    int64_t want = (total + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)lc.sm_count * 8;
    const int grid = (int)(want < cap ? want : cap);
    if (cv.model == 0) {
        k_grid_hist<2><<<grid, kThreads, 0, lc.stream>>>(gv, n_agents, nv, nodes, 0);
        k_grid_scan<<<n_agents, 1024, 0, lc.stream>>>(gv);
        k_grid_hist<2><<<grid, kThreads, 0, lc.stream>>>(gv, n_agents, nv, nodes, 1);
    } else if (cv.model == 1) {
        k_grid_hist<3><<<grid, kThreads, 0, lc.stream>>>(gv, n_agents, nv, nodes, 0);
        k_grid_scan<<<n_agents, 1024, 0, lc.stream>>>(gv);
        k_grid_hist<3><<<grid, kThreads, 0, lc.stream>>>(gv, n_agents, nv, nodes, 1);
    } else {
        return cudaErrorNotSupported;
    }
    count_launch(3);
    return cudaGetLastError();
}

cudaError_t launch_prm_grid_adjacency(const CtxView &cv, const LaunchCfg &lc, int agent0,
                                      int n_agents, int nv, int g, int cells,
                                      const int32_t *cell_start, const int32_t *ids,
                                      const double *rad_r2, const double *rad, const double *nodes,
                                      int32_t *row_cnt, int64_t *row_off, int32_t *tmp,
                                      int64_t tmp_cap, unsigned long long *tmp_cursor) {
    GridView gv{g, cells, const_cast<int32_t *>(cell_start), nullptr, const_cast<int32_t *>(ids)};
    const int words = (nv + 31) / 32;
    int seg = seg_smem_bytes(cv.con_len);
    const int per_warp = (words + 64) * 4;
    int warps = kWarps;
    while (warps > 1 && seg + warps * per_warp > 200 * 1024) warps >>= 1;
    if (seg + warps * per_warp > 200 * 1024) seg = 0;
    if (warps * per_warp > 220 * 1024) return cudaErrorInvalidValue;  // nv beyond ~1.7 M
    const int sm = seg + warps * per_warp;
    const int64_t n_rows = (int64_t)n_agents * nv;
    int per_sm = (220 * 1024) / (sm > 0 ? sm : 1);
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    int64_t want = (n_rows + warps - 1) / warps;
    const int64_t cap = (int64_t)lc.sm_count * per_sm;
    const int grid = (int)(want < cap ? want : cap);
    cudaError_t e = cudaMemsetAsync(tmp_cursor, 0, 8, lc.stream);
    if (e != cudaSuccess) return e;
    if (cv.model == 0) {
        using M = PointModel<2>;
        e = set_smem(k_prm_grid_adjacency<M>, sm);
        k_prm_grid_adjacency<M><<<grid, kThreads, sm, lc.stream>>>(
            cv, gv, agent0, n_agents, nv, rad, rad_r2, nodes, words, warps, row_cnt, row_off, tmp,
            tmp_cap, tmp_cursor, seg > 0, seg);
    } else if (cv.model == 1) {
        using M = PointModel<3>;
        e = set_smem(k_prm_grid_adjacency<M>, sm);
        k_prm_grid_adjacency<M><<<grid, kThreads, sm, lc.stream>>>(
            cv, gv, agent0, n_agents, nv, rad, rad_r2, nodes, words, warps, row_cnt, row_off, tmp,
            tmp_cap, tmp_cursor, seg > 0, seg);
    } else {
        return cudaErrorNotSupported;
    }
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_grid_gather_rows(const LaunchCfg &lc, int64_t n_rows, const int32_t *row_ptr,
                                    const int64_t *row_off, const int32_t *tmp, int32_t *col_idx,
                                    int64_t col_cap) {
    if (n_rows <= 0) return cudaSuccess;
    int64_t want = (n_rows + kWarps - 1) / kWarps;
    const int64_t cap = (int64_t)lc.sm_count * 8;
    const int grid = (int)(want < cap ? want : cap);
    k_grid_gather_rows<<<grid, kThreads, 0, lc.stream>>>(n_rows, row_ptr, row_off, tmp, col_idx,
                                                         col_cap);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_csr_pack_shard(const LaunchCfg &lc, int64_t n_rows, int64_t rows_cap,
                                  const int32_t *row_ptr, const int32_t *col_idx, int32_t *slot) {
    k_csr_pack_shard<<<lc.sm_count * 4, kThreads, 0, lc.stream>>>(n_rows, rows_cap, row_ptr, col_idx,
                                                                 slot);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_csr_shard_counts(const LaunchCfg &lc, int64_t n_rows, int64_t rows_cap,
                                    int64_t slot_stride, const int32_t *slots, int32_t *row_cnt) {
    k_csr_shard_counts<<<lc.sm_count * 2, kThreads, 0, lc.stream>>>(n_rows, rows_cap, slot_stride,
                                                                   slots, row_cnt);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_csr_stitch(const LaunchCfg &lc, int64_t n_rows, int64_t rows_cap,
                              int64_t slot_stride, const int32_t *slots, const int32_t *row_ptr,
                              int32_t *col_idx, int64_t col_cap) {
    k_csr_stitch<<<lc.sm_count * 4, kThreads, 0, lc.stream>>>(n_rows, rows_cap, slot_stride, slots,
                                                             row_ptr, col_idx, col_cap);
    count_launch();
    return cudaGetLastError();
}

}  // namespace mrmp
