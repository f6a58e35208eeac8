// kernels_cross.cu -- cross-product collide: every row edge x every column edge -> bits.
//
// This is the batched form of the reference's inter-agent check
//   collide(qi_from, qi_to, qj_from, qj_to, i, j) = dist(seg_i, seg_j) < rads[i] + rads[j]
// (src/models/point.jl:9-12) for the call shapes that issue it in bulk: PP's `invalid`
// (pp.jl:179-189: candidate edge x planned agents' path edges), CBS's soft conflict count
// (cbs.jl:336-347), the TPG construction (smoother.jl:86-103), SSSP's successor filter
// (sssp.jl:196-223: candidate successor edges x other agents' current edges).
//
// Per CTA: a tile of 256 row edges (one per thread, in registers) sweeps all column edges,
// which stream through shared memory in 64-column tiles staged by 1-D TMA bulk copies
// (double-buffered, mbarrier completion).  Two phases per column tile:
//   phase 1 (broad phase, ~7 FP64 ops per pair): midpoint-distance bound.  Every point of a
//     segment lies within L/2 of its midpoint m, so
//        dist(seg1, seg2) >= |m1 - m2| - L1/2 - L2/2,
//     and with (x+y+z)^2 <= 3(x^2+y^2+z^2):  |m1-m2|^2 > 3((thr+pad)^2 + L1^2/4 + L2^2/4)
//     implies dist > thr + pad.  pad = 2^-20 dwarfs the rounding error of both this bound
//     and the exact path (coordinates O(1)), so culled pairs are exactly the ones the
//     reference arithmetic would also call "no collision".  NaN/inf never pass the `>` test
//     and fall through to the exact path.
//   phase 2 (exact): survivors are compacted into a shared-memory queue and distributed
//     round-robin over all 256 threads, each running the bit-exact reference arithmetic
//     (geom.cuh: segseg_lt).  Hits are OR-ed into a shared-memory bit tile.
// Output: full bit matrix, or OR-reduced over consecutive column groups (e.g. all agents'
// path edges of one time step).
#include "point_model.cuh"

namespace mrmp {

extern __shared__ __align__(16) unsigned char smem_raw[];

constexpr int kRows = 256;      // row edges per CTA == threads per CTA
constexpr int kCols = 64;       // column edges per smem tile
constexpr int kRedWords = 8;    // reduced mode: up to 256 groups per launch

template <int D>
struct ColTile {                 // one staged tile; sizeof is a multiple of 16
    double mm[kCols][D];         // midpoints
    struct SA { double s2; int agent; int pad; } sa[kCols];  // 0.75*L2^2*(1+1e-9), agent id
    double ab[kCols][2 * D];     // q_from, q_to (exact path operands)
};

__device__ __forceinline__ void mbar_init(uint64_t *bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load(void *smem_dst, const void *gsrc, uint32_t bytes,
                                         uint64_t *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- column preparation: explicit (agent, from, to) or ids into a roadmap node table ------
template <int D>
__global__ void __launch_bounds__(256)
k_cross_prep_cols(int64_t n_cols, const int32_t *__restrict__ agent,
                  const double *__restrict__ qf, const double *__restrict__ qt,
                  const int32_t *__restrict__ vf, const int32_t *__restrict__ vt,
                  const double *__restrict__ nodes, int agent0, int nv_stride,
                  ColTile<D> *__restrict__ tiles, int64_t n_tiles) {
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n_tiles * kCols;
         c += (int64_t)gridDim.x * blockDim.x) {
        ColTile<D> &T = tiles[c / kCols];
        const int k = (int)(c % kCols);
        if (c >= n_cols) {  // padding column: never selected
#pragma unroll
            for (int x = 0; x < D; ++x) T.mm[k][x] = 0.0;
            T.sa[k].s2 = 0.0;
            T.sa[k].agent = -1;
            T.sa[k].pad = 0;
#pragma unroll
            for (int x = 0; x < 2 * D; ++x) T.ab[k][x] = 0.0;
            continue;
        }
        const int j = agent[c];
        double a[D], b[D];
        if (nodes) {
            const int64_t base = (int64_t)(j - agent0) * nv_stride;
            load_state<D>(nodes, base + vf[c], a);
            load_state<D>(nodes, base + vt[c], b);
        } else {
            load_state<D>(qf, c, a);
            load_state<D>(qt, c, b);
        }
        double v[D];
#pragma unroll
        for (int x = 0; x < D; ++x) {
            v[x] = dsub(b[x], a[x]);
            T.mm[k][x] = dmul(dadd(a[x], b[x]), 0.5);
            T.ab[k][x] = a[x];
            T.ab[k][D + x] = b[x];
        }
        T.sa[k].s2 = dmul(dmul(0.75, dotd<D>(v, v)), 1.0 + 1e-9);
        T.sa[k].agent = j;
        T.sa[k].pad = 0;
    }
}

// ---- rows from the CSR of a roadmap set: edge e of row (agent a, vertex v) -----------------
template <int D>
__global__ void __launch_bounds__(256)
k_csr_expand_edges(int64_t n_rows_csr, int nv, int agent0, const int32_t *__restrict__ row_ptr,
                   const int32_t *__restrict__ col_idx, const double *__restrict__ nodes,
                   int32_t *__restrict__ e_agent, double *__restrict__ e_from,
                   double *__restrict__ e_to) {
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n_rows_csr;
         row += n_warps) {
        const int a = (int)(row / nv), v = (int)(row - (int64_t)a * nv);
        const double *tab = nodes + (int64_t)a * nv * D;
        double qv[D];
        load_state<D>(tab, v, qv);
        for (int e = row_ptr[row] + lane; e < row_ptr[row + 1]; e += 32) {
            double qk[D];
            load_state<D>(tab, col_idx[e], qk);
            e_agent[e] = agent0 + a;
#pragma unroll
            for (int x = 0; x < D; ++x) {
                e_from[(int64_t)e * D + x] = qv[x];
                e_to[(int64_t)e * D + x] = qk[x];
            }
        }
    }
}

// ---- the cross kernel -------------------------------------------------------------------------
template <int D>
struct CrossSmem {
    uint64_t bar[2];
    ColTile<D> tile[2];
    double row_ab[kRows][2 * D];
    int row_agent[kRows];
    uint32_t obits[kRows][kRedWords];
    int warp_tot[kRows / 32];
    uint16_t queue[kRows * kCols];
};

template <int D, bool REDUCED>
__global__ void __launch_bounds__(kRows)
k_cross_collide(CtxView cv, int64_t n_rows, const int32_t *__restrict__ row_agent,
                const double *__restrict__ row_from, const double *__restrict__ row_to,
                int64_t n_cols, const ColTile<D> *__restrict__ tiles, int n_tiles,
                int group_size, int words_per_row, int64_t out_stride, uint32_t *__restrict__ out,
                double max_rad, unsigned long long *__restrict__ stats) {
    CrossSmem<D> &S = *reinterpret_cast<CrossSmem<D> *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *rads = cv.blob + cv.crads_off;
    const double *t2tab = cv.blob + cv.t2_pair_off;
    const int N = cv.n_agents;

    if (tid == 0) {
        mbar_init(&S.bar[0]);
        mbar_init(&S.bar[1]);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t phase[2] = {0u, 0u};

    for (int64_t row0 = (int64_t)blockIdx.x * kRows; row0 < n_rows;
         row0 += (int64_t)gridDim.x * kRows) {
        const int64_t r = row0 + tid;
        const bool live = r < n_rows;
        int i = -2;
        double m1[D], s1 = 0.0;
        if (live) {
            double a[D], b[D], v[D];
            i = __ldg(row_agent + r);
            load_state<D>(row_from, r, a);
            load_state<D>(row_to, r, b);
#pragma unroll
            for (int x = 0; x < D; ++x) {
                v[x] = dsub(b[x], a[x]);
                m1[x] = dmul(dadd(a[x], b[x]), 0.5);
                S.row_ab[tid][x] = a[x];
                S.row_ab[tid][D + x] = b[x];
            }
            const double tp = dadd(dadd(__ldg(rads + i), max_rad), 9.5367431640625e-07);
            s1 = dmul(dadd(dmul(3.0, dmul(tp, tp)), dmul(0.75, dotd<D>(v, v))), 1.0 + 1e-9);
        } else {
#pragma unroll
            for (int x = 0; x < D; ++x) m1[x] = 0.0;
        }
        S.row_agent[tid] = i;
        if (REDUCED) {
#pragma unroll
            for (int w = 0; w < kRedWords; ++w) S.obits[tid][w] = 0u;
        }
        // tile pipeline prologue.  Parities continue across row tiles: tile index tt counts
        // every staged tile of this CTA.
        __syncthreads();
        if (tid == 0) tma_load(&S.tile[0], &tiles[0], sizeof(ColTile<D>), &S.bar[0]);
        for (int t = 0; t < n_tiles; ++t) {
            const int buf = t & 1;
            if (tid == 0 && t + 1 < n_tiles)
                tma_load(&S.tile[buf ^ 1], &tiles[t + 1], sizeof(ColTile<D>), &S.bar[buf ^ 1]);
            // every use of a barrier completes one phase; phase[] tracks the parity to wait for
            mbar_wait(&S.bar[buf], phase[buf]);
            phase[buf] ^= 1u;
            const ColTile<D> &T = S.tile[buf];
            if (!REDUCED) {
                S.obits[tid][0] = 0u;
                S.obits[tid][1] = 0u;
            }
            // ---- phase 1: broad phase over the 64 columns of the tile
            uint32_t mask[2] = {0u, 0u};
            if (live) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t mk = 0u;
#pragma unroll 8
                    for (int c = 0; c < 32; ++c) {
                        const int cc = h * 32 + c;
                        double w2;
                        if constexpr (D == 2) {
                            const double2 mm = *reinterpret_cast<const double2 *>(&T.mm[cc][0]);
                            const double wx = dsub(mm.x, m1[0]), wy = dsub(mm.y, m1[1]);
                            w2 = dadd(dmul(wx, wx), dmul(wy, wy));
                        } else {
                            double w[D];
#pragma unroll
                            for (int x = 0; x < D; ++x) w[x] = dsub(T.mm[cc][x], m1[x]);
                            w2 = sumsq<D>(w);
                        }
                        const double s2 = T.sa[cc].s2;
                        const int j = T.sa[cc].agent;
                        const bool far = w2 > dadd(s1, s2);
                        const bool cand = !far && j != i && j >= 0;
                        mk |= (cand ? 1u : 0u) << c;
                    }
                    mask[h] = mk;
                }
            }
            // ---- compaction: block-wide exclusive scan of the survivor counts
            const int cnt = __popc(mask[0]) + __popc(mask[1]);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) S.warp_tot[warp] = incl;
            __syncthreads();
            int base = incl - cnt;
            int total = 0;
#pragma unroll
            for (int w = 0; w < kRows / 32; ++w) {
                const int wt = S.warp_tot[w];
                if (w < warp) base += wt;
                total += wt;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t mk = mask[h];
                while (mk) {
                    const int c = __ffs(mk) - 1;
                    mk &= mk - 1;
                    S.queue[base++] = (uint16_t)((tid << 6) | (h * 32 + c));
                }
            }
            __syncthreads();
            if (stats && tid == 0 && total) atomicAdd(stats, (unsigned long long)total);
            // ---- phase 2: exact reference arithmetic on the compacted survivors
            for (int e = tid; e < total; e += kRows) {
                const int ent = S.queue[e];
                const int rr = ent >> 6, cc = ent & 63;
                const int ri = S.row_agent[rr], cj = T.sa[cc].agent;
                const double thr = dadd(__ldg(rads + ri), __ldg(rads + cj));  // point.jl:11
                bool hit;
                if (cv.has_pair_tab)
                    hit = segseg_lt<D, true>(&S.row_ab[rr][0], &S.row_ab[rr][D], &T.ab[cc][0],
                                             &T.ab[cc][D], thr, __ldg(t2tab + ri * N + cj));
                else
                    hit = segseg_lt<D, false>(&S.row_ab[rr][0], &S.row_ab[rr][D], &T.ab[cc][0],
                                              &T.ab[cc][D], thr, 0.0);
                if (hit) {
                    if (REDUCED) {
                        const int g = (int)(((int64_t)t * kCols + cc) / group_size);
                        atomicOr(&S.obits[rr][g >> 5], 1u << (g & 31));
                    } else {
                        atomicOr(&S.obits[rr][cc >> 5], 1u << (cc & 31));
                    }
                }
            }
            __syncthreads();
            if (!REDUCED && live) {
                const int64_t w0 = (int64_t)t * 2;
                if (w0 < words_per_row) out[r * out_stride + w0] = S.obits[tid][0];
                if (w0 + 1 < words_per_row) out[r * out_stride + w0 + 1] = S.obits[tid][1];
            }
        }
        if (REDUCED && live) {
            for (int w = 0; w < words_per_row; ++w) out[r * out_stride + w] = S.obits[tid][w];
        }
        __syncthreads();
    }
}


// ---- launchers --------------------------------------------------------------------------------
size_t cross_tile_bytes(int d) { return d == 2 ? sizeof(ColTile<2>) : sizeof(ColTile<3>); }

cudaError_t launch_cross_prep_cols(const CtxView &cv, const LaunchCfg &lc, int64_t n_cols,
                                   const int32_t *agent, const double *qf, const double *qt,
                                   const int32_t *vf, const int32_t *vt, const double *nodes,
                                   int agent0, int nv_stride, void *tiles) {
    const int64_t n_tiles = (n_cols + kCols - 1) / kCols;
    if (n_tiles <= 0) return cudaSuccess;
    int64_t want = (n_tiles * kCols + 255) / 256;
    int64_t cap = (int64_t)lc.sm_count * 8;
    const int grid = (int)(want < cap ? want : cap);
    if (cv.model == 0)
        k_cross_prep_cols<2><<<grid, 256, 0, lc.stream>>>(n_cols, agent, qf, qt, vf, vt, nodes, agent0,
                                                         nv_stride, (ColTile<2> *)tiles, n_tiles);
    else if (cv.model == 1)
        k_cross_prep_cols<3><<<grid, 256, 0, lc.stream>>>(n_cols, agent, qf, qt, vf, vt, nodes, agent0,
                                                         nv_stride, (ColTile<3> *)tiles, n_tiles);
    else
        return cudaErrorNotSupported;
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_csr_expand_edges(const CtxView &cv, const LaunchCfg &lc, int64_t n_rows_csr,
                                    int nv, int agent0, const int32_t *row_ptr,
                                    const int32_t *col_idx, const double *nodes, int32_t *e_agent,
                                    double *e_from, double *e_to) {
    if (n_rows_csr <= 0) return cudaSuccess;
    int64_t want = (n_rows_csr + 7) / 8;
    int64_t cap = (int64_t)lc.sm_count * 8;
    const int grid = (int)(want < cap ? want : cap);
    if (cv.model == 0 || cv.model == 2)  // StateArm22 is two doubles as well
        k_csr_expand_edges<2><<<grid, 256, 0, lc.stream>>>(n_rows_csr, nv, agent0, row_ptr, col_idx,
                                                          nodes, e_agent, e_from, e_to);
    else if (cv.model == 1)
        k_csr_expand_edges<3><<<grid, 256, 0, lc.stream>>>(n_rows_csr, nv, agent0, row_ptr, col_idx,
                                                          nodes, e_agent, e_from, e_to);
    else
        return cudaErrorNotSupported;
    count_launch();
    return cudaGetLastError();
}

template <int D>
static cudaError_t launch_cross_t(const CtxView &cv, const LaunchCfg &lc, int64_t n_rows,
                                  const int32_t *row_agent, const double *row_from,
                                  const double *row_to, int64_t n_cols, const void *tiles,
                                  int group_size, int words_per_row, int64_t out_stride,
                                  uint32_t *out, double max_rad, unsigned long long *stats) {
    const int n_tiles = (int)((n_cols + kCols - 1) / kCols);
    const int smem = (int)sizeof(CrossSmem<D>);
    int per_sm = (220 * 1024) / smem;
    if (per_sm < 1) per_sm = 1;
    int64_t want = (n_rows + kRows - 1) / kRows;
    int64_t cap = (int64_t)lc.sm_count * per_sm;
    const int grid = (int)(want < cap ? want : cap);
    cudaError_t e;
    if (group_size > 0) {
        e = cudaFuncSetAttribute(k_cross_collide<D, true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        k_cross_collide<D, true><<<grid, kRows, smem, lc.stream>>>(
            cv, n_rows, row_agent, row_from, row_to, n_cols, (const ColTile<D> *)tiles, n_tiles,
            group_size, words_per_row, out_stride, out, max_rad, stats);
    } else {
        e = cudaFuncSetAttribute(k_cross_collide<D, false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        k_cross_collide<D, false><<<grid, kRows, smem, lc.stream>>>(
            cv, n_rows, row_agent, row_from, row_to, n_cols, (const ColTile<D> *)tiles, n_tiles, 1,
            words_per_row, out_stride, out, max_rad, stats);
    }
    count_launch();
    return cudaGetLastError();
}

// group_size == 0: full bit matrix (words_per_row = ceil(n_cols/32)); otherwise OR-reduced
// over consecutive groups of group_size columns (at most 32*kRedWords groups per launch).
cudaError_t launch_cross_collide(const CtxView &cv, const LaunchCfg &lc, int64_t n_rows,
                                 const int32_t *row_agent, const double *row_from,
                                 const double *row_to, int64_t n_cols, const void *tiles,
                                 int group_size, int words_per_row, int64_t out_stride,
                                 uint32_t *out, double max_rad, unsigned long long *stats) {
    if (n_rows <= 0 || n_cols <= 0) return cudaSuccess;
    if (group_size > 0 && words_per_row > kRedWords) return cudaErrorInvalidValue;
    if (cv.model == 0)
        return launch_cross_t<2>(cv, lc, n_rows, row_agent, row_from, row_to, n_cols, tiles,
                                 group_size, words_per_row, out_stride, out, max_rad, stats);
    if (cv.model == 1)
        return launch_cross_t<3>(cv, lc, n_rows, row_agent, row_from, row_to, n_cols, tiles,
                                 group_size, words_per_row, out_stride, out, max_rad, stats);
    return cudaErrorNotSupported;
}

}  // namespace mrmp
