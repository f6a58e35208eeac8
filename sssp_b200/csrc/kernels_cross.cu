// kernels_cross.cu -- cross-product collide: every row edge x every column edge -> bits.
//
// This is the batched form of the reference's inter-agent check
//   collide(qi_from, qi_to, qj_from, qj_to, i, j) = dist(seg_i, seg_j) < rads[i] + rads[j]
// (src/models/point.jl:9-12) for the call shapes that issue it in bulk: PP's `invalid`
// (pp.jl:179-189: candidate edge x planned agents' path edges), CBS's soft conflict count
// (cbs.jl:336-347), the TPG construction (smoother.jl:86-103), SSSP's successor filter
// (sssp.jl:196-223: candidate successor edges x other agents' current edges).
//
// Design (second version; the measurements that led here are in profiles/):
//  * Both sides are sorted by the Morton cell of their midpoint (one-pass radix / counting
//    sort over 4096 cells: histogram -> scan -> scatter).  32 consecutive rows then form a
//    spatially compact WARP TILE, 64 consecutive columns a compact COLUMN TILE with a stored
//    bounding box.
//  * A warp owns a warp tile (one row per lane, in registers).  For every column tile it first
//    tests box against box (lanes test 32 tile headers at once, one ballot): a column tile
//    whose box is farther from the warp's box than the largest possible reach is skipped
//    without being read.  Surviving tiles are read through L1/L2 with warp-uniform
//    (broadcast) loads -- all column tiles of a launch are a few hundred KB and shared by
//    every warp, so they stay cache resident.
//  * Broad phase per pair (8 FP64 ops): every point of a segment lies within h = L/2 of its
//    midpoint m, so dist(seg1, seg2) >= |m1 - m2| - h1 - h2; a pair with
//    |m1 - m2|^2 > (thr_max + 2^-20 + h1 + h2)^2 cannot collide.  2^-20 dwarfs the rounding of
//    both this bound and the exact path (coordinates up to 1e6), so culled pairs are pairs the
//    reference arithmetic also calls "no collision".  NaN/inf never pass the `>` test and
//    fall through to the exact path.
//  * Survivors are compacted into a per-warp shared-memory queue (popc + warp prefix), then
//    the exact reference arithmetic (geom.cuh: segseg_lt) runs on dense lanes.  No block-wide
//    barrier anywhere: warps are autonomous and pull warp tiles from an atomic counter.
// Output: full bit matrix (bit = original column index), or OR-reduced over consecutive
// column groups (e.g. all agents' path edges of one time step).
#include "point_model.cuh"

namespace mrmp {

extern __shared__ __align__(16) unsigned char smem_raw[];

constexpr int kCols = 32;        // column edges per tile (must equal kCrossCols)
constexpr int kRedWords = 8;     // reduced mode: up to 256 groups per launch
constexpr int kWarpsPerCta = 8;
constexpr int kSortCells = 4096;  // Morton grid: 64 x 64 (2-D), 16^3 (3-D)

template <int D>
struct RowRec {  // one row edge, in Morton order
    double a[D], b[D];
    int agent, orig;  // orig = row index in the caller's order
};

template <int D>
struct ColTile {                 // 64 column edges in Morton order; sizeof is a multiple of 16
    double mm[kCols][D];         // midpoints
    struct SA { double h; int agent; int bit; } sa[kCols];  // half length (rounded up), agent
                                 // id (-1 = padding), output bit (group or original column)
    double ab[kCols][2 * D];     // q_from, q_to (exact path operands)
};

template <int D>
struct TileHdr {  // bounding box of the tile's midpoints and its largest half length
    double lo[D], hi[D], hmax, pad;
};

// ---- Morton cell of an edge midpoint -------------------------------------------------------
__device__ __forceinline__ unsigned spread2(unsigned x) {  // 6 bits -> every 2nd bit
    x = (x | (x << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}
__device__ __forceinline__ unsigned spread3(unsigned x) {  // 4 bits -> every 3rd bit
    x = (x | (x << 8)) & 0x0000f00fu;
    x = (x | (x << 4)) & 0x000c30c3u;
    x = (x | (x << 2)) & 0x00249249u;
    return x;
}
template <int D>
__device__ __forceinline__ int morton_cell(const double *a, const double *b) {
    constexpr int G = D == 2 ? 64 : 16;
    unsigned c[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const double m = (a[k] + b[k]) * (0.5 * G);
        c[k] = m >= 0.0 ? (m < (double)G ? (unsigned)m : (unsigned)(G - 1)) : 0u;  // NaN -> 0
    }
    if constexpr (D == 2) return (int)(spread2(c[0]) | (spread2(c[1]) << 1));
    return (int)(spread3(c[0]) | (spread3(c[1]) << 1) | (spread3(c[2]) << 2));
}

// edge k of a list given either by explicit states or by vertex ids into a node table
template <int D>
struct EdgeList {
    const int32_t *agent;
    const double *qf, *qt;      // explicit states, or
    const int32_t *vf, *vt;     // ids into nodes
    const double *nodes;
    int agent0, nv_stride;
    __device__ __forceinline__ int get(int64_t k, double *a, double *b) const {
        const int j = __ldg(agent + k);
        if (nodes) {
            const int64_t base = (int64_t)(j - agent0) * nv_stride;
            load_state<D>(nodes, base + __ldg(vf + k), a);
            load_state<D>(nodes, base + __ldg(vt + k), b);
        } else {
            load_state<D>(qf, k, a);
            load_state<D>(qt, k, b);
        }
        return j;
    }
};

// ---- counting sort by Morton cell: histogram, scan, scatter ------------------------------------
template <int D>
__global__ void __launch_bounds__(256)
k_cell_hist(EdgeList<D> el, int64_t n, int32_t *__restrict__ counts) {
    __shared__ int h[kSortCells];
    for (int c = threadIdx.x; c < kSortCells; c += blockDim.x) h[c] = 0;
    __syncthreads();
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
         k += (int64_t)gridDim.x * blockDim.x) {
        double a[D], b[D];
        el.get(k, a, b);
        atomicAdd(&h[morton_cell<D>(a, b)], 1);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < kSortCells; c += blockDim.x)
        if (h[c]) atomicAdd(counts + c, h[c]);
}

// counts[4096] -> exclusive offsets in place (used as scatter cursors)
__global__ void __launch_bounds__(1024)
k_cell_scan(int32_t *__restrict__ counts) {
    __shared__ int part[1024];
    const int tid = threadIdx.x;
    int v[4], s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[k] = counts[4 * tid + k];
        s += v[k];
    }
    part[tid] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int t = tid >= o ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += t;
        __syncthreads();
    }
    int base = tid == 0 ? 0 : part[tid - 1];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        counts[4 * tid + k] = base;
        base += v[k];
    }
}

// rows: scatter whole records (order inside a cell is arbitrary; results are written back by
// `orig`, so the output does not depend on it)
template <int D>
__global__ void __launch_bounds__(256)
k_rows_scatter(EdgeList<D> el, int64_t n, int32_t *__restrict__ cursor,
               RowRec<D> *__restrict__ rows) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
         k += (int64_t)gridDim.x * blockDim.x) {
        RowRec<D> r;
        r.agent = el.get(k, r.a, r.b);
        r.orig = (int)k;
        rows[atomicAdd(cursor + morton_cell<D>(r.a, r.b), 1)] = r;
    }
}

// columns: scatter indices, the tiles are built from them by k_cross_prep_cols
template <int D>
__global__ void __launch_bounds__(256)
k_cols_scatter(EdgeList<D> el, int64_t n, int32_t *__restrict__ cursor,
               int32_t *__restrict__ order) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
         k += (int64_t)gridDim.x * blockDim.x) {
        double a[D], b[D];
        el.get(k, a, b);
        order[atomicAdd(cursor + morton_cell<D>(a, b), 1)] = (int)k;
    }
}

// ---- column tiles: one warp builds one tile (2 columns per lane) and its header ------------
template <int D>
__global__ void __launch_bounds__(256)
k_cross_prep_cols(EdgeList<D> el, int64_t n_cols, const int32_t *__restrict__ order,
                  int group_size, ColTile<D> *__restrict__ tiles, TileHdr<D> *__restrict__ hdrs,
                  int64_t n_tiles, const double *__restrict__ rads, int n_agents) {
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_tiles;
         t += n_warps) {
        ColTile<D> &T = tiles[t];
        double lo[D], hi[D], hmax = 0.0;
#pragma unroll
        for (int x = 0; x < D; ++x) {
            lo[x] = 1e300;
            hi[x] = -1e300;
        }
        bool poison = false;  // a NaN midpoint must never be culled by the box test
#pragma unroll
        for (int half = 0; half < kCols / 32; ++half) {
            const int k = half * 32 + lane;
            const int64_t c = t * kCols + k;
            if (c >= n_cols) {  // padding column: infinitely far, never selected
#pragma unroll
                for (int x = 0; x < D; ++x) T.mm[k][x] = 1e300;
                T.sa[k].h = 0.0;
                T.sa[k].agent = -1;
                T.sa[k].bit = 0;
#pragma unroll
                for (int x = 0; x < 2 * D; ++x) T.ab[k][x] = 0.0;
                continue;
            }
            const int64_t src = order[c];
            double a[D], b[D];
            const int j = el.get(src, a, b);
            // reach of the column: half length + its agent's radius (an id outside the ctx is
            // rejected by the exact stage; it must not be culled here)
            const double h = half_len_up<D>(a, b) + (j >= 0 && j < n_agents ? __ldg(rads + j) : 1e300);
#pragma unroll
            for (int x = 0; x < D; ++x) {
                const double m = dmul(dadd(a[x], b[x]), 0.5);
                T.mm[k][x] = m;
                T.ab[k][x] = a[x];
                T.ab[k][D + x] = b[x];
                lo[x] = fmin(lo[x], m);
                hi[x] = fmax(hi[x], m);
                poison |= !(m == m);
            }
            poison |= !(h == h);
            hmax = fmax(hmax, h);
            T.sa[k].h = h;
            T.sa[k].agent = j;
            T.sa[k].bit = (int)(group_size > 0 ? src / group_size : src);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int x = 0; x < D; ++x) {
                lo[x] = fmin(lo[x], __shfl_xor_sync(0xffffffffu, lo[x], o));
                hi[x] = fmax(hi[x], __shfl_xor_sync(0xffffffffu, hi[x], o));
            }
            hmax = fmax(hmax, __shfl_xor_sync(0xffffffffu, hmax, o));
        }
        poison = __any_sync(0xffffffffu, poison);
        if (lane == 0) {
            TileHdr<D> H;
#pragma unroll
            for (int x = 0; x < D; ++x) {
                H.lo[x] = poison ? -1e300 : lo[x];
                H.hi[x] = poison ? 1e300 : hi[x];
            }
            H.hmax = poison ? 1e300 : hmax;
            H.pad = 0.0;
            hdrs[t] = H;
        }
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
struct WarpSmem {
    double row_ab[32][2 * D];
    int row_agent[32];
    int row_orig[32];
    uint32_t obits[32][kRedWords];
    uint16_t queue[32 * kCols];
};

#ifndef MRMP_CROSS_MINB
#define MRMP_CROSS_MINB 3
#endif

template <int D, bool REDUCED>
__global__ void __launch_bounds__(32 * kWarpsPerCta, MRMP_CROSS_MINB)
k_cross_collide(CtxView cv, int64_t n_rows, const RowRec<D> *__restrict__ rows, int n_tiles,
                const ColTile<D> *__restrict__ tiles, const TileHdr<D> *__restrict__ hdrs,
                int words_per_row, int64_t out_stride, uint32_t *__restrict__ out, double max_rad,
                unsigned long long *__restrict__ stats, unsigned int *__restrict__ work) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpSmem<D> &S = reinterpret_cast<WarpSmem<D> *>(smem_raw)[warp];
    const double *rads = cv.blob + cv.crads_off;
    const double *t2tab = cv.blob + cv.t2_pair_off;
    const int N = cv.n_agents;
    const int64_t n_wt = (n_rows + 31) >> 5;
    unsigned long long n_tiles_read = 0, n_exact = 0;  // warp-uniform counters

    for (;;) {
        unsigned wt = 0;
        if (lane == 0) wt = atomicAdd(work, 1u);
        wt = __shfl_sync(0xffffffffu, wt, 0);
        if ((int64_t)wt >= n_wt) break;
        const int64_t r = (int64_t)wt * 32 + lane;
        const bool live = r < n_rows;
        int i = -2, orig = 0;
        double m1[D], R = 0.0;
        double u1[D], h1 = 0.0, Rc = 0.0;  // unit direction, half length, reach without h1
        double lo[D], hi[D];
#pragma unroll
        for (int x = 0; x < D; ++x) u1[x] = 0.0;
        if (live) {
            const RowRec<D> rec = rows[r];
            i = rec.agent;
            orig = rec.orig;
#pragma unroll
            for (int x = 0; x < D; ++x) {
                m1[x] = dmul(dadd(rec.a[x], rec.b[x]), 0.5);
                S.row_ab[lane][x] = rec.a[x];
                S.row_ab[lane][D + x] = rec.b[x];
                lo[x] = hi[x] = m1[x];
            }
            // reach of this row: its radius + slack + its own half length (the column's
            // radius and half length are in T.sa[c].h)
            h1 = half_len_up<D>(rec.a, rec.b);
            Rc = __ldg(rads + i) + kPad;
            R = Rc + h1;
            // direction of the row edge for the capsule bound (a zero-length edge keeps u1 = 0:
            // its capsule is the disc of the midpoint bound)
            double len2 = 0.0;
#pragma unroll
            for (int x = 0; x < D; ++x) len2 += (rec.b[x] - rec.a[x]) * (rec.b[x] - rec.a[x]);
            if (len2 > 0.0 && len2 < 1e300) {
                const double inv = 1.0 / sqrt(len2);
#pragma unroll
                for (int x = 0; x < D; ++x) u1[x] = (rec.b[x] - rec.a[x]) * inv;
            }
        } else {
#pragma unroll
            for (int x = 0; x < D; ++x) {
                m1[x] = 0.0;
                lo[x] = 1e300;
                hi[x] = -1e300;
            }
        }
        S.row_agent[lane] = i;
        S.row_orig[lane] = orig;
        if (REDUCED) {
#pragma unroll
            for (int w = 0; w < kRedWords; ++w) S.obits[lane][w] = 0u;
        }
        // warp box and largest reach; a NaN row disables the box test for the whole warp tile
        double Rmax = R;
        bool poison = live && !(R == R);
#pragma unroll
        for (int x = 0; x < D; ++x) poison |= live && !(m1[x] == m1[x]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int x = 0; x < D; ++x) {
                lo[x] = fmin(lo[x], __shfl_xor_sync(0xffffffffu, lo[x], o));
                hi[x] = fmax(hi[x], __shfl_xor_sync(0xffffffffu, hi[x], o));
            }
            Rmax = fmax(Rmax, __shfl_xor_sync(0xffffffffu, Rmax, o));
        }
        poison = __any_sync(0xffffffffu, poison);
        __syncwarp();

        for (int t0 = 0; t0 < n_tiles; t0 += 32) {
            // box-vs-box test of 32 column tiles at once
            bool near = false;
            if (t0 + lane < n_tiles) {
                const TileHdr<D> H = hdrs[t0 + lane];
                double d2 = 0.0;
#pragma unroll
                for (int x = 0; x < D; ++x) {
                    const double gap = fmax(fmax(H.lo[x] - hi[x], lo[x] - H.hi[x]), 0.0);
                    d2 += gap * gap;
                }
                const double reach = Rmax + H.hmax;
                near = poison || !(d2 > reach * reach);
            }
            unsigned tmask = __ballot_sync(0xffffffffu, near);
            while (tmask) {
                const int t = t0 + __ffs(tmask) - 1;
                tmask &= tmask - 1;
                const ColTile<D> &T = tiles[t];
                // ---- broad phase: this lane's row against the 64 columns of the tile
                unsigned long long mask = 0ull;
#pragma unroll 8
                for (int c = 0; c < kCols; ++c) {
                    // capsule bound: the column's midpoint against the row SEGMENT --
                    // dist(seg_r, seg_c) >= dist(m_c, seg_r) - h_c, and
                    // dist(m_c, seg_r)^2 = |w_perp|^2 + max(|w_par| - h_r, 0)^2 with w = m_c - m_r
                    double w2, par;
                    if constexpr (D == 2) {
                        const double2 mm = __ldg(reinterpret_cast<const double2 *>(&T.mm[c][0]));
                        const double wx = mm.x - m1[0], wy = mm.y - m1[1];
                        w2 = wx * wx + wy * wy;
                        par = wx * u1[0] + wy * u1[1];
                    } else {
                        w2 = 0.0;
                        par = 0.0;
#pragma unroll
                        for (int x = 0; x < D; ++x) {
                            const double w = __ldg(&T.mm[c][x]) - m1[x];
                            w2 += w * w;
                            par += w * u1[x];
                        }
                    }
                    const double ex = fmax(fabs(par) - h1, 0.0);
                    const double d2 = (w2 - par * par) + ex * ex;
                    const double reach = Rc + __ldg(&T.sa[c].h);
                    const bool far = d2 > reach * reach;
                    mask |= (unsigned long long)(far ? 0 : 1) << c;
                }
                if (!live) mask = 0ull;
                if (REDUCED) {
                    // the output is an OR over each column group: a pair whose bit this row
                    // has already set (by an earlier tile) cannot change it -- drop it here,
                    // before the compaction, so the exact stage stays dense
                    unsigned long long m2 = mask;
                    while (m2) {
                        const int c = __ffsll((long long)m2) - 1;
                        m2 &= m2 - 1;
                        const int bit = __ldg(&T.sa[c].bit);
                        if ((S.obits[lane][bit >> 5] >> (bit & 31)) & 1u) mask &= ~(1ull << c);
                    }
                }
                ++n_tiles_read;
                // ---- compaction into the warp queue
                const int cnt = __popcll(mask);
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                if (total == 0) continue;
                int base = incl - cnt;
                while (mask) {
                    const int c = __ffsll((long long)mask) - 1;
                    mask &= mask - 1;
                    S.queue[base++] = (uint16_t)((lane << 6) | c);
                }
                __syncwarp();
                n_exact += total;
                // ---- exact reference arithmetic on dense lanes
                for (int e = lane; e < total; e += 32) {
                    const int ent = S.queue[e];
                    const int rr = ent >> 6, cc = ent & 63;
                    const int ri = S.row_agent[rr], cj = __ldg(&T.sa[cc].agent);
                    if (cj < 0 || cj == ri) continue;  // padding / same agent: never a conflict
                    double ca[D], cb[D];
#pragma unroll
                    for (int x = 0; x < D; ++x) {
                        ca[x] = __ldg(&T.ab[cc][x]);
                        cb[x] = __ldg(&T.ab[cc][D + x]);
                    }
                    const double thr = dadd(__ldg(rads + ri), __ldg(rads + cj));  // point.jl:11
                    bool hit;
                    if (cv.has_pair_tab)
                        hit = segseg_lt<D, true>(&S.row_ab[rr][0], &S.row_ab[rr][D], ca, cb, thr,
                                                 __ldg(t2tab + ri * N + cj));
                    else
                        hit = segseg_lt<D, false>(&S.row_ab[rr][0], &S.row_ab[rr][D], ca, cb, thr,
                                                  0.0);
                    if (hit) {
                        const int bit = __ldg(&T.sa[cc].bit);
                        if (REDUCED)
                            atomicOr(&S.obits[rr][bit >> 5], 1u << (bit & 31));
                        else
                            atomicOr(out + (int64_t)S.row_orig[rr] * out_stride + (bit >> 5),
                                     1u << (bit & 31));
                    }
                }
                __syncwarp();
            }
        }
        if (REDUCED && live) {
            // every word of the row is written exactly once (the output may be mapped host
            // memory: plain stores, two words per transaction where the layout allows)
            uint32_t *dst = out + (int64_t)orig * out_stride;
            if ((words_per_row & 1) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
                for (int w = 0; w < words_per_row; w += 2)
                    *reinterpret_cast<uint2 *>(dst + w) = make_uint2(S.obits[lane][w], S.obits[lane][w + 1]);
            } else {
                for (int w = 0; w < words_per_row; ++w) dst[w] = S.obits[lane][w];
            }
        }
        __syncwarp();
    }
    if (stats && lane == 0) {
        if (n_exact) atomicAdd(stats, n_exact);                          // pairs on the exact path
        if (n_tiles_read) atomicAdd(stats + 2, n_tiles_read * 32ull * kCols);  // pairs bound-tested
    }
}

// ---- launchers --------------------------------------------------------------------------------
size_t cross_tile_bytes(int d) { return d == 2 ? sizeof(ColTile<2>) : sizeof(ColTile<3>); }
size_t cross_hdr_bytes(int d) { return d == 2 ? sizeof(TileHdr<2>) : sizeof(TileHdr<3>); }
size_t cross_row_bytes(int d) { return d == 2 ? sizeof(RowRec<2>) : sizeof(RowRec<3>); }

static inline int grid_for_n(const LaunchCfg &lc, int64_t n, int threads) {
    int64_t want = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)lc.sm_count * 8;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

template <int D>
static cudaError_t sort_rows_t(const LaunchCfg &lc, const EdgeList<D> &el, int64_t n,
                               int32_t *cells, void *rows_out) {
    cudaError_t e = cudaMemsetAsync(cells, 0, sizeof(int32_t) * kSortCells, lc.stream);
    if (e != cudaSuccess) return e;
    const int grid = grid_for_n(lc, n, 256 * 4);
    k_cell_hist<D><<<grid, 256, 0, lc.stream>>>(el, n, cells);
    k_cell_scan<<<1, 1024, 0, lc.stream>>>(cells);
    k_rows_scatter<D><<<grid, 256, 0, lc.stream>>>(el, n, cells, (RowRec<D> *)rows_out);
    count_launch(3);
    return cudaGetLastError();
}

// rows (explicit states) -> Morton-sorted RowRec array; cells = 4096 ints of device scratch
cudaError_t launch_cross_sort_rows(const CtxView &cv, const LaunchCfg &lc, int64_t n_rows,
                                   const int32_t *agent, const double *qf, const double *qt,
                                   int32_t *cells, void *rows_out) {
    if (n_rows <= 0) return cudaSuccess;
    if (n_rows >= ((int64_t)1 << 31)) return cudaErrorInvalidValue;
    if (cv.model == 0)
        return sort_rows_t<2>(lc, EdgeList<2>{agent, qf, qt, nullptr, nullptr, nullptr, 0, 0},
                              n_rows, cells, rows_out);
    if (cv.model == 1)
        return sort_rows_t<3>(lc, EdgeList<3>{agent, qf, qt, nullptr, nullptr, nullptr, 0, 0},
                              n_rows, cells, rows_out);
    return cudaErrorNotSupported;
}

template <int D>
static cudaError_t prep_cols_t(const CtxView &cv, const LaunchCfg &lc, const EdgeList<D> &el,
                               int64_t n_cols, int group_size, int32_t *cells, int32_t *order,
                               void *tiles, void *hdrs) {
    const int64_t n_tiles = (n_cols + kCols - 1) / kCols;
    cudaError_t e = cudaMemsetAsync(cells, 0, sizeof(int32_t) * kSortCells, lc.stream);
    if (e != cudaSuccess) return e;
    const int grid = grid_for_n(lc, n_cols, 256 * 4);
    k_cell_hist<D><<<grid, 256, 0, lc.stream>>>(el, n_cols, cells);
    k_cell_scan<<<1, 1024, 0, lc.stream>>>(cells);
    k_cols_scatter<D><<<grid, 256, 0, lc.stream>>>(el, n_cols, cells, order);
    k_cross_prep_cols<D><<<grid_for_n(lc, n_tiles * 32, 256), 256, 0, lc.stream>>>(
        el, n_cols, order, group_size, (ColTile<D> *)tiles, (TileHdr<D> *)hdrs, n_tiles,
        cv.blob + cv.crads_off, cv.n_agents);
    count_launch(4);
    return cudaGetLastError();
}

// columns (explicit states or ids into a node table) -> Morton-sorted tiles + headers
cudaError_t launch_cross_prep_cols(const CtxView &cv, const LaunchCfg &lc, int64_t n_cols,
                                   const int32_t *agent, const double *qf, const double *qt,
                                   const int32_t *vf, const int32_t *vt, const double *nodes,
                                   int agent0, int nv_stride, int group_size, int32_t *cells,
                                   int32_t *order, void *tiles, void *hdrs) {
    if (n_cols <= 0) return cudaSuccess;
    if (n_cols >= ((int64_t)1 << 31)) return cudaErrorInvalidValue;
    if (cv.model == 0)
        return prep_cols_t<2>(cv, lc, EdgeList<2>{agent, qf, qt, vf, vt, nodes, agent0, nv_stride},
                              n_cols, group_size, cells, order, tiles, hdrs);
    if (cv.model == 1)
        return prep_cols_t<3>(cv, lc, EdgeList<3>{agent, qf, qt, vf, vt, nodes, agent0, nv_stride},
                              n_cols, group_size, cells, order, tiles, hdrs);
    return cudaErrorNotSupported;
}

cudaError_t launch_csr_expand_edges(const CtxView &cv, const LaunchCfg &lc, int64_t n_rows_csr,
                                    int nv, int agent0, const int32_t *row_ptr,
                                    const int32_t *col_idx, const double *nodes, int32_t *e_agent,
                                    double *e_from, double *e_to) {
    if (n_rows_csr <= 0) return cudaSuccess;
    int64_t want = (n_rows_csr + 7) / 8;
    int64_t cap = (int64_t)lc.sm_count * 8;
    const int grid = (int)(want < cap ? want : cap);
    if (cv.d == 2)  // StatePoint2D, StateArm22
        k_csr_expand_edges<2><<<grid, 256, 0, lc.stream>>>(n_rows_csr, nv, agent0, row_ptr, col_idx,
                                                          nodes, e_agent, e_from, e_to);
    else if (cv.d == 3)  // StatePoint3D, StateLine2D
        k_csr_expand_edges<3><<<grid, 256, 0, lc.stream>>>(n_rows_csr, nv, agent0, row_ptr, col_idx,
                                                          nodes, e_agent, e_from, e_to);
    else if (cv.d == 6)  // StateSnake2D, StateArm33, StateCapsule3D
        k_csr_expand_edges<6><<<grid, 256, 0, lc.stream>>>(n_rows_csr, nv, agent0, row_ptr, col_idx,
                                                          nodes, e_agent, e_from, e_to);
    else
        return cudaErrorNotSupported;
    count_launch();
    return cudaGetLastError();
}

template <int D>
static cudaError_t launch_cross_t(const CtxView &cv, const LaunchCfg &lc, int64_t n_rows,
                                  const void *rows, int64_t n_cols, const void *tiles,
                                  const void *hdrs, int group_size, int words_per_row,
                                  int64_t out_stride, uint32_t *out, double max_rad,
                                  unsigned long long *stats, unsigned int *work) {
    const int n_tiles = (int)((n_cols + kCols - 1) / kCols);
    const int smem = (int)sizeof(WarpSmem<D>) * kWarpsPerCta;
    int per_sm = (220 * 1024) / smem;
    if (per_sm > MRMP_CROSS_MINB) per_sm = MRMP_CROSS_MINB;  // register-limited residency
    if (per_sm < 1) per_sm = 1;
    const int64_t n_wt = (n_rows + 31) / 32;
    int64_t want = (n_wt + kWarpsPerCta - 1) / kWarpsPerCta;
    const int64_t cap = (int64_t)lc.sm_count * per_sm;
    const int grid = (int)(want < cap ? want : cap);
    cudaError_t e = cudaMemsetAsync(work, 0, sizeof(unsigned int), lc.stream);
    if (e != cudaSuccess) return e;
    if (group_size > 0) {
        e = cudaFuncSetAttribute(k_cross_collide<D, true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        k_cross_collide<D, true><<<grid, 32 * kWarpsPerCta, smem, lc.stream>>>(
            cv, n_rows, (const RowRec<D> *)rows, n_tiles, (const ColTile<D> *)tiles,
            (const TileHdr<D> *)hdrs, words_per_row, out_stride, out, max_rad, stats, work);
    } else {
        e = cudaFuncSetAttribute(k_cross_collide<D, false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        k_cross_collide<D, false><<<grid, 32 * kWarpsPerCta, smem, lc.stream>>>(
            cv, n_rows, (const RowRec<D> *)rows, n_tiles, (const ColTile<D> *)tiles,
            (const TileHdr<D> *)hdrs, words_per_row, out_stride, out, max_rad, stats, work);
    }
    count_launch();
    return cudaGetLastError();
}

// group_size == 0: full bit matrix (words_per_row = ceil(n_cols/32), `out` must be zeroed by the
// caller: hits are OR-ed in); otherwise OR-reduced over consecutive groups of group_size
// columns (at most 32*kRedWords groups per launch, every word of a row is written).
cudaError_t launch_cross_collide(const CtxView &cv, const LaunchCfg &lc, int64_t n_rows,
                                 const void *rows, int64_t n_cols, const void *tiles,
                                 const void *hdrs, int group_size, int words_per_row,
                                 int64_t out_stride, uint32_t *out, double max_rad,
                                 unsigned long long *stats, unsigned int *work) {
    if (n_rows <= 0 || n_cols <= 0) return cudaSuccess;
    if (group_size > 0 && words_per_row > kRedWords) return cudaErrorInvalidValue;
    if (cv.model == 0)
        return launch_cross_t<2>(cv, lc, n_rows, rows, n_cols, tiles, hdrs, group_size,
                                 words_per_row, out_stride, out, max_rad, stats, work);
    if (cv.model == 1)
        return launch_cross_t<3>(cv, lc, n_rows, rows, n_cols, tiles, hdrs, group_size,
                                 words_per_row, out_stride, out, max_rad, stats, work);
    return cudaErrorNotSupported;
}

}  // namespace mrmp
