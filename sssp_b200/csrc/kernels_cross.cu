// kernels_cross.cu -- cross-product collide: every row edge x every column edge -> bits / counts.
//
// This is the batched form of the reference's inter-agent check
//   collide(qi_from, qi_to, qj_from, qj_to, i, j) = dist(seg_i, seg_j) < rads[i] + rads[j]
// (src/models/point.jl:9-12) for the call shapes that issue it in bulk: PP's `invalid`
// (pp.jl:179-189: candidate edge x planned agents' path edges), CBS's soft conflict count
// (cbs.jl:336-347), the TPG construction (smoother.jl:78-106), SSSP's successor filter
// (sssp.jl:196-223: candidate successor edges x other agents' current edges).
//
// Design (third version; the measurements that led here are in profiles/):
//  * Both sides are sorted by the Morton cell of their midpoint (counting sort over 4096 cells:
//    histogram -> scan -> scatter).  32 consecutive rows form a spatially compact WARP
//    TILE, 32 consecutive columns a COLUMN TILE with a stored bounding box.
//  * A warp owns a warp tile (rows in registers).  Column tiles are first tested box against box
//    (32 tile headers per ballot, FP64); a tile farther from the warp's box than the largest
//    possible reach is skipped without being read.
//  * Broad phase per pair in SINGLE precision (geom.cuh: bound32_far, 14 FP32 instructions, one
//    16-byte warp-uniform load per column): capsule bound dist(seg_r, seg_c) >= dist(m_c, seg_r)
//    - h_c against r_i + r_j + 2^-20 with a one-sided error budget, so a rejected pair is a pair
//    the reference arithmetic calls "no collision".  The FP64 pipe is left to the pairs that
//    matter.
//  * Survivors are compacted into a per-warp shared-memory queue (popc + warp prefix) and
//    evaluated on dense lanes by the verdict filter (geom.cuh: segseg_filter, FP64 with FMA and
//    per-segment reciprocals, ~100 instructions instead of ~380): it decides every pair outside a
//    2^-20 band around contact; the rest (parallel segments, contact within the band, coordinates
//    outside [-2, 2]) runs the reference's own sequence of roundings (segseg_lt), so every verdict
//    is the reference's bit for bit.  No block-wide barrier: warps are autonomous and pull warp
//    tiles from an atomic counter.
// Output modes: full bit matrix; OR over column groups (e.g. all agents' path edges of one time
// step); COUNT per (row, group) (CBS soft conflicts); FIRST = smallest column index per
// (row, group) (TPG dependency scan), optionally restricted to col_key > row_key.
#include <stdlib.h>

#include "point_model.cuh"

namespace mrmp {

extern __shared__ __align__(16) unsigned char smem_raw[];

constexpr int kCols = 32;        // column edges per tile (must equal kCrossCols)
constexpr int kRedWords = 8;     // OR mode: up to 256 groups per launch
constexpr int kWarpsPerCta = 8;
constexpr int kSortCells = 4096;  // Morton grid: 64 x 64 (2-D), 16^3 (3-D)

template <int D>
struct alignas(16) RowRec {  // one row edge, in Morton order (16-byte aligned: the scatter of the
                             // sort writes a record with 128-bit stores)
    double a[D], b[D];
    int agent, orig;  // orig = row index in the caller's order
};

struct ColB {  // broad-phase record of a column (16 bytes, one warp-uniform load)
    float m[3];  // midpoint (z = 0 in 2-D)
    float h;     // half length + agent radius, rounded up; +inf outside the FP32 domain
};

template <int D>
struct ColX {  // exact-stage record of a column (64 bytes in 2-D, 80 in 3-D)
    double a[D], b[D];  // q_from, q_to
    double inv;         // geom.cuh: seg_filter_inv
    double rad;         // rads[agent]
    int agent;          // -1 = padding
    int bit;            // output slot: column index (bits), group (or / count / first)
    int orig;           // column index in the caller's order (first mode)
    int key;            // col_key (0 without keys)
};

// Bounding box of the tile's midpoints and its largest reach as order-preserving integer keys of
// floats rounded OUTWARD (a box that is 1e-7 too large only culls that much less): they can be
// reduced with integer atomicMin / atomicMax by the threads that write the columns.
template <int D>
struct TileHdr {
    int lo[D], hi[D];
    int hmax;
    int poison;  // a NaN midpoint / reach: the tile must never be culled by the box test
};

__device__ __forceinline__ int fkey(float f) {  // order-preserving float -> int
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float fkey_inv(int k) {
    return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff);
}

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

// ---- column tiles: one warp builds one tile (one column per lane) and its header -----------
template <int D>
struct ColPrep {
    EdgeList<D> el;
    int64_t n_cols;
    const int32_t *order;      // Morton order -> source index
    int group_size;            // > 0: bit = src / group_size (unless col_group)
    const int32_t *col_group;  // explicit group ids (may be NULL)
    const int32_t *col_key;    // may be NULL
    int64_t orig_base;         // added to src for ColX::orig
    ColB *cb;
    ColX<D> *cx;
    TileHdr<D> *hdrs;
    int64_t n_tiles;
    const double *rads;
    int n_agents;
};

// broad-phase and exact-stage records of one column; m / h = its midpoint and reach in FP64
// (inputs of the tile header)
template <int D>
__device__ __forceinline__ void make_col(const ColPrep<D> &P, int64_t src, int j, const double *a,
                                         const double *b, ColB &B, ColX<D> &X, double *m,
                                         double &h) {
    const bool known = j >= 0 && j < P.n_agents;
    // reach of the column: half length + its agent's radius (an id outside the ctx is rejected
    // by the exact stage; it must not be culled here)
    const double rj = known ? __ldg(P.rads + j) : 1e300;
    h = half_len_up<D>(a, b) + rj;
    X.inv = seg_filter_inv<D>(a, b);
    X.rad = known ? rj : 0.0;
    X.agent = known ? j : -1;
    X.bit = (int)(P.col_group ? __ldg(P.col_group + src)
                              : (P.group_size > 0 ? src / P.group_size : src));
    X.orig = (int)(src + P.orig_base);
    X.key = P.col_key ? __ldg(P.col_key + src) : 0;
    // 2-D: the z slot carries the output slot for the kernel's in-loop "already set" test
    B.m[2] = D == 2 ? __int_as_float(X.bit) : 0.0f;
#pragma unroll
    for (int x = 0; x < D; ++x) {
        X.a[x] = a[x];
        X.b[x] = b[x];
        m[x] = dmul(dadd(a[x], b[x]), 0.5);
        B.m[x] = (float)m[x];
    }
    B.h = (X.inv >= 0.0 && known) ? f32_up(h) : INFINITY;
}

template <int D>
__device__ __forceinline__ void make_pad(ColB &B, ColX<D> &X) {  // far away, never selected
    B.m[0] = B.m[1] = 1e15f;
    B.m[2] = 0.0f;
    B.h = 0.0f;
#pragma unroll
    for (int x = 0; x < D; ++x) X.a[x] = X.b[x] = 0.0;
    X.inv = -1.0;
    X.rad = 0.0;
    X.agent = -1;
    X.bit = X.orig = X.key = 0;
}

// header of tile t from this lane's column (m, h; `real` false for padding): warp reduction
template <int D>
__device__ __forceinline__ void tile_header(const ColPrep<D> &P, int64_t t, int lane, bool real,
                                            const double *m, double h) {
    double lo[D], hi[D], hmax = real ? h : 0.0;
    bool poison = real && !(h == h);  // a NaN midpoint must never be culled by the box test
#pragma unroll
    for (int x = 0; x < D; ++x) {
        lo[x] = real ? m[x] : 1e300;
        hi[x] = real ? m[x] : -1e300;
        poison |= real && !(m[x] == m[x]);
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
            H.lo[x] = fkey(__double2float_rd(lo[x]));
            H.hi[x] = fkey(__double2float_ru(hi[x]));
        }
        H.hmax = fkey(__double2float_ru(hmax));
        H.poison = poison ? 1 : 0;
        P.hdrs[t] = H;
    }
}

// one warp builds tile t (one column per lane) and its header
template <int D>
__device__ __forceinline__ void prep_tile(const ColPrep<D> &P, int64_t t, int lane) {
    const int64_t c = t * kCols + lane;
    ColB B;
    ColX<D> X;
    double m[D], h = 0.0;
    const bool real = c < P.n_cols;
    if (real) {
        const int64_t src = P.order[c];
        double a[D], b[D];
        const int j = P.el.get(src, a, b);
        make_col<D>(P, src, j, a, b, B, X, m, h);
    } else {
        make_pad<D>(B, X);
#pragma unroll
        for (int x = 0; x < D; ++x) m[x] = 0.0;
    }
    P.cb[c] = B;
    P.cx[c] = X;
    tile_header<D>(P, t, lane, real, m, h);
}

template <int D>
__global__ void __launch_bounds__(256) k_cross_prep_cols(ColPrep<D> P) {
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < P.n_tiles;
         t += n_warps)
        prep_tile<D>(P, t, lane);
}

// The column side for the usual case of a few thousand columns (a conflict table has N x T of
// them) in TWO launches instead of four: `place` -- one CTA -- counts the Morton cells in shared
// memory, scans, and gives every column its sorted position (and resets the tile headers and the
// work counter of the table kernel); `fill` -- a thread per column over many SMs -- builds the
// records (a division and a square root each: too much for one SM) at their position and reduces
// the tile headers with integer atomics.
constexpr int kPrepSmallCpt = 4;
constexpr int kPrepSmallMax = 1024 * kPrepSmallCpt;

template <int D>
__global__ void __launch_bounds__(1024)
k_cross_cols_place(ColPrep<D> P, int32_t *__restrict__ place, unsigned int *__restrict__ work) {
    __shared__ int h[kSortCells];
    __shared__ int part[32];
    const int tid = threadIdx.x;
    for (int c = tid; c < kSortCells; c += 1024) h[c] = 0;
    for (int64_t t = tid; t < P.n_tiles; t += 1024) {
        TileHdr<D> H;
#pragma unroll
        for (int x = 0; x < D; ++x) {
            H.lo[x] = fkey(INFINITY);
            H.hi[x] = fkey(-INFINITY);
        }
        H.hmax = fkey(0.0f);
        H.poison = 0;
        P.hdrs[t] = H;
    }
    if (tid == 0 && work) *work = 0u;
    __syncthreads();
    int cell[kPrepSmallCpt];
#pragma unroll
    for (int i = 0; i < kPrepSmallCpt; ++i) {
        const int64_t k = tid + 1024 * i;
        cell[i] = -1;
        if (k < P.n_cols) {
            double a[D], b[D];
            P.el.get(k, a, b);
            cell[i] = morton_cell<D>(a, b);
            atomicAdd(&h[cell[i]], 1);
        }
    }
    __syncthreads();
    int v[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[k] = h[4 * tid + k];
        sum += v[k];
    }
    // exclusive scan of the 1024 partial sums: warp shuffles + one pass over the 32 warp totals
    const int lane = tid & 31, warp = tid >> 5;
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int tv = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += tv;
    }
    if (lane == 31) part[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int wt = part[lane];
        int wi = wt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int tv = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += tv;
        }
        part[lane] = wi - wt;
    }
    __syncthreads();
    int base = part[warp] + incl - sum;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        h[4 * tid + k] = base;
        base += v[k];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kPrepSmallCpt; ++i)
        if (cell[i] >= 0) place[tid + 1024 * i] = atomicAdd(&h[cell[i]], 1);
}

template <int D>
__global__ void __launch_bounds__(128)
k_cross_cols_fill(ColPrep<D> P, const int32_t *__restrict__ place) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= P.n_tiles * kCols) return;
    ColB B;
    ColX<D> X;
    if (k >= P.n_cols) {  // the padding slots behind the last column
        make_pad<D>(B, X);
        P.cb[k] = B;
        P.cx[k] = X;
        return;
    }
    double a[D], b[D], m[D], hh;
    const int j = P.el.get(k, a, b);
    make_col<D>(P, k, j, a, b, B, X, m, hh);
    const int pos = place[k];
    P.cb[pos] = B;
    P.cx[pos] = X;
    TileHdr<D> &H = P.hdrs[pos / kCols];
    bool bad = !(hh == hh);
#pragma unroll
    for (int x = 0; x < D; ++x) {
        bad |= !(m[x] == m[x]);
        atomicMin(&H.lo[x], fkey(__double2float_rd(m[x])));
        atomicMax(&H.hi[x], fkey(__double2float_ru(m[x])));
    }
    atomicMax(&H.hmax, fkey(__double2float_ru(hh)));
    if (bad) H.poison = 1;
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

// ---- CSR of a roadmap set -> Morton-sorted row records in two launches (point models) ------------
// kCsrLanes lanes per CSR row (a PRM vertex has about 8 neighbours at the BASELINE density; a full
// warp per row left 3/4 of the lanes idle and the row loop serial), lane per edge: count the cells,
// then -- with the exclusive scan of the 4096 counts redone by every CTA in shared memory instead
// of a launch of its own -- scatter the records.
constexpr int kCsrLanes = 8;

template <int D>
__global__ void __launch_bounds__(256)
k_csr_cell_hist(int64_t n_rows_csr, int nv, const int32_t *__restrict__ row_ptr,
                const int32_t *__restrict__ col_idx, const double *__restrict__ nodes,
                int32_t *__restrict__ counts, int32_t *__restrict__ cursor,
                unsigned int *__restrict__ done, int64_t e_cap) {
    __shared__ int h[kSortCells];
    __shared__ int part[8];
    __shared__ bool s_last;
    for (int c = threadIdx.x; c < kSortCells; c += blockDim.x) h[c] = 0;
    __syncthreads();
    const int lane = threadIdx.x & (kCsrLanes - 1);
    const int64_t n_sub = ((int64_t)gridDim.x * blockDim.x) / kCsrLanes;
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kCsrLanes; row < n_rows_csr;
         row += n_sub) {
        const int a = (int)(row / nv), v = (int)(row - (int64_t)a * nv);
        const double *tab = nodes + (int64_t)a * nv * D;
        double qv[D];
        load_state<D>(tab, v, qv);
        // (e_cap: capacity of col_idx / the row records; only an overflowing asynchronous build
        // ever reaches it, and that build is rejected by mrmp_roadmaps_finish)
        for (int e = row_ptr[row] + lane; e < row_ptr[row + 1] && e < e_cap; e += kCsrLanes) {
            double qk[D];
            load_state<D>(tab, col_idx[e], qk);
            atomicAdd(&h[morton_cell<D>(qv, qk)], 1);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < kSortCells; c += blockDim.x)
        if (h[c]) atomicAdd(counts + c, h[c]);
    // the CTA that finishes last turns the counts into the cursors of the scatter (exclusive
    // scan, 16 cells per thread): no scan launch, and no scan in every CTA of the scatter
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int tid = threadIdx.x;
    int v[16], sum = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        v[k] = __ldcg(counts + 16 * tid + k);
        sum += v[k];
    }
    const int ln = tid & 31, wp = tid >> 5;
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int tv = __shfl_up_sync(0xffffffffu, incl, o);
        if (ln >= o) incl += tv;
    }
    if (ln == 31) part[wp] = incl;
    __syncthreads();
    int base = incl - sum;
    for (int w = 0; w < wp; ++w) base += part[w];  // 8 warps
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        cursor[16 * tid + k] = base;
        base += v[k];
    }
}

template <int D>
__global__ void __launch_bounds__(256)
k_csr_rows_scatter(int64_t n_rows_csr, int nv, int agent0, const int32_t *__restrict__ row_ptr,
                   const int32_t *__restrict__ col_idx, const double *__restrict__ nodes,
                   int32_t *__restrict__ cursor, RowRec<D> *__restrict__ rows, int64_t e_cap) {
    const int lane = threadIdx.x & (kCsrLanes - 1);
    const int64_t n_sub = ((int64_t)gridDim.x * blockDim.x) / kCsrLanes;
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kCsrLanes; row < n_rows_csr;
         row += n_sub) {
        const int a = (int)(row / nv), v = (int)(row - (int64_t)a * nv);
        const double *tab = nodes + (int64_t)a * nv * D;
        RowRec<D> r;
        load_state<D>(tab, v, r.a);
        r.agent = agent0 + a;
        for (int e = row_ptr[row] + lane; e < row_ptr[row + 1] && e < e_cap; e += kCsrLanes) {
            load_state<D>(tab, col_idx[e], r.b);
            r.orig = e;
            rows[atomicAdd(cursor + morton_cell<D>(r.a, r.b), 1)] = r;
        }
    }
}

// ---- the cross kernel -------------------------------------------------------------------------
enum { kModeBits = kCrossBits, kModeOr = kCrossOr, kModeCount = kCrossCount, kModeFirst = kCrossFirst };

template <int D, int R>
struct WarpSmem {
    double rab[R][2 * D];  // q_from, q_to of the warp tile's rows
    double inv[R];         // seg_filter_inv
    double rad[R];         // rads[agent]
    int agent[R];
    int orig[R];
    int key[R];
    uint32_t obits[R][kRedWords];
    uint16_t queue[R * kCols];
};

template <int D>
struct CrossArgs {
    int64_t n_rows;
    const RowRec<D> *rows;
    int n_tiles;
    const ColB *cb;
    const ColX<D> *cx;
    const TileHdr<D> *hdrs;
    int words_per_row;   // or mode: words written per row
    int64_t out_stride;  // 32-bit words (bits / or) or ints (count / first) per row
    uint32_t *out;
    const int32_t *row_key;  // NULL: every pair is admitted; else col_key > row_key[row]
    int cbs;                 // cbs.jl:343-346: vertex test OR edge test with the column first
    int split;               // work items per warp tile (column tiles divided into `split` ranges)
    const long long *n_rows_dev;  // NULL, or the actual row count (<= n_rows) on the device
    int64_t out_rows;             // rows `out` has room for: a row whose index is beyond is dropped
    unsigned long long *stats;
    unsigned int *work;
};

#ifndef MRMP_CROSS_MINB
#define MRMP_CROSS_MINB 3
#endif
#ifndef MRMP_CROSS_UNROLL
#define MRMP_CROSS_UNROLL 32   // broad-phase loop over the 32 columns of a tile
#endif
constexpr int kBroadUnroll = MRMP_CROSS_UNROLL;

// The reference's own sequence of roundings for one pair (utils.jl:59-94): rare -- the filter
// decides all but the pairs within 2^-20 of contact, near-parallel pairs and coordinates outside
// [-2, 2].
template <int D>
__device__ __forceinline__ bool cross_exact(const double *p11, const double *p12,
                                            const double *p21, const double *p22, double thr,
                                            double t2, int has_tab) {
    return has_tab ? segseg_lt<D, true>(p11, p12, p21, p22, thr, t2)
                   : segseg_lt<D, false>(p11, p12, p21, p22, thr, 0.0);
}

// verdict of one (row, column) pair: filter first, reference arithmetic when undecided
template <int D>
__device__ __forceinline__ bool cross_pair_verdict(const CtxView &cv, const double *ra,
                                                   const double *rb, double rinv, double rrad,
                                                   int ri, const double *ca, const double *cb,
                                                   double cinv, double crad, int cj, int cbs,
                                                   bool &ran_exact) {
    const double *t2tab = cv.blob + cv.t2_pair_off;
    const double thr = dadd(rrad, crad);  // rads[i] + rads[j], point.jl:11 (commutative)
    ran_exact = false;
    if (cbs) {
        // collide(q_j_to, q_i_to, j, i)  cbs.jl:345 / point.jl:14-16
        double v[D];
#pragma unroll
        for (int x = 0; x < D; ++x) v[x] = dsub(cb[x], rb[x]);
        const double s = sumsq<D>(v);
        const bool vertex = cv.has_pair_tab
                                ? norm_lt<D, true>(v, s, thr, __ldg(t2tab + cj * cv.n_agents + ri))
                                : norm_lt<D, false>(v, s, thr, 0.0);
        if (vertex) return true;
        // collide(q_j_from, q_j_to, q_i_from, q_i_to, j, i)  cbs.jl:346: the column is segment 1
        const int f = segseg_filter<D>(ca, cb, ra, rb, cinv, rinv, thr);
        if (f >= 0) return f != 0;
        ran_exact = true;
        return cross_exact<D>(ca, cb, ra, rb, thr,
                              cv.has_pair_tab ? __ldg(t2tab + cj * cv.n_agents + ri) : 0.0,
                              cv.has_pair_tab);
    }
    const int f = segseg_filter<D>(ra, rb, ca, cb, rinv, cinv, thr);
    if (f >= 0) return f != 0;
    ran_exact = true;
    return cross_exact<D>(ra, rb, ca, cb, thr,
                          cv.has_pair_tab ? __ldg(t2tab + ri * cv.n_agents + cj) : 0.0,
                          cv.has_pair_tab);
}

// FD ("fast drop", 2-D OR mode with at most 64 groups): the group id of a column travels in the
// unused z slot of its broad-phase record and the row's output words live in registers, so the
// "bit already set" test is part of the broad phase instead of a loop over the survivors.
template <int D, int MODE, bool FD>
__global__ void __launch_bounds__(32 * kWarpsPerCta, MRMP_CROSS_MINB)
k_cross_collide(CtxView cv, CrossArgs<D> A) {
    constexpr int R = 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpSmem<D, R> &S = reinterpret_cast<WarpSmem<D, R> *>(smem_raw)[warp];
    const double *rads = cv.blob + cv.crads_off;
    int64_t n_rows_k = A.n_rows;
    if (A.n_rows_dev) {  // launched behind an asynchronous build: the count is only known here
        const long long nd = *A.n_rows_dev;
        if (nd < n_rows_k) n_rows_k = nd;
    }
    const int64_t n_wt = (n_rows_k + R - 1) / R;
    const int64_t n_items = n_wt * A.split;
    unsigned long long n_tiles_read = 0, n_filter = 0, n_exact = 0;  // per-warp / per-lane counters

    for (;;) {
        // work item = (warp tile, range of column tiles).  A large table has several warp tiles
        // per resident warp and split == 1; a small one (a rank's share of a strong-scaled
        // table) is cut finer so that the dynamic distribution can balance the SMs.
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(A.work, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if ((int64_t)item >= n_items) break;
        // (column tiles are dealt to the items round robin: the tiles near a warp tile are
        // neighbours in Morton order, a contiguous range would give one item all the work)
        const unsigned wt = item / (unsigned)A.split, chunk = item - wt * (unsigned)A.split;
        RowB32 rb;
        double lo[D], hi[D], Rmax = 0.0;
        bool poison = false;
#pragma unroll
        for (int x = 0; x < D; ++x) {
            lo[x] = 1e300;
            hi[x] = -1e300;
        }
        const int64_t r = (int64_t)wt * R + lane;
        // (orig beyond the output's room: only behind an asynchronous build whose edge count
        // exceeded the room the caller gave -- rejected afterwards by mrmp_roadmaps_finish)
        const bool live = r < n_rows_k && A.rows[r].orig < A.out_rows;
        if (live) {
            const RowRec<D> rec = A.rows[r];
            const bool known = rec.agent >= 0 && rec.agent < cv.n_agents;
            const double ri = known ? __ldg(rads + rec.agent) : 1e300;
            const double inv = seg_filter_inv<D>(rec.a, rec.b);
            // reach of this row: its radius + slack (+ its own half length in the box test);
            // the column's radius and half length are in its record
            const double Rc = ri + kPad;
            rb = bound32_row<D>(rec.a, rec.b, Rc, inv >= 0.0 && known);
            Rmax = Rc + half_len_up<D>(rec.a, rec.b);
            poison |= !(Rmax == Rmax);
#pragma unroll
            for (int x = 0; x < D; ++x) {
                const double m = dmul(dadd(rec.a[x], rec.b[x]), 0.5);
                S.rab[lane][x] = rec.a[x];
                S.rab[lane][D + x] = rec.b[x];
                lo[x] = hi[x] = m;
                poison |= !(m == m);
            }
            S.inv[lane] = inv;
            S.rad[lane] = known ? ri : 0.0;
            S.agent[lane] = known ? rec.agent : -2;
            S.orig[lane] = rec.orig;
            S.key[lane] = A.row_key ? __ldg(A.row_key + rec.orig) : 0;
        } else {
#pragma unroll
            for (int x = 0; x < 3; ++x) rb.m[x] = rb.u[x] = 0.0f;
            rb.h = rb.rc = 0.0f;
            S.agent[lane] = -2;
            S.orig[lane] = 0;
        }
        if (MODE == kModeOr) {
#pragma unroll
            for (int w = 0; w < kRedWords; ++w) S.obits[lane][w] = 0u;
        }
        uint32_t ob0 = 0u, ob1 = 0u;  // FD: this row's output words
        // warp box and largest reach; a NaN row disables the box test for the whole warp tile
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

        for (int t0 = 0; t0 < A.n_tiles; t0 += 32) {
            // box-vs-box test of 32 column tiles at once
            bool near = false;
            if (t0 + lane < A.n_tiles &&
                (A.split == 1 || (unsigned)(t0 + lane) % (unsigned)A.split == chunk)) {
                const TileHdr<D> H = A.hdrs[t0 + lane];
                double d2 = 0.0;
#pragma unroll
                for (int x = 0; x < D; ++x) {
                    const double hlo = (double)fkey_inv(H.lo[x]), hhi = (double)fkey_inv(H.hi[x]);
                    const double gap = fmax(fmax(hlo - hi[x], lo[x] - hhi), 0.0);
                    d2 += gap * gap;
                }
                const double reach = Rmax + (double)fkey_inv(H.hmax);  // +inf: never culled
                near = poison || H.poison != 0 || !(d2 > reach * reach);
            }
            unsigned tmask = __ballot_sync(0xffffffffu, near);
            while (tmask) {
                const int t = t0 + __ffs(tmask) - 1;
                tmask &= tmask - 1;
                const float4 *cbp = reinterpret_cast<const float4 *>(A.cb + (int64_t)t * kCols);
                const ColX<D> *cxp = A.cx + (int64_t)t * kCols;
                // ---- broad phase (FP32): this lane's row against the 32 columns of the tile
                uint32_t mask = 0u;
#pragma unroll kBroadUnroll
                for (int c = 0; c < kCols; ++c) {
                    const float4 q = __ldg(cbp + c);
                    bool keep = !bound32_far<D>(rb, q.x, q.y, FD ? 0.0f : q.z, q.w);
                    if (FD) {
                        // the output is an OR over each column group: a pair whose bit this row
                        // has already set (by an earlier tile) cannot change it
                        const int bit = __float_as_int(q.z);  // warp-uniform
                        keep &= !(((bit & 32 ? ob1 : ob0) >> (bit & 31)) & 1u);
                    }
                    mask |= (keep ? 1u : 0u) << c;
                }
                ++n_tiles_read;
                if (!live) mask = 0u;
                if (MODE == kModeOr && !FD) {
                    // same drop for the general layout (3-D, more than 64 groups): walk the
                    // survivors, before the compaction, so that the exact stage stays dense
                    uint32_t m2 = mask;
                    while (m2) {
                        const int c = __ffs(m2) - 1;
                        m2 &= m2 - 1;
                        const int bit = __ldg(&cxp[c].bit);
                        if ((S.obits[lane][bit >> 5] >> (bit & 31)) & 1u) mask &= ~(1u << c);
                    }
                }
                // ---- compaction into the warp queue
                const int cnt = __popc(mask);
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
                    const int c = __ffs(mask) - 1;
                    mask &= mask - 1;
                    S.queue[base++] = (uint16_t)((lane << 5) | c);
                }
                __syncwarp();
                // ---- verdict filter / reference arithmetic on dense lanes
                for (int e = lane; e < total; e += 32) {
                    const int ent = S.queue[e];
                    const int rr = ent >> 5, cc = ent & 31;
                    const ColX<D> &X = cxp[cc];
                    const int4 ids = __ldg(reinterpret_cast<const int4 *>(&X.agent));
                    const int ri = S.agent[rr], cj = ids.x;
                    // padding / unknown agent / same agent: never a conflict
                    if (cj < 0 || ri < 0 || cj == ri) continue;
                    if (A.row_key && !(ids.w > S.key[rr])) continue;
                    double ca[D], cb[D];
                    if constexpr (D == 2) {
                        const double2 a2 = __ldg(reinterpret_cast<const double2 *>(&X.a[0]));
                        const double2 b2 = __ldg(reinterpret_cast<const double2 *>(&X.b[0]));
                        ca[0] = a2.x, ca[1] = a2.y, cb[0] = b2.x, cb[1] = b2.y;
                    } else {
#pragma unroll
                        for (int x = 0; x < D; ++x) {
                            ca[x] = __ldg(&X.a[x]);
                            cb[x] = __ldg(&X.b[x]);
                        }
                    }
                    const double2 ir = __ldg(reinterpret_cast<const double2 *>(&X.inv));
                    bool ran_exact;
                    const bool hit = cross_pair_verdict<D>(cv, &S.rab[rr][0], &S.rab[rr][D],
                                                           S.inv[rr], S.rad[rr], ri, ca, cb, ir.x,
                                                           ir.y, cj, A.cbs, ran_exact);
                    ++n_filter;
                    n_exact += ran_exact ? 1 : 0;
                    if (hit) {
                        const int bit = ids.y;
                        if (MODE == kModeOr)
                            atomicOr(&S.obits[rr][bit >> 5], 1u << (bit & 31));
                        else if (MODE == kModeBits)
                            atomicOr(A.out + (int64_t)S.orig[rr] * A.out_stride + (bit >> 5),
                                     1u << (bit & 31));
                        else if (MODE == kModeCount)
                            atomicAdd(reinterpret_cast<int *>(A.out) +
                                          (int64_t)S.orig[rr] * A.out_stride + bit, 1);
                        else
                            atomicMin(reinterpret_cast<int *>(A.out) +
                                          (int64_t)S.orig[rr] * A.out_stride + bit, ids.z);
                    }
                }
                __syncwarp();
                if (FD) {
                    ob0 = S.obits[lane][0];
                    ob1 = S.obits[lane][1];
                }
            }
        }
        if (MODE == kModeOr && live && A.split > 1) {
            // several work items share the row: OR into the (zeroed, device-local) output
            uint32_t *dst = A.out + (int64_t)S.orig[lane] * A.out_stride;
            for (int w = 0; w < A.words_per_row; ++w)
                if (S.obits[lane][w]) atomicOr(dst + w, S.obits[lane][w]);
        } else if (MODE == kModeOr && live) {
            // every word of the row is written exactly once (the output may be mapped host or
            // peer memory: plain stores, two words per transaction where the layout allows)
            uint32_t *dst = A.out + (int64_t)S.orig[lane] * A.out_stride;
            if ((A.words_per_row & 1) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
                for (int w = 0; w < A.words_per_row; w += 2)
                    *reinterpret_cast<uint2 *>(dst + w) =
                        make_uint2(S.obits[lane][w], S.obits[lane][w + 1]);
            } else {
                for (int w = 0; w < A.words_per_row; ++w) dst[w] = S.obits[lane][w];
            }
        }
        __syncwarp();
    }
    if (A.stats) {
        // [0] pairs that ran the reference arithmetic, [2] pairs bound-tested, [3] pairs that ran
        // the verdict filter
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_filter += __shfl_xor_sync(0xffffffffu, n_filter, o);
            n_exact += __shfl_xor_sync(0xffffffffu, n_exact, o);
        }
        if (lane == 0) {
            if (n_exact) atomicAdd(A.stats, n_exact);
            if (n_tiles_read) atomicAdd(A.stats + 2, n_tiles_read * (unsigned long long)(R * kCols));
            if (n_filter) atomicAdd(A.stats + 3, n_filter);
        }
    }
}

// first mode: rows without a hit hold INT_MAX after the kernel -> -1
__global__ void __launch_bounds__(256) k_first_finalize(int32_t *__restrict__ out, int64_t n) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
         k += (int64_t)gridDim.x * blockDim.x)
        if (out[k] == 0x7fffffff) out[k] = -1;
}
__global__ void __launch_bounds__(256) k_fill_i32(int32_t *__restrict__ out, int64_t n, int32_t v) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
         k += (int64_t)gridDim.x * blockDim.x)
        out[k] = v;
}

// ---- launchers --------------------------------------------------------------------------------
size_t cross_tile_bytes(int d) {
    return kCols * (sizeof(ColB) + (d == 2 ? sizeof(ColX<2>) : sizeof(ColX<3>)));
}
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
                               int64_t n_cols, const CrossCols &cc, int32_t *cells,
                               int32_t *order, void *tiles, void *hdrs) {
    const int64_t n_tiles = (n_cols + kCols - 1) / kCols;
    ColPrep<D> P;
    P.el = el;
    P.n_cols = n_cols;
    P.order = order;
    P.group_size = cc.group_size;
    P.col_group = cc.col_group;
    P.col_key = cc.col_key;
    P.orig_base = cc.orig_base;
    P.cb = (ColB *)tiles;
    P.cx = (ColX<D> *)((unsigned char *)tiles + sizeof(ColB) * kCols * (size_t)n_tiles);
    P.hdrs = (TileHdr<D> *)hdrs;
    P.n_tiles = n_tiles;
    P.rads = cv.blob + cv.crads_off;
    P.n_agents = cv.n_agents;
    if (n_cols <= kPrepSmallMax) {
        k_cross_cols_place<D><<<1, 1024, 0, lc.stream>>>(P, order, nullptr);
        const int64_t slots = n_tiles * kCols;
        k_cross_cols_fill<D><<<(int)((slots + 127) / 128), 128, 0, lc.stream>>>(P, order);
        count_launch(2);
        return cudaGetLastError();
    }
    cudaError_t e = cudaMemsetAsync(cells, 0, sizeof(int32_t) * kSortCells, lc.stream);
    if (e != cudaSuccess) return e;
    const int grid = grid_for_n(lc, n_cols, 256 * 4);
    k_cell_hist<D><<<grid, 256, 0, lc.stream>>>(el, n_cols, cells);
    k_cell_scan<<<1, 1024, 0, lc.stream>>>(cells);
    k_cols_scatter<D><<<grid, 256, 0, lc.stream>>>(el, n_cols, cells, order);
    k_cross_prep_cols<D><<<grid_for_n(lc, n_tiles * 32, 256), 256, 0, lc.stream>>>(P);
    count_launch(4);
    return cudaGetLastError();
}

// columns (explicit states or ids into a node table) -> Morton-sorted tiles + headers
cudaError_t launch_cross_prep_cols(const CtxView &cv, const LaunchCfg &lc, int64_t n_cols,
                                   const int32_t *agent, const double *qf, const double *qt,
                                   const int32_t *vf, const int32_t *vt, const double *nodes,
                                   int agent0, int nv_stride, const CrossCols &cc, int32_t *cells,
                                   int32_t *order, void *tiles, void *hdrs) {
    if (n_cols <= 0) return cudaSuccess;
    if (n_cols >= ((int64_t)1 << 31)) return cudaErrorInvalidValue;
    if (cv.model == 0)
        return prep_cols_t<2>(cv, lc, EdgeList<2>{agent, qf, qt, vf, vt, nodes, agent0, nv_stride},
                              n_cols, cc, cells, order, tiles, hdrs);
    if (cv.model == 1)
        return prep_cols_t<3>(cv, lc, EdgeList<3>{agent, qf, qt, vf, vt, nodes, agent0, nv_stride},
                              n_cols, cc, cells, order, tiles, hdrs);
    return cudaErrorNotSupported;
}

// CSR edges of a point-model roadmap set -> Morton-sorted RowRec array (row e <-> col_idx[e]);
// cells = 2 * 4096 + 1 ints of device scratch (counts | cursors | CTAs done)
cudaError_t launch_cross_sort_csr_rows(const CtxView &cv, const LaunchCfg &lc, int64_t n_rows_csr,
                                       int nv, int agent0, const int32_t *row_ptr,
                                       const int32_t *col_idx, const double *nodes, int32_t *cells,
                                       void *rows_out, int64_t e_cap) {
    if (n_rows_csr <= 0) return cudaSuccess;
    if (cv.model > 1) return cudaErrorNotSupported;
    cudaError_t e = cudaMemsetAsync(cells, 0, sizeof(int32_t) * (2 * kSortCells + 1), lc.stream);
    if (e != cudaSuccess) return e;
    int64_t want = (n_rows_csr * kCsrLanes + 255) / 256;
    const int64_t cap = (int64_t)lc.sm_count * 8;
    const int grid = (int)(want < cap ? want : cap);
    int32_t *cursor = cells + kSortCells;
    unsigned int *done = reinterpret_cast<unsigned int *>(cells + 2 * kSortCells);
    if (cv.model == 0) {
        k_csr_cell_hist<2><<<grid, 256, 0, lc.stream>>>(n_rows_csr, nv, row_ptr, col_idx, nodes, cells,
                                                       cursor, done, e_cap);
        k_csr_rows_scatter<2><<<grid, 256, 0, lc.stream>>>(n_rows_csr, nv, agent0, row_ptr, col_idx,
                                                          nodes, cursor, (RowRec<2> *)rows_out,
                                                          e_cap);
    } else {
        k_csr_cell_hist<3><<<grid, 256, 0, lc.stream>>>(n_rows_csr, nv, row_ptr, col_idx, nodes, cells,
                                                       cursor, done, e_cap);
        k_csr_rows_scatter<3><<<grid, 256, 0, lc.stream>>>(n_rows_csr, nv, agent0, row_ptr, col_idx,
                                                          nodes, cursor, (RowRec<3> *)rows_out,
                                                          e_cap);
    }
    count_launch(2);
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

template <int D, int MODE, bool FD>
static cudaError_t launch_cross_v(const CtxView &cv, const LaunchCfg &lc, const CrossArgs<D> &A) {
    constexpr int R = 32;
    const int smem = (int)sizeof(WarpSmem<D, R>) * kWarpsPerCta;
    int per_sm = (220 * 1024) / smem;
    if (per_sm > MRMP_CROSS_MINB) per_sm = MRMP_CROSS_MINB;  // register-limited residency
    if (per_sm < 1) per_sm = 1;
    const int64_t n_wt = (A.n_rows + R - 1) / R * A.split;
    int64_t want = (n_wt + kWarpsPerCta - 1) / kWarpsPerCta;
    const int64_t cap = (int64_t)lc.sm_count * per_sm;
    const int grid = (int)(want < cap ? want : cap);
    cudaError_t e = cudaFuncSetAttribute(k_cross_collide<D, MODE, FD>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    k_cross_collide<D, MODE, FD><<<grid, 32 * kWarpsPerCta, smem, lc.stream>>>(cv, A);
    count_launch();
    return cudaGetLastError();
}

// Work items per warp tile.  A warp tile (32 rows x all column tiles) is ~25 k dependent-latency
// bound instructions, 60-140 us of one warp's time, and every finished item can drop the pairs
// whose output bit it has already set -- so items are only cut finer (column tiles dealt round
// robin) when there is less than one warp tile per resident warp, i.e. when SMs would idle
// (measured on B200, profiles/r2_shard_split.md: a rank's 1/8 share gains 10 %, larger shares lose).
int cross_split(const LaunchCfg &lc, int64_t n_rows, int n_tiles, int may_split) {
    static int forced = -1;  // MRMP_CROSS_SPLIT=<n> pins the factor (experiments)
    if (forced < 0) {
        const char *e = getenv("MRMP_CROSS_SPLIT");
        forced = e ? atoi(e) : 0;
    }
    if (!may_split || n_tiles <= 1) return 1;
    if (forced > 0) return forced < n_tiles ? forced : n_tiles;
    const int64_t n_wt = (n_rows + 31) / 32;
    const int64_t warps = (int64_t)lc.sm_count * MRMP_CROSS_MINB * kWarpsPerCta;
    if (n_wt >= warps) return 1;
    int64_t s = (2 * warps + n_wt - 1) / n_wt;
    if (s > 4) s = 4;
    if (s > n_tiles) s = n_tiles;
    return (int)(s < 1 ? 1 : s);
}

template <int D>
static cudaError_t launch_cross_t(const CtxView &cv, const LaunchCfg &lc, int64_t n_rows,
                                  const void *rows, int64_t n_cols, const void *tiles,
                                  const void *hdrs, const CrossOut &co, unsigned long long *stats,
                                  unsigned int *work) {
    const int n_tiles = (int)((n_cols + kCols - 1) / kCols);
    CrossArgs<D> A;
    A.n_rows = n_rows;
    A.rows = (const RowRec<D> *)rows;
    A.n_tiles = n_tiles;
    A.cb = (const ColB *)tiles;
    A.cx = (const ColX<D> *)((const unsigned char *)tiles + sizeof(ColB) * kCols * (size_t)n_tiles);
    A.hdrs = (const TileHdr<D> *)hdrs;
    A.words_per_row = co.words_per_row;
    A.out_stride = co.out_stride;
    A.out = co.out;
    A.row_key = co.row_key;
    A.cbs = co.cbs;
    A.split = cross_split(lc, n_rows, n_tiles, co.may_split);
    A.n_rows_dev = co.n_rows_dev;
    A.out_rows = n_rows;   // the host bound: room of `out` behind an asynchronous build
    A.stats = stats;
    A.work = work;
    cudaError_t e = cudaMemsetAsync(work, 0, sizeof(unsigned int), lc.stream);
    if (e != cudaSuccess) return e;
    switch (co.mode) {
    case kModeOr:
        if constexpr (D == 2)
            if (co.words_per_row <= 2) return launch_cross_v<D, kModeOr, true>(cv, lc, A);
        return launch_cross_v<D, kModeOr, false>(cv, lc, A);
    case kModeBits: return launch_cross_v<D, kModeBits, false>(cv, lc, A);
    case kModeCount: return launch_cross_v<D, kModeCount, false>(cv, lc, A);
    case kModeFirst: return launch_cross_v<D, kModeFirst, false>(cv, lc, A);
    default: return cudaErrorInvalidValue;
    }
}

// mode bits: `out` must be zeroed by the caller (hits are OR-ed in); or: at most 32*kRedWords
// groups per launch, every word of a row is written; count: zeroed int32 [n_rows][out_stride];
// first: int32 [n_rows][out_stride] filled with INT_MAX (launch_cross_first_init / _finalize).
cudaError_t launch_cross_collide(const CtxView &cv, const LaunchCfg &lc, int64_t n_rows,
                                 const void *rows, int64_t n_cols, const void *tiles,
                                 const void *hdrs, const CrossOut &co, unsigned long long *stats,
                                 unsigned int *work) {
    if (n_rows <= 0 || n_cols <= 0) return cudaSuccess;
    if (co.mode == kModeOr && co.words_per_row > kRedWords) return cudaErrorInvalidValue;
    if (cv.model == 0)
        return launch_cross_t<2>(cv, lc, n_rows, rows, n_cols, tiles, hdrs, co, stats, work);
    if (cv.model == 1)
        return launch_cross_t<3>(cv, lc, n_rows, rows, n_cols, tiles, hdrs, co, stats, work);
    return cudaErrorNotSupported;
}

cudaError_t launch_fill_i32(const LaunchCfg &lc, int32_t *out, int64_t n, int32_t v) {
    if (n <= 0) return cudaSuccess;
    k_fill_i32<<<grid_for_n(lc, n, 256 * 4), 256, 0, lc.stream>>>(out, n, v);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_cross_first_finalize(const LaunchCfg &lc, int32_t *out, int64_t n) {
    if (n <= 0) return cudaSuccess;
    k_first_finalize<<<grid_for_n(lc, n, 256 * 4), 256, 0, lc.stream>>>(out, n);
    count_launch();
    return cudaGetLastError();
}

}  // namespace mrmp
