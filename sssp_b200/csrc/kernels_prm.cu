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

constexpr int kSampleThreads = 1024;
constexpr int kSampleWarps = kSampleThreads / 32;

// One CTA per agent (1024 threads: a 500-vertex roadmap usually fills in a single round).
// Round r tests tries [r*T, (r+1)*T); accepted tries are compacted in try order (ballot + warp
// prefix), so the result equals the sequential loop of prm.jl:193-199.  fixed_src (may be
// mapped host memory) holds the n_fixed leading vertices of every agent (init / goal).
template <class M>
__global__ void __launch_bounds__(kSampleThreads)
k_sample_free(CtxView cv, int agent0, int n_fixed, int nv, int nv_stride, uint64_t seed,
              double *__restrict__ nodes, int64_t *__restrict__ tries_out, int64_t max_tries,
              const double *__restrict__ fixed_src, int staged) {
    __shared__ int warp_cnt[kSampleWarps];
    __shared__ int accepted_sh;
    const ConSeg cs(cv, stage_segment(cv, cv.con_off, cv.con_len, smem_raw, staged));
    const int a = blockIdx.x;
    const int agent = agent0 + a;
    const int wanted = nv - n_fixed;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *out = nodes + ((int64_t)a * nv_stride + n_fixed) * M::SD;
    if (fixed_src && threadIdx.x < n_fixed * M::SD)
        nodes[(int64_t)a * nv_stride * M::SD + threadIdx.x] =
            fixed_src[(int64_t)a * n_fixed * M::SD + threadIdx.x];
    int accepted = 0;
    int64_t tries = wanted <= 0 ? 0 : -1;
    for (int64_t t0 = 0; accepted < wanted && t0 < max_tries; t0 += kSampleThreads) {
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
            for (int w = 0; w < kSampleWarps; ++w) tot += warp_cnt[w];
            accepted_sh = accepted + tot;
        }
        __syncthreads();
        accepted = accepted_sh;
        __syncthreads();
    }
    if (threadIdx.x == 0 && tries_out && (accepted < wanted || wanted <= 0)) tries_out[a] = tries;
}

// ---- all-pairs neighbour pass -------------------------------------------------------------
// One CTA = (agent, chunk of that agent's rows); the agent's vertex table is staged in shared
// memory once per CTA.  A warp walks its rows j through a small pipeline of per-warp rings so
// that every expensive stage runs on 32 dense lanes, across row boundaries:
//   gate   every lane tests kGateGroup candidates k per round (prm.jl:207, branch-free,
//          independent FP64 chains); survivors (j,k) are compacted into ring 1;
//   edge   (prm.jl:208) generic models: connect(q_j, q_k) per lane, a passing edge sets its bit
//          in the (row-zeroed) global bitmap and bumps the row count with L2 reductions;
//          point models: the field test of q_k, then the bit is set optimistically and the edge
//          is tested against every obstacle with the conservative midpoint cull of
//          point_model.cuh; only (edge, obstacle) pairs that survive go to ring 2;
//   exact  (point models) the reference's segment-point test on ring 2; a hit clears the bit
//          again (the first clear also takes the row count back).
// Bits and counts are order-independent, so it does not matter when a pair leaves a ring.
constexpr int kRing1 = 256;      // (j,k) entries per warp; in flight <= 31 + 32 * kGateGroup
constexpr int kGateGroup = 4;    // bitmap words gated per compaction round
constexpr int kRing2 = 512;      // (j,k,m) entries per warp; in flight <= 31 + 32 * kCullGroup
constexpr int kCullGroup = 8;    // obstacles culled per compaction round
constexpr int kAdjStageBytes = 96 * 1024;  // vertex tables up to this size are staged in smem
constexpr double kGateRadMin = 1e-150;     // above this, dist <= rad  <=>  dist_sq <= r2

template <class M>
struct AdjTraits {
    static constexpr bool kCull = false;
};
template <int D>
struct AdjTraits<PointModel<D>> {
    static constexpr bool kCull = true;
};

template <int SD>
__device__ __forceinline__ void load_node(const double *tab, int k, double *q) {
    if constexpr (SD % 2 == 0) {
        const double2 *p = reinterpret_cast<const double2 *>(tab + (int64_t)SD * k);
#pragma unroll
        for (int c = 0; c < SD / 2; ++c) {
            const double2 t = p[c];
            q[2 * c] = t.x;
            q[2 * c + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int c = 0; c < SD; ++c) q[c] = tab[(int64_t)SD * k + c];
    }
}

// exclusive prefix of c over the warp; total = sum over all lanes
__device__ __forceinline__ int warp_offsets(int c, int lane, int &total) {
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    total = __shfl_sync(0xffffffffu, incl, 31);
    return incl - c;
}

// SMEM: the connect segment and the agent's vertex table (padded to 32 * words entries) are
// staged in shared memory -- the normal case; the other instantiation reads both from global.
// The per-agent radii of a build travel in the kernel parameters when there are at most 64 agents:
// read from the mapped staging block (a PCIe round trip per CTA) they were 18 % of the kernel's
// stall samples (profiles/r2_i_adj_ncu_summary.md).
struct RadTab {
    int n;  // 0: read rad / rad_r2 through the pointers
    double rad[64], r2[64];
};

template <class M, bool SMEM, bool OPAD>
__global__ void __launch_bounds__(kThreads, 3)
k_prm_adjacency(CtxView cv, int agent0, int n_agents, int nv, int chunks,
                const double *__restrict__ rad, const double *__restrict__ rad_r2,
                const double *__restrict__ nodes, uint32_t *__restrict__ bitmap, int words,
                int32_t *__restrict__ row_cnt, int seg_bytes, int obs_pad,
                const __grid_constant__ RadTab rt) {
    constexpr int SD = M::SD;
    constexpr bool kCull = AdjTraits<M>::kCull;
    const int a = blockIdx.x / chunks, chunk = blockIdx.x - a * chunks;
    const double r = rt.n ? rt.rad[a] : rad[a], r2 = rt.n ? rt.r2[a] : rad_r2[a];
    const ConSeg cs(cv, stage_segment(cv, cv.con_off, cv.con_len, smem_raw, SMEM));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int agent = agent0 + a;
    unsigned char *sp = smem_raw + (SMEM ? seg_bytes : 0);
    int2 *ring1 = reinterpret_cast<int2 *>(sp) + warp * kRing1;
    sp += kWarps * kRing1 * 8;
    unsigned long long *ring2 = reinterpret_cast<unsigned long long *>(sp) + warp * kRing2;
    if (kCull) sp += kWarps * kRing2 * 8;
    // point models: obstacle SoA re-packed with the count padded to a multiple of kCullGroup
    // (far-away, zero-radius entries), so that a cull round reads at immediate offsets
    double *opad = reinterpret_cast<double *>(sp);
    if (kCull && OPAD) {
        for (int t = threadIdx.x; t < (SD + 1) * obs_pad; t += kThreads) {
            const int x = t / obs_pad, m = t - x * obs_pad;
            opad[t] = m < cs.n_obs ? cs.oc(x, m) : (x < SD ? 1.0e300 : 0.0);
        }
        sp += (SD + 1) * obs_pad * 8;
        if (!SMEM) __syncthreads();
    }
    const double *tab = nodes + (int64_t)a * nv * SD;
    if constexpr (SMEM) {
        double *st = reinterpret_cast<double *>(sp);
        const int n_dbl = nv * SD, n_pad = words * 32 * SD;
        for (int i = threadIdx.x; i < n_pad; i += kThreads) st[i] = i < n_dbl ? __ldg(tab + i) : 0.0;
        __syncthreads();
        tab = st;
    }
    const bool gate_all = isnan(r);
    const bool fast_gate = r >= kGateRadMin;  // false for NaN
    uint32_t *bits_a = bitmap + (int64_t)a * nv * words;
    int32_t *cnt_a = row_cnt + (int64_t)a * nv;
    const double rad_i = cs.rad(agent);
    (void)rad_i;
    (void)ring2;

    int head = 0, tail = 0;    // ring 1, warp-uniform
    int head2 = 0, tail2 = 0;  // ring 2 (point models)

    // generic models: the whole edge test on one lane
    auto connect_lane = [&](int2 e) {
        double qa[SD], qb[SD];
        load_node<SD>(tab, e.x, qa);
        load_node<SD>(tab, e.y, qb);
        if (M::connect_edge(cv, cs, qa, qb, agent)) {
            atomicOr(bits_a + (int64_t)e.x * words + (e.y >> 5), 1u << (e.y & 31));
            atomicAdd(cnt_a + e.x, 1);
        }
    };
    // point models, exact stage: dist(q_j, q_k, o_m) < o_m.r + rads[i]  (point2d.jl:63-65)
    auto exact_lane = [&](unsigned long long e) {
        if constexpr (kCull) {
            const int j = (int)(e >> 44), k = (int)((e >> 24) & 0xfffffu), m = (int)(e & 0xffffffu);
            double qa[SD], qb[SD], u[SD], c[SD];
            load_node<SD>(tab, j, qa);
            load_node<SD>(tab, k, qb);
#pragma unroll
            for (int x = 0; x < SD; ++x) {
                u[x] = dsub(qa[x], qb[x]);
                c[x] = cs.oc(x, m);
            }
            if (segpt_lt_simt<SD, true>(qa, qb, u, dotd<SD>(u, u), c, dadd(cs.oc(SD, m), rad_i),
                                        cs.t2_im(agent, m))) {
                const uint32_t bit = 1u << (k & 31);
                const uint32_t old = atomicAnd(bits_a + (int64_t)j * words + (k >> 5), ~bit);
                if (old & bit) atomicSub(cnt_a + j, 1);
            }
        }
    };
    // point models, edge stage for up to 32 ring-1 entries (warp-collective)
    auto cull_batch = [&](int n_valid) {
        if constexpr (kCull) {
            const bool valid = lane < n_valid;
            const int2 e = ring1[(head + lane) & (kRing1 - 1)];
            double mid[SD], R = 0.0;
            bool ok = false, wide = false;
            if (valid) {
                double qa[SD], qb[SD];
                load_node<SD>(tab, e.x, qa);
                load_node<SD>(tab, e.y, qb);
                ok = !out_of_field<SD>(qb, rad_i);  // q_to only (point2d.jl:60)
                if (ok) {
                    atomicOr(bits_a + (int64_t)e.x * words + (e.y >> 5), 1u << (e.y & 31));
                    atomicAdd(cnt_a + e.x, 1);
                }
                const double h = half_len_up<SD>(qa, qb);
                R = rad_i + kPad + h;
                wide = !(h < kCullCoordMax);
#pragma unroll
                for (int x = 0; x < SD; ++x) {
                    mid[x] = 0.5 * (qa[x] + qb[x]);
                    wide |= !(fabs(mid[x]) < kCullCoordMax);
                }
            }
            const unsigned long long key = ((unsigned long long)(unsigned)e.x << 44) |
                                           ((unsigned long long)(unsigned)e.y << 24);
            for (int m0 = 0; m0 < cs.n_obs; m0 += kCullGroup) {
                unsigned near = 0;
                if (ok && OPAD) {
                    const double *oc = opad + m0;
#pragma unroll
                    for (int u = 0; u < kCullGroup; ++u) {
                        double w2 = 0.0;
#pragma unroll
                        for (int x = 0; x < SD; ++x) {
                            const double w = oc[x * obs_pad + u] - mid[x];
                            w2 += w * w;
                        }
                        const double orad = oc[SD * obs_pad + u];
                        const double reach = R + orad;
                        const bool far = w2 > reach * reach && fabs(orad) < kCullCoordMax;
                        near |= (m0 + u < cs.n_obs && (wide || !far) ? 1u : 0u) << u;
                    }
                } else if (ok && !OPAD) {  // very many obstacles: read the ctx arrays at a clamped index
                    const double *oc = cs.s + cs.obs;
#pragma unroll
                    for (int u = 0; u < kCullGroup; ++u) {
                        const int m = m0 + u < cs.n_obs ? m0 + u : cs.n_obs - 1;
                        double w2 = 0.0;
#pragma unroll
                        for (int x = 0; x < SD; ++x) {
                            const double w = oc[x * cs.mp + m] - mid[x];
                            w2 += w * w;
                        }
                        const double orad = oc[SD * cs.mp + m];
                        const double reach = R + orad;
                        const bool far = w2 > reach * reach && fabs(orad) < kCullCoordMax;
                        near |= (m0 + u < cs.n_obs && (wide || !far) ? 1u : 0u) << u;
                    }
                }
                int total;
                int pos = tail2 + warp_offsets(__popc(near), lane, total);
                while (near) {
                    const int u = __ffs(near) - 1;
                    near &= near - 1;
                    ring2[pos & (kRing2 - 1)] = key | (unsigned)(m0 + u);
                    ++pos;
                }
                tail2 += total;
                __syncwarp();
                while (tail2 - head2 >= 32) {
                    exact_lane(ring2[(head2 + lane) & (kRing2 - 1)]);
                    head2 += 32;
                }
            }
        }
    };

    for (int j = chunk * kWarps + warp; j < nv; j += chunks * kWarps) {
        for (int t = lane; t < words; t += 32) bits_a[(int64_t)j * words + t] = 0u;
        if (lane == 0) cnt_a[j] = 0;
        double qj[SD];
        load_node<SD>(tab, j, qj);
        for (int t0 = 0; t0 < words; t0 += kGateGroup) {
            unsigned m = 0;
            const int k0 = 32 * t0 + lane;
            if (fast_gate) {
                // branch-free, independent chains; candidates past nv are masked out (staged
                // tables are padded, global ones are read at a clamped index)
#pragma unroll
                for (int u = 0; u < kGateGroup; ++u) {
                    const int k = k0 + 32 * u;
                    const bool valid = k < nv && k != j;
                    double qk[SD];
                    if constexpr (SMEM) {
                        if (t0 + u < words) load_node<SD>(tab, k, qk);  // warp-uniform guard
                        else load_node<SD>(tab, j, qk);
                    } else {
                        load_node<SD>(tab, valid ? k : j, qk);
                    }
                    const double s = M::dist_sq(qj, qk);  // prm.jl:207
                    m |= (valid && s <= r2 ? 1u : 0u) << u;
                }
            } else {
                for (int u = 0; u < kGateGroup; ++u) {
                    const int k = k0 + 32 * u;
                    if (k < nv && k != j) {
                        double qk[SD];
                        load_node<SD>(tab, k, qk);
                        if (gate_all || M::dist_le(cv, qj, qk, r, r2)) m |= 1u << u;
                    }
                }
            }
            // compaction by ballots: with only kGateGroup slots per lane this is cheaper than a
            // prefix sum plus a divergent loop over the set bits
            const unsigned lt = (1u << lane) - 1u;
#pragma unroll
            for (int u = 0; u < kGateGroup; ++u) {
                const bool pass = (m >> u) & 1u;
                const unsigned bal = __ballot_sync(0xffffffffu, pass);
                if (pass) ring1[(tail + __popc(bal & lt)) & (kRing1 - 1)] = make_int2(j, k0 + 32 * u);
                tail += __popc(bal);
            }
            __syncwarp();  // ring writes (and this row's zeroing) before the reads / reductions
            while (tail - head >= 32) {
                if constexpr (kCull) cull_batch(32);
                else connect_lane(ring1[(head + lane) & (kRing1 - 1)]);
                head += 32;
            }
        }
    }
    __syncwarp();
    if (tail - head > 0) {
        if constexpr (kCull) {
            cull_batch(tail - head);
        } else {
            if (lane < tail - head) connect_lane(ring1[(head + lane) & (kRing1 - 1)]);
        }
    }
    if constexpr (kCull) {
        __syncwarp();
        if (lane < tail2 - head2) exact_lane(ring2[(head2 + lane) & (kRing2 - 1)]);
    }
}

// ---- exclusive scan of the row counts ------------------------------------------------------
// CTA b owns rows [1024 b, 1024 b + 1024).  Its base offset is the sum of everything before:
// read from chunk_sum[0..b) when the caller ran k_chunk_sums first (many chunks), otherwise
// summed straight from row_cnt (few chunks: every load is independent, so the redundant reads
// pipeline and the whole scan is one short, fully parallel launch).
__device__ __forceinline__ long long block_sum_1024(long long v, long long *sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    long long w = sh[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    __syncthreads();
    return w;
}

__global__ void __launch_bounds__(1024)
k_chunk_sums(int64_t n_rows, const int32_t *__restrict__ row_cnt, long long *__restrict__ chunk_sum) {
    __shared__ long long sh[32];
    const int64_t r = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    const long long s = block_sum_1024(r < n_rows ? row_cnt[r] : 0, sh);
    if (threadIdx.x == 0) chunk_sum[blockIdx.x] = s;
}

__global__ void __launch_bounds__(1024)
k_row_scan(int64_t n_rows, const int32_t *__restrict__ row_cnt,
           const long long *__restrict__ chunk_sum, int32_t *__restrict__ row_ptr,
           int64_t *__restrict__ nnz_total, int64_t *__restrict__ nnz_mirror) {
    __shared__ long long sh[32];
    __shared__ long long warp_tot[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t base_row = (int64_t)blockIdx.x * 1024;
    long long part = 0;
    if (chunk_sum) {
        for (int64_t c = tid; c < blockIdx.x; c += 1024) part += chunk_sum[c];
    } else {
        for (int64_t r = tid; r < base_row; r += 1024) part += row_cnt[r];
    }
    const int64_t r = base_row + tid;
    const long long v = r < n_rows ? row_cnt[r] : 0;
    const long long base = block_sum_1024(part, sh);
    long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        long long w = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        warp_tot[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const long long before = base + (warp ? warp_tot[warp - 1] : 0) + incl - v;
    if (r < n_rows) row_ptr[r] = (int32_t)before;
    if (r == n_rows - 1) {
        row_ptr[n_rows] = (int32_t)(before + v);
        *nnz_total = before + v;
        if (nnz_mirror) *nnz_mirror = before + v;  // mapped host copy: no D2H copy node
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
        for (int t0 = 0; t0 < words; t0 += 32) {
            // lane t owns word t0 + t: the bitmap is sparse (a few bits per word), so a prefix
            // over the popcounts plus a short loop over the set bits beats one pass per word
            uint32_t w = t0 + lane < words ? __ldg(bitmap + row * words + t0 + lane) : 0u;
            int total;
            int64_t pos = base + warp_offsets(__popc(w), lane, total);
            const int k0 = 32 * (t0 + lane);
            while (w) {
                const int bit = __ffs(w) - 1;
                w &= w - 1;
                if (pos < col_cap) col_idx[pos] = k0 + bit;
                ++pos;
            }
            base += total;
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
                               int64_t *tries_out, int64_t max_tries, const double *fixed_src) {
    if (n_agents <= 0) return cudaSuccess;
    if (fixed_src && n_fixed * cv.d > kSampleThreads) return cudaErrorInvalidValue;
    const int sm = seg_smem_bytes(cv.con_len);
    cudaError_t e = cudaSuccess;
    MRMP_MODEL_SWITCH(cv, (e = set_smem(k_sample_free<M>, sm),
                           k_sample_free<M><<<n_agents, kSampleThreads, sm, lc.stream>>>(
                               cv, agent0, n_fixed, nv, nv_stride, seed, nodes, tries_out,
                               max_tries, fixed_src, sm > 0)));
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

// resident CTAs per SM of one adjacency instantiation, cached per (model, staging, smem size)
template <class K>
static int adj_occupancy(K kernel, int slot, int smem) {
    static int cached_smem[32], cached_occ[32];
    if (cached_occ[slot] == 0 || cached_smem[slot] != smem) {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kThreads, smem) !=
                cudaSuccess || occ < 1)
            occ = 1;
        cached_smem[slot] = smem;
        cached_occ[slot] = occ;
    }
    return cached_occ[slot];
}

// one wave: (agent, chunk) CTAs fill the resident slots without a ragged second wave
template <class M, bool SMEM, bool OPAD>
static cudaError_t launch_adj_t(const CtxView &cv, const LaunchCfg &lc, int agent0, int n_agents,
                                int nv, const double *rad_r2, const double *rad,
                                const double *nodes, uint32_t *bitmap, int words,
                                int32_t *row_cnt, int seg, int sm, int obs_pad,
                                const double *rad_host, const double *r2_host) {
    auto kern = k_prm_adjacency<M, SMEM, OPAD>;
    RadTab rt;
    rt.n = 0;
    if (rad_host && r2_host && n_agents <= 64) {
        rt.n = n_agents;
        for (int a = 0; a < n_agents; ++a) {
            rt.rad[a] = rad_host[a];
            rt.r2[a] = r2_host[a];
        }
    }
    cudaError_t e = set_smem(kern, sm);
    if (e != cudaSuccess) return e;
    const int slots = lc.sm_count * adj_occupancy(kern, 4 * cv.model + (SMEM ? 1 : 0) + (OPAD ? 2 : 0), sm);
    const int max_chunks = (nv + kWarps - 1) / kWarps;  // at most one row per warp
    int chunks = slots / n_agents;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    kern<<<n_agents * chunks, kThreads, sm, lc.stream>>>(cv, agent0, n_agents, nv, chunks, rad,
                                                         rad_r2, nodes, bitmap, words, row_cnt, seg,
                                                         obs_pad, rt);
    return cudaSuccess;
}

cudaError_t launch_prm_adjacency(const CtxView &cv, const LaunchCfg &lc, int agent0, int n_agents,
                                 int nv, const double *rad_r2, const double *rad,
                                 const double *nodes, uint32_t *bitmap, int words,
                                 int32_t *row_cnt, const double *rad_host, const double *r2_host) {
    if (n_agents <= 0 || nv <= 0) return cudaSuccess;
    const bool cull = cv.model <= 1;  // AdjTraits<PointModel<D>>::kCull: second ring
    if (cull && (cv.n_obs >= (1 << 24) || nv > (1 << 20)))
        return cudaErrorInvalidValue;
    const int rings = kWarps * 8 * (kRing1 + (cull ? kRing2 : 0));
    const int seg = (seg_smem_bytes(cv.con_len) + 15) & ~15;
    const int64_t tab = (int64_t)words * 32 * cv.d * 8;
    const bool smem_on = seg > 0 && tab <= kAdjStageBytes;
    int obs_pad = 0;  // padded obstacle count of the re-packed smem copy (point models)
    if (cull && cv.n_obs > 0 && (int64_t)cv.n_obs * (cv.d + 1) * 8 <= 32 * 1024)
        obs_pad = (cv.n_obs + kCullGroup - 1) / kCullGroup * kCullGroup;
    const int sm = rings + obs_pad * (cv.d + 1) * 8 + (smem_on ? seg + (int)tab : 0);
    cudaError_t e = cudaSuccess;
#define ADJ_ARGS cv, lc, agent0, n_agents, nv, rad_r2, rad, nodes, bitmap, words, row_cnt, seg, sm, obs_pad, rad_host, r2_host
    if (obs_pad > 0) {  // point models with a re-packed obstacle copy
        if (smem_on) {
            if (cv.model == 0) e = launch_adj_t<PointModel<2>, true, true>(ADJ_ARGS);
            else e = launch_adj_t<PointModel<3>, true, true>(ADJ_ARGS);
        } else {
            if (cv.model == 0) e = launch_adj_t<PointModel<2>, false, true>(ADJ_ARGS);
            else e = launch_adj_t<PointModel<3>, false, true>(ADJ_ARGS);
        }
    } else if (smem_on) {
        MRMP_MODEL_SWITCH(cv, (e = launch_adj_t<M, true, false>(ADJ_ARGS)));
    } else {
        MRMP_MODEL_SWITCH(cv, (e = launch_adj_t<M, false, false>(ADJ_ARGS)));
    }
#undef ADJ_ARGS
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_row_scan(const LaunchCfg &lc, int64_t n_rows, const int32_t *row_cnt,
                            int32_t *row_ptr, int64_t *nnz_total, long long *chunk_ws,
                            int64_t chunk_ws_len, int64_t *nnz_mirror) {
    if (n_rows <= 0) return cudaErrorInvalidValue;
    const int64_t chunks = (n_rows + 1023) / 1024;
    const bool two_level = chunks > 64;
    if (two_level) {
        if (!chunk_ws || chunk_ws_len < chunks) return cudaErrorInvalidValue;
        k_chunk_sums<<<(int)chunks, 1024, 0, lc.stream>>>(n_rows, row_cnt, chunk_ws);
        count_launch();
    }
    k_row_scan<<<(int)chunks, 1024, 0, lc.stream>>>(n_rows, row_cnt, two_level ? chunk_ws : nullptr,
                                                     row_ptr, nnz_total, nnz_mirror);
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

// the same pack, written to the slot of this rank in the buffers of up to 8 ranks at once (peer
// mappings over NVLink; consecutive threads write consecutive words, so every warp store is one
// 128-byte transaction per destination)
__global__ void __launch_bounds__(kThreads)
k_csr_pack_push(int64_t n_rows, int64_t rows_cap, int64_t cap_nnz,
                const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col_idx,
                CsrPushDst dst) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (int64_t r = tid; r < rows_cap; r += stride) {
        const int32_t v = r < n_rows ? row_ptr[r + 1] - row_ptr[r] : 0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < dst.n) dst.slot[k][r] = v;
    }
    // (a shard larger than its slot is truncated here; the host detects it from nnz)
    int64_t nnz = n_rows > 0 ? row_ptr[n_rows] : 0;
    if (nnz > cap_nnz) nnz = cap_nnz;
    for (int64_t e = tid; e < nnz; e += stride) {
        const int32_t v = col_idx[e];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < dst.n) dst.slot[k][rows_cap + e] = v;
    }
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

cudaError_t launch_csr_pack_push(const LaunchCfg &lc, int64_t n_rows, int64_t rows_cap,
                                 int64_t cap_nnz, const int32_t *row_ptr, const int32_t *col_idx,
                                 const CsrPushDst &dst) {
    k_csr_pack_push<<<lc.sm_count * 4, kThreads, 0, lc.stream>>>(n_rows, rows_cap, cap_nnz, row_ptr,
                                                                 col_idx, dst);
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
