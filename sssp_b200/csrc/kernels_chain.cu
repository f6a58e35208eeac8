// kernels_chain.cu -- connect / collide kernels of line2d, snake2d, arm33 and capsule3d
// (src/models/line2d.jl, snake2d.jl, arm33.jl, capsule3d.jl), one template over the model.
//
// Same execution scheme as kernels_arm.cu (a WARP per check): connect spreads the postures of
// a move over the lanes with a ballot per 32 postures; collide builds the posture tables of
// both moves once per check in per-warp shared memory, lanes own the entries of the second
// table and read the first by broadcast, a conservative midpoint bound per link pair
// (|m1 - m2|^2 > (thr + 2^-20 + h1 + h2)^2 => farther than thr) discards most link pairs, the
// survivors are compacted through a per-warp queue and the exact reference arithmetic
// (segseg_lt) runs on dense lanes; the first hit ends the check.
#include "chain_model.cuh"
#include "point_model.cuh"

namespace mrmp {

extern __shared__ __align__(16) unsigned char smem_raw[];

constexpr int kThreads = 256;
constexpr int kPairWarps = 4;
constexpr int kPairThreads = 32 * kPairWarps;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kTableBudget = 12 * 1024;       // bytes of posture tables per warp

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

template <int MODEL>
struct Lay {  // table geometry of a model
    using C = Chain<MODEL>;
    static constexpr int PW = C::P * C::W, LW = C::L * C::W, LL = C::L * C::L;
    static constexpr int IE = PW + LW;                   // I entry: points + link midpoints
    static constexpr int JE = (PW + 3 + 1) & ~1;         // J entry: points, G, thr, T2 (even)
    static constexpr int CAP = kTableBudget / ((IE + JE) * 8) < 1023 ? kTableBudget / ((IE + JE) * 8) : 1023;
    static constexpr int QCAP = 32 + 32 * LL;
    static constexpr int WARP_BYTES = CAP * (IE + JE) * 8 + QCAP * 4 + ((QCAP & 1) ? 4 : 0);
};

__device__ __forceinline__ ChainPar chain_par(const CtxView &cv) {
    ChainPar p;
    p.sr.ok = cv.step_rat_ok;
    p.sr.num = cv.step_rat_num;
    p.sr.den = cv.step_rat_den;
    p.step = cv.step_dist;
    p.safety = cv.safety_dist;
    p.safety_t2 = cv.safety_t2;
    p.max_dist = cv.max_dist;
    return p;
}

// per-agent geometry from the connect (con = true) or collide segment
template <int MODEL>
__device__ __forceinline__ ChainGeom chain_geom(const CtxView &cv, const double *seg, bool con,
                                                int i) {
    using T = ChainTraits<MODEL>;
    const int rads = con ? cv.rads_off - cv.con_off : cv.crads_off - cv.col_off;
    const int base = con ? cv.base_off - cv.con_off : cv.cbase_off - cv.col_off;
    const int aux = con ? cv.aux_off - cv.con_off : cv.caux_off - cv.col_off;
    ChainGeom g;
    g.rad = seg[rads + i];
    g.len = MODEL == kModelCapsule3D ? seg[aux + i] : g.rad;
#pragma unroll
    for (int x = 0; x < 3; ++x) g.base[x] = (T::kHasBase && x < T::W) ? seg[base + T::W * i + x] : 0.0;
    return g;
}

template <int MODEL>
__device__ __forceinline__ ChainObs chain_obs(const CtxView &cv, const double *con, int i) {
    ChainObs o;
    o.s = con;
    o.obs = cv.obs_off - cv.con_off;
    o.t2 = cv.t2_obs_off - cv.con_off + (ChainTraits<MODEL>::kObsAddsRad ? i * cv.mp : 0);
    o.mp = cv.mp;
    o.n_obs = cv.n_obs;
    return o;
}

// pair threshold of agents (i, j): rads[i] + rads[j] with its exact-compare value from the pair
// table (capsule3d.jl:163), or safety_dist
template <int MODEL>
__device__ __forceinline__ void pair_threshold(const CtxView &cv, const double *col, int i, int j,
                                               double ri, double rj, double &thr, double &T2) {
    if constexpr (ChainTraits<MODEL>::kPairRadSum) {
        thr = dadd(ri, rj);
        T2 = cv.has_pair_tab ? col[cv.t2_pair_off - cv.col_off + i * cv.n_agents + j] : -1.0;
    } else {
        thr = cv.safety_dist;
        T2 = cv.safety_t2;
    }
}

// ---- warp-cooperative move setup: the diff_angles of up to two moves on separate lanes ------
template <int MODEL>
__device__ __forceinline__ void warp_chain_moves(const ChainPar &cp, int lane, const double *qf0,
                                                 const double *qt0, const double *qf1,
                                                 const double *qt1, ChainMove<MODEL> &m0,
                                                 ChainMove<MODEL> &m1, bool two) {
    using C = Chain<MODEL>;
    constexpr int SD = C::SD;
    // lane = move * 2*SD + kind * SD + component; kind 0: interpolation target, 1: dist component
    double val = 0.0;
    if (lane < (two ? 4 : 2) * SD && lane < 32) {
        const int mv = lane / (2 * SD), r = lane - mv * 2 * SD, kind = r / SD, c = r - kind * SD;
        const double *qf = mv ? qf1 : qf0, *qt = mv ? qt1 : qt0;
        if (c < C::NPOS)
            val = kind ? dsub(qf[c], qt[c]) : qt[c];
        else
            val = kind ? ddiv(diff_angles(qf[c], qt[c]), kPi64) : diff_angles(qt[c], qf[c]);
    }
    double v0[SD], v1[SD];
#pragma unroll
    for (int c = 0; c < SD; ++c) {
        m0.tgt[c] = __shfl_sync(kFull, val, c);
        v0[c] = __shfl_sync(kFull, val, SD + c);
        m1.tgt[c] = __shfl_sync(kFull, val, (2 * SD + c) & 31);
        v1[c] = __shfl_sync(kFull, val, (3 * SD + c) & 31);
        m0.from[c] = qf0[c];
        m1.from[c] = qf1[c];
    }
    const double D0 = norm_val<SD>(v0);
    const double D1 = two ? norm_val<SD>(v1) : 0.0;
    ArmSched s;
    s.len = 0;
    s.rational = 0;
    if (lane < (two ? 2 : 1)) s = arm_schedule(cp.sr, cp.step, lane ? D1 : D0);
    m0.sch.len = __shfl_sync(kFull, s.len, 0);
    m0.sch.rational = __shfl_sync(kFull, s.rational, 0);
    m1.sch.len = __shfl_sync(kFull, s.len, 1);
    m1.sch.rational = __shfl_sync(kFull, s.rational, 1);
    m0.sch.num = m1.sch.num = cp.sr.num;
    m0.sch.den = m1.sch.den = cp.sr.den;
    m0.sch.step = m1.sch.step = cp.step;
    m0.sch.D = D0;
    m1.sch.D = D1;
}
static_assert(4 * 6 <= 32, "two moves of a 6-DOF model fit the lanes of a warp");

// ---- thread-per-check kernels --------------------------------------------------------------------
template <int MODEL>
__global__ void __launch_bounds__(kThreads)
k_chain_connect_points(CtxView cv, int64_t n, const int32_t *__restrict__ agent,
                       const double *__restrict__ q, uint8_t *__restrict__ out, int staged) {
    constexpr int SD = Chain<MODEL>::SD;
    const double *con = stage_segment(cv, cv.con_off, cv.con_len, smem_raw, staged);
    const ChainPar cp = chain_par(cv);
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
         k += (int64_t)gridDim.x * blockDim.x) {
        double s[SD];
        load_state<SD>(q, k, s);
        const int i = __ldg(agent + k);
        out[k] = chain_connect_point<MODEL>(cp, chain_obs<MODEL>(cv, con, i),
                                            chain_geom<MODEL>(cv, con, true, i), s)
                     ? 1
                     : 0;
    }
}

template <int MODEL>
__global__ void __launch_bounds__(kThreads)
k_chain_collide_static(CtxView cv, int64_t n, const int32_t *__restrict__ ai,
                       const int32_t *__restrict__ aj, const double *__restrict__ qi,
                       const double *__restrict__ qj, uint8_t *__restrict__ out, int staged) {
    constexpr int SD = Chain<MODEL>::SD;
    const double *col = stage_segment(cv, cv.col_off, cv.col_len, smem_raw, staged);
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
         k += (int64_t)gridDim.x * blockDim.x) {
        double a[SD], b[SD];
        load_state<SD>(qi, k, a);
        load_state<SD>(qj, k, b);
        const int i = __ldg(ai + k), j = __ldg(aj + k);
        const ChainGeom gi = chain_geom<MODEL>(cv, col, false, i), gj = chain_geom<MODEL>(cv, col, false, j);
        double thr, T2;
        pair_threshold<MODEL>(cv, col, i, j, gi.rad, gj.rad, thr, T2);
        out[k] = chain_collide_static<MODEL>(gi, gj, a, b, thr, T2) ? 1 : 0;
    }
}

// ---- connect(q_from, q_to, i): warp per check ---------------------------------------------------
template <int MODEL>
__global__ void __launch_bounds__(kThreads)
k_chain_connect_edges(CtxView cv, int64_t n, const int32_t *__restrict__ agent,
                      const double *__restrict__ qf, const double *__restrict__ qt,
                      uint8_t *__restrict__ out, int staged) {
    using C = Chain<MODEL>;
    constexpr int SD = C::SD;
    const double *con = stage_segment(cv, cv.con_off, cv.con_len, smem_raw, staged);
    const ChainPar cp = chain_par(cv);
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n; k += n_warps) {
        double from[SD], to[SD];
        load_state<SD>(qf, k, from);
        load_state<SD>(qt, k, to);
        const int i = __ldg(agent + k);
        ChainMove<MODEL> mv, unused;
        warp_chain_moves<MODEL>(cp, lane, from, to, from, to, mv, unused, false);
        mv.g = chain_geom<MODEL>(cv, con, true, i);
        const ChainObs ob = chain_obs<MODEL>(cv, con, i);
        bool ok = !(cp.max_dist == cp.max_dist && mv.sch.D > cp.max_dist);
        const long long ns = mv.sch.n();
        for (long long s0 = 0; ok && s0 < ns; s0 += 32) {
            const long long s = s0 + lane;
            bool bad = false;
            if (s < ns) {
                double pts[C::P * C::W];
                mv.posture(s, pts);
                bad = !C::posture_free(cp, ob, mv.g, pts, false);
            }
            ok = !__any_sync(kFull, bad);
        }
        if (lane == 0) out[k] = ok ? 1 : 0;
    }
}

// ---- the table sweep --------------------------------------------------------------------------------
// I entry: pts[PW] | link midpoints[LW];  J entry: pts[PW] | G | thr | T2
template <int MODEL>
__device__ __forceinline__ bool queue_entry_hit(unsigned ent, const double *I, const double *J) {
    using Y = Lay<MODEL>;
    const double *ei = I + (ent & 0x3ffu) * Y::IE;
    const double *ej = J + ((ent >> 10) & 0x3ffu) * Y::JE;
    double pi[Y::PW], pj[Y::PW];
#pragma unroll
    for (int x = 0; x < Y::PW; ++x) {
        pi[x] = ei[x];
        pj[x] = ej[x];
    }
    return Chain<MODEL>::link_pair_hit((int)(ent >> 20), pi, pj, ej[Y::PW + 1], ej[Y::PW + 2]);
}

template <int MODEL>
__device__ __forceinline__ bool warp_sweep(const double *I, int nI, const double *J, int nJ,
                                           unsigned *queue, int lane, unsigned &n_exact) {
    using C = Chain<MODEL>;
    using Y = Lay<MODEL>;
    constexpr int W = C::W, L = C::L;
    int qlen = 0;
    const unsigned lt = (1u << lane) - 1u;
    for (int j0 = 0; j0 < nJ; j0 += 32) {
        const int jj = j0 + lane;
        const bool jv = jj < nJ;
        const double *ej = J + (jv ? jj : 0) * Y::JE;
        double mj[L * W];
#pragma unroll
        for (int l = 0; l < L; ++l)
#pragma unroll
            for (int x = 0; x < W; ++x) mj[l * W + x] = dmul(dadd(ej[l * W + x], ej[(l + 1) * W + x]), 0.5);
        const double G = ej[Y::PW];
        for (int ii = 0; ii < nI; ++ii) {
            const double *mi = I + ii * Y::IE + Y::PW;  // broadcast reads
            unsigned mask = 0;
#pragma unroll
            for (int li = 0; li < L; ++li)
#pragma unroll
                for (int lj = 0; lj < L; ++lj) {
                    double d2 = 0.0;
#pragma unroll
                    for (int x = 0; x < W; ++x) {
                        const double dx = dsub(mi[li * W + x], mj[lj * W + x]);
                        d2 = x == 0 ? dmul(dx, dx) : dadd(d2, dmul(dx, dx));
                    }
                    mask |= !(d2 > G) ? 1u << (li * L + lj) : 0u;
                }
            if (!jv) mask = 0;
            if (!__any_sync(kFull, mask != 0)) continue;
#pragma unroll
            for (int lp = 0; lp < L * L; ++lp) {
                const bool on = (mask >> lp) & 1u;
                const unsigned bal = __ballot_sync(kFull, on);
                if (on)
                    queue[qlen + __popc(bal & lt)] = (unsigned)ii | ((unsigned)jj << 10) | ((unsigned)lp << 20);
                qlen += __popc(bal);
            }
            __syncwarp();
            while (qlen >= 32) {
                qlen -= 32;
                const bool hit = queue_entry_hit<MODEL>(queue[qlen + lane], I, J);
                n_exact += 32;
                if (__any_sync(kFull, hit)) return true;
            }
            __syncwarp();
        }
    }
    bool hit = false;
    if (lane < qlen) hit = queue_entry_hit<MODEL>(queue[lane], I, J);
    n_exact += qlen;
    __syncwarp();
    return __any_sync(kFull, hit);
}

template <int MODEL>
__device__ __forceinline__ void fill_I(double *I, const ChainMove<MODEL> &mv, long long s0, int cnt,
                                       int lane) {
    using C = Chain<MODEL>;
    using Y = Lay<MODEL>;
    for (int s = lane; s < cnt; s += 32) {
        double pts[Y::PW];
        mv.posture(s0 + s, pts);
        double *e = I + s * Y::IE;
#pragma unroll
        for (int x = 0; x < Y::PW; ++x) e[x] = pts[x];
#pragma unroll
        for (int l = 0; l < C::L; ++l)
#pragma unroll
            for (int x = 0; x < C::W; ++x)
                e[Y::PW + l * C::W + x] = dmul(dadd(pts[l * C::W + x], pts[(l + 1) * C::W + x]), 0.5);
    }
}

template <int MODEL>
__device__ __forceinline__ void write_J(double *e, const double *pts, double G, double thr, double T2) {
    using Y = Lay<MODEL>;
#pragma unroll
    for (int x = 0; x < Y::PW; ++x) e[x] = pts[x];
    e[Y::PW] = G;
    e[Y::PW + 1] = thr;
    e[Y::PW + 2] = T2;
}

template <int MODEL>
__device__ __forceinline__ void fill_J(double *J, const ChainMove<MODEL> &mv, long long s0, int cnt,
                                       double G, double thr, double T2, int lane) {
    using Y = Lay<MODEL>;
    for (int s = lane; s < cnt; s += 32) {
        double pts[Y::PW];
        mv.posture(s0 + s, pts);
        write_J<MODEL>(J + s * Y::JE, pts, G, thr, T2);
    }
}

// cull bound of an agent pair: link pairs whose midpoints are farther apart than
// thr + 2^-20 + (len_i + len_j)/2 cannot be closer than thr
__device__ __forceinline__ double chain_pair_bound(double thr, double li, double lj) {
    const double r = (thr > 0.0 ? thr : 0.0) + kPad + 0.5 * (fabs(li) + fabs(lj)) * (1.0 + 1e-9);
    return r * r * (1.0 + 1e-9);
}

template <int MODEL>
__device__ __forceinline__ bool warp_chain_collide_pair(const CtxView &cv, const ChainPar &cp,
                                                        const double *col, const double *qif,
                                                        const double *qit, const double *qjf,
                                                        const double *qjt, int i, int j, double *I,
                                                        double *J, unsigned *queue, int lane,
                                                        unsigned &n_exact) {
    using Y = Lay<MODEL>;
    ChainMove<MODEL> mi, mj;
    warp_chain_moves<MODEL>(cp, lane, qif, qit, qjf, qjt, mi, mj, true);
    mi.g = chain_geom<MODEL>(cv, col, false, i);
    mj.g = chain_geom<MODEL>(cv, col, false, j);
    double thr, T2;
    pair_threshold<MODEL>(cv, col, i, j, mi.g.rad, mj.g.rad, thr, T2);
    const double G = chain_pair_bound(thr, mi.g.len, mj.g.len);
    const long long ni = mi.sch.n(), nj = mj.sch.n();
    for (long long i0 = 0; i0 < ni; i0 += Y::CAP) {
        const int ci = (int)(ni - i0 < Y::CAP ? ni - i0 : Y::CAP);
        fill_I<MODEL>(I, mi, i0, ci, lane);
        for (long long j0 = 0; j0 < nj; j0 += Y::CAP) {
            const int cj = (int)(nj - j0 < Y::CAP ? nj - j0 : Y::CAP);
            if (i0 == 0 || nj > Y::CAP) fill_J<MODEL>(J, mj, j0, cj, G, thr, T2, lane);
            __syncwarp();
            if (warp_sweep<MODEL>(I, ci, J, cj, queue, lane, n_exact)) return true;
            __syncwarp();
        }
    }
    return false;
}

// ---- sources of 6-arity checks ----------------------------------------------------------------
template <int SD>
struct PairCheckN {
    int i, j;
    double qif[SD], qit[SD], qjf[SD], qjt[SD];
};

template <int SD>
struct ListSrcN {
    const int32_t *ai, *aj;
    const double *qif, *qit, *qjf, *qjt;
    uint8_t *out;
    __device__ bool fetch(int64_t k, PairCheckN<SD> &p) const {
        p.i = __ldg(ai + k);
        p.j = __ldg(aj + k);
        load_state<SD>(qif, k, p.qif);
        load_state<SD>(qit, k, p.qit);
        load_state<SD>(qjf, k, p.qjf);
        load_state<SD>(qjt, k, p.qjt);
        return true;
    }
    __device__ void emit(int64_t k, bool ran, bool hit) const { out[k] = (ran && hit) ? 1 : 0; }
    __device__ bool cbs() const { return false; }
};

template <int SD>
struct IndexedSrcN {
    const int32_t *ai, *aj, *vif, *vit, *vjf, *vjt;
    const double *nodes;
    int agent0, nv_stride;
    uint8_t *out;
    __device__ bool fetch(int64_t k, PairCheckN<SD> &p) const {
        p.i = __ldg(ai + k);
        p.j = __ldg(aj + k);
        const int64_t bi = (int64_t)(p.i - agent0) * nv_stride, bj = (int64_t)(p.j - agent0) * nv_stride;
        load_state<SD>(nodes, bi + __ldg(vif + k), p.qif);
        load_state<SD>(nodes, bi + __ldg(vit + k), p.qit);
        load_state<SD>(nodes, bj + __ldg(vjf + k), p.qjf);
        load_state<SD>(nodes, bj + __ldg(vjt + k), p.qjt);
        return true;
    }
    __device__ void emit(int64_t k, bool ran, bool hit) const { out[k] = (ran && hit) ? 1 : 0; }
    __device__ bool cbs() const { return false; }
};

template <int SD>
struct CrossSrcN {
    const int32_t *row_agent;
    const double *row_from, *row_to;
    int64_t n_cols;
    const int32_t *col_agent;
    const double *col_from, *col_to;
    const int32_t *col_vf, *col_vt;
    const double *nodes;
    int agent0, nv_stride;
    CrossReq rq;
    int64_t out_stride;
    uint32_t *out;  // initialised by the caller (abi.cu: cross_run_dev)
    __device__ int64_t slot_of(int64_t c) const {
        return rq.col_group ? __ldg(rq.col_group + c) : (rq.group_size > 0 ? c / rq.group_size : c);
    }
    __device__ bool cbs() const { return rq.cbs != 0; }  // cbs.jl:343-346: column agent first
    __device__ bool fetch(int64_t k, PairCheckN<SD> &p) const {
        const int64_t r = k / n_cols, c = k - r * n_cols;
        const int ri = __ldg(row_agent + r), cj = __ldg(col_agent + c);
        if (ri == cj) return false;
        if (rq.row_key && !(__ldg(rq.col_key + c) > __ldg(rq.row_key + r))) return false;
        if (rq.mode == kCrossOr) {
            const int64_t b = slot_of(c);
            if ((*(volatile uint32_t *)(out + r * out_stride + (b >> 5)) >> (b & 31)) & 1u) return false;
        }
        double *rf = rq.cbs ? p.qjf : p.qif, *rt = rq.cbs ? p.qjt : p.qit;
        double *cf = rq.cbs ? p.qif : p.qjf, *ct = rq.cbs ? p.qit : p.qjt;
        p.i = rq.cbs ? cj : ri;
        p.j = rq.cbs ? ri : cj;
        load_state<SD>(row_from, r, rf);
        load_state<SD>(row_to, r, rt);
        if (col_from) {
            load_state<SD>(col_from, c, cf);
            load_state<SD>(col_to, c, ct);
        } else {
            const int64_t bj = (int64_t)(cj - agent0) * nv_stride;
            load_state<SD>(nodes, bj + __ldg(col_vf + c), cf);
            load_state<SD>(nodes, bj + __ldg(col_vt + c), ct);
        }
        return true;
    }
    __device__ void emit(int64_t k, bool ran, bool hit) const {
        if (!(ran && hit)) return;
        const int64_t r = k / n_cols, c = k - r * n_cols, b = slot_of(c);
        if (rq.mode == kCrossCount)
            atomicAdd(reinterpret_cast<int *>(out) + r * out_stride + b, 1);
        else if (rq.mode == kCrossFirst)
            atomicMin(reinterpret_cast<int *>(out) + r * out_stride + b, (int)c);
        else
            atomicOr(out + r * out_stride + (b >> 5), 1u << (b & 31));
    }
};

template <int SD>
struct ConfigSrcN {
    const double *Qf, *Qt;
    int N;
    int64_t pairs;
    unsigned long long *first;
    __device__ void decode(int64_t k, int64_t &cfg, int &i, int &j) const {
        cfg = k / pairs;
        int64_t r = k - cfg * pairs;
        i = 0;
        while (r >= N - 1 - i) {
            r -= N - 1 - i;
            ++i;
        }
        j = i + 1 + (int)r;
    }
    __device__ bool fetch(int64_t k, PairCheckN<SD> &p) const {
        int64_t cfg;
        decode(k, cfg, p.i, p.j);
        if (*(volatile unsigned long long *)(first + cfg) < (unsigned long long)(p.i * N + p.j))
            return false;
        const int64_t base = cfg * N;
        load_state<SD>(Qf, base + p.i, p.qif);
        load_state<SD>(Qt, base + p.i, p.qit);
        load_state<SD>(Qf, base + p.j, p.qjf);
        load_state<SD>(Qt, base + p.j, p.qjt);
        return true;
    }
    __device__ void emit(int64_t k, bool ran, bool hit) const {
        if (ran && hit) {
            int64_t cfg;
            int i, j;
            decode(k, cfg, i, j);
            atomicMin(first + cfg, (unsigned long long)(i * N + j));
        }
    }
    __device__ bool cbs() const { return false; }
};

template <int MODEL, class Src>
__global__ void __launch_bounds__(kPairThreads)
k_chain_collide(CtxView cv, int64_t n, Src src, unsigned long long *stats,
                unsigned long long *work, int staged, int seg_bytes) {
    using Y = Lay<MODEL>;
    constexpr int SD = Chain<MODEL>::SD;
    const double *col = stage_segment(cv, cv.col_off, cv.col_len, smem_raw, staged);
    const ChainPar cp = chain_par(cv);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *wb = smem_raw + seg_bytes + (size_t)warp * Y::WARP_BYTES;
    double *I = reinterpret_cast<double *>(wb);
    double *J = I + Y::CAP * Y::IE;
    unsigned *queue = reinterpret_cast<unsigned *>(J + Y::CAP * Y::JE);
    unsigned n_exact = 0;
    // checks differ in cost by orders of magnitude (S_i x S_j posture pairs, early exits):
    // warps pull the next check from a counter instead of striding
    for (;;) {
        unsigned long long kk = 0;
        if (lane == 0) kk = atomicAdd(work, 1ull);
        const int64_t k = (int64_t)__shfl_sync(kFull, kk, 0);
        if (k >= n) break;
        PairCheckN<SD> p;
        const bool run = src.fetch(k, p);
        bool hit = false;
        if (run && src.cbs()) {  // collide(q_j_to, q_i_to, j, i), cbs.jl:345
            const ChainGeom gi = chain_geom<MODEL>(cv, col, false, p.i), gj = chain_geom<MODEL>(cv, col, false, p.j);
            double thr, T2;
            pair_threshold<MODEL>(cv, col, p.i, p.j, gi.rad, gj.rad, thr, T2);
            hit = chain_collide_static<MODEL>(gi, gj, p.qit, p.qjt, thr, T2);
        }
        if (run && !hit)
            hit = warp_chain_collide_pair<MODEL>(cv, cp, col, p.qif, p.qit, p.qjf, p.qjt, p.i, p.j, I,
                                                 J, queue, lane, n_exact);
        if (lane == 0) src.emit(k, run, hit);
        __syncwarp();
    }
    if (lane == 0 && stats && n_exact) atomicAdd(stats, (unsigned long long)n_exact);
}

// ---- collide(Q, q_to, i): warp per candidate, CTA-shared table of the static others --------------
template <int MODEL>
__global__ void __launch_bounds__(kPairThreads)
k_chain_config_moves(CtxView cv, int agent_i, const double *__restrict__ Q, int64_t n_cand,
                     const double *__restrict__ q_to, uint8_t *__restrict__ out,
                     unsigned long long *stats, int staged, int seg_bytes, int j_cap) {
    using C = Chain<MODEL>;
    using Y = Lay<MODEL>;
    constexpr int SD = C::SD;
    const double *col = stage_segment(cv, cv.col_off, cv.col_len, smem_raw, staged);
    const ChainPar cp = chain_par(cv);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = cv.n_agents;
    double *J = reinterpret_cast<double *>(smem_raw + seg_bytes);
    const size_t warp_bytes = (size_t)Y::CAP * Y::IE * 8 + (size_t)((Y::QCAP + 1) & ~1) * 4;
    unsigned char *wb = smem_raw + seg_bytes + (size_t)j_cap * Y::JE * 8 + (size_t)warp * warp_bytes;
    double *I = reinterpret_cast<double *>(wb);
    unsigned *queue = reinterpret_cast<unsigned *>(I + Y::CAP * Y::IE);
    const ChainGeom gi = chain_geom<MODEL>(cv, col, false, agent_i);
    for (int t = threadIdx.x; t < N - 1; t += blockDim.x) {
        const int j = t < agent_i ? t : t + 1;
        double qj[SD], pts[Y::PW];
        load_state<SD>(Q, j, qj);
        const ChainGeom gj = chain_geom<MODEL>(cv, col, false, j);
        C::points(qj, gj, pts);
        double thr, T2;
        pair_threshold<MODEL>(cv, col, agent_i, j, gi.rad, gj.rad, thr, T2);
        write_J<MODEL>(J + t * Y::JE, pts, chain_pair_bound(thr, gi.len, gj.len), thr, T2);
    }
    __syncthreads();
    double from[SD];
    load_state<SD>(Q, agent_i, from);
    unsigned n_exact = 0;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n_cand;
         k += n_warps) {
        double to[SD];
        load_state<SD>(q_to, k, to);
        ChainMove<MODEL> mv, unused;
        warp_chain_moves<MODEL>(cp, lane, from, to, from, to, mv, unused, false);
        mv.g = gi;
        const long long ns = mv.sch.n();
        bool hit = false;
        for (long long s0 = 0; !hit && s0 < ns; s0 += Y::CAP) {
            const int cnt = (int)(ns - s0 < Y::CAP ? ns - s0 : Y::CAP);
            fill_I<MODEL>(I, mv, s0, cnt, lane);
            __syncwarp();
            for (int j0 = 0; !hit && j0 < N - 1; j0 += 1023) {  // J indices carry 10 bits
                const int cj = N - 1 - j0 < 1023 ? N - 1 - j0 : 1023;
                hit = warp_sweep<MODEL>(I, cnt, J + (size_t)j0 * Y::JE, cj, queue, lane, n_exact);
            }
            __syncwarp();
        }
        if (lane == 0) out[k] = hit ? 1 : 0;
    }
    if (lane == 0 && stats && n_exact) atomicAdd(stats, (unsigned long long)n_exact);
}

__global__ void k_chain_configs_finish(int64_t n_cfg, const unsigned long long *__restrict__ first,
                                       uint8_t *__restrict__ out, int64_t *__restrict__ first_pair) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n_cfg;
         k += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long f = first[k];
        const bool hit = f != ~0ull;
        out[k] = hit ? 1 : 0;
        if (first_pair) first_pair[k] = hit ? (int64_t)f : -1;
    }
}

// ---- launchers ---------------------------------------------------------------------------------------
static inline int grid_threads(const LaunchCfg &lc, int64_t n, int per_sm) {
    int64_t want = (n + kThreads - 1) / kThreads;
    int64_t cap = (int64_t)lc.sm_count * per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

#define MRMP_CHAIN_SWITCH(cv, ...)                                                    \
    switch ((cv).model) {                                                             \
    case kModelLine2D: { constexpr int MODEL = kModelLine2D; __VA_ARGS__; } break;       \
    case kModelSnake2D: { constexpr int MODEL = kModelSnake2D; __VA_ARGS__; } break;     \
    case kModelArm33: { constexpr int MODEL = kModelArm33; __VA_ARGS__; } break;         \
    case kModelCapsule3D: { constexpr int MODEL = kModelCapsule3D; __VA_ARGS__; } break; \
    default: return cudaErrorNotSupported;                                            \
    }

cudaError_t launch_chain_connect_points(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                        const int32_t *agent, const double *q, uint8_t *out) {
    const int sm = seg_smem_bytes(cv.con_len);
    cudaError_t e = cudaSuccess;
    MRMP_CHAIN_SWITCH(cv, (e = set_smem(k_chain_connect_points<MODEL>, sm),
                           k_chain_connect_points<MODEL><<<grid_threads(lc, n, 4), kThreads, sm, lc.stream>>>(
                               cv, n, agent, q, out, sm > 0)));
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_chain_collide_static_pairs(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                              const int32_t *ai, const int32_t *aj,
                                              const double *qi, const double *qj, uint8_t *out) {
    const int sm = seg_smem_bytes(cv.col_len);
    cudaError_t e = cudaSuccess;
    MRMP_CHAIN_SWITCH(cv, (e = set_smem(k_chain_collide_static<MODEL>, sm),
                           k_chain_collide_static<MODEL><<<grid_threads(lc, n, 4), kThreads, sm, lc.stream>>>(
                               cv, n, ai, aj, qi, qj, out, sm > 0)));
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_chain_connect_edges(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                       const int32_t *agent, const double *qf, const double *qt,
                                       uint8_t *out) {
    const int sm = seg_smem_bytes(cv.con_len);
    cudaError_t e = cudaSuccess;
    MRMP_CHAIN_SWITCH(cv, (e = set_smem(k_chain_connect_edges<MODEL>, sm),
                           k_chain_connect_edges<MODEL><<<grid_threads(lc, n * 32, 6), kThreads, sm, lc.stream>>>(
                               cv, n, agent, qf, qt, out, sm > 0)));
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

template <int MODEL, class Src>
static cudaError_t launch_chain_collide_src(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                            const Src &src, unsigned long long *stats) {
    int seg = seg_smem_bytes(cv.col_len);
    const int tables = kPairWarps * Lay<MODEL>::WARP_BYTES;
    if (seg + tables > 200 * 1024) seg = 0;
    const int sm = seg + tables;
    cudaError_t e = set_smem(k_chain_collide<MODEL, Src>, sm);
    if (e != cudaSuccess) return e;
    int per_sm = (220 * 1024) / sm;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    int64_t want = (n + kPairWarps - 1) / kPairWarps;
    const int64_t cap = (int64_t)lc.sm_count * per_sm;
    if (want < 1) want = 1;
    const int grid = (int)(want < cap ? want : cap);
    unsigned long long *work = lc.stats + 3;  // ctx counter slot, zeroed in stream order
    e = cudaMemsetAsync(work, 0, sizeof(unsigned long long), lc.stream);
    if (e != cudaSuccess) return e;
    k_chain_collide<MODEL, Src><<<grid, kPairThreads, sm, lc.stream>>>(cv, n, src, stats, work, seg > 0,
                                                                      seg);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_chain_collide_pairs(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                       const int32_t *ai, const int32_t *aj, const double *qif,
                                       const double *qit, const double *qjf, const double *qjt,
                                       uint8_t *out, unsigned long long *stats) {
    MRMP_CHAIN_SWITCH(cv, {
        ListSrcN<Chain<MODEL>::SD> s{ai, aj, qif, qit, qjf, qjt, out};
        return launch_chain_collide_src<MODEL>(cv, lc, n, s, stats);
    });
    return cudaErrorNotSupported;
}

cudaError_t launch_chain_collide_edges_indexed(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                               const int32_t *ai, const int32_t *aj,
                                               const int32_t *vif, const int32_t *vit,
                                               const int32_t *vjf, const int32_t *vjt,
                                               const double *nodes, int agent0, int nv_stride,
                                               uint8_t *out, unsigned long long *stats) {
    MRMP_CHAIN_SWITCH(cv, {
        IndexedSrcN<Chain<MODEL>::SD> s{ai, aj, vif, vit, vjf, vjt, nodes, agent0, nv_stride, out};
        return launch_chain_collide_src<MODEL>(cv, lc, n, s, stats);
    });
    return cudaErrorNotSupported;
}

cudaError_t launch_chain_collide_cross(const CtxView &cv, const LaunchCfg &lc, int64_t n_rows,
                                       const int32_t *row_agent, const double *row_from,
                                       const double *row_to, int64_t n_cols,
                                       const int32_t *col_agent, const double *col_from,
                                       const double *col_to, const int32_t *col_vf,
                                       const int32_t *col_vt, const double *nodes, int agent0,
                                       int nv_stride, const CrossReq &rq, int out_stride,
                                       uint32_t *out, unsigned long long *stats) {
    MRMP_CHAIN_SWITCH(cv, {
        CrossSrcN<Chain<MODEL>::SD> s{row_agent, row_from, row_to, n_cols, col_agent, col_from, col_to,
                                      col_vf, col_vt, nodes, agent0, nv_stride, rq,
                                      (int64_t)out_stride, out};
        return launch_chain_collide_src<MODEL>(cv, lc, n_rows * n_cols, s, stats);
    });
    return cudaErrorNotSupported;
}

cudaError_t launch_chain_collide_configs(const CtxView &cv, const LaunchCfg &lc, int64_t n_cfg,
                                         const double *Qf, const double *Qt, uint8_t *out,
                                         int64_t *first_pair, unsigned long long *first_scratch,
                                         unsigned long long *stats) {
    cudaError_t e = cudaMemsetAsync(first_scratch, 0xff, sizeof(unsigned long long) * (size_t)n_cfg,
                                    lc.stream);
    if (e != cudaSuccess) return e;
    const int N = cv.n_agents;
    const int64_t pairs = (int64_t)N * (N - 1) / 2;
    if (pairs > 0) {
        e = cudaErrorNotSupported;
        MRMP_CHAIN_SWITCH(cv, {
            ConfigSrcN<Chain<MODEL>::SD> s{Qf, Qt, N, pairs, first_scratch};
            e = launch_chain_collide_src<MODEL>(cv, lc, n_cfg * pairs, s, stats);
        });
        if (e != cudaSuccess) return e;
    }
    k_chain_configs_finish<<<grid_threads(lc, n_cfg, 4), kThreads, 0, lc.stream>>>(
        n_cfg, first_scratch, out, first_pair);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_chain_collide_config_moves(const CtxView &cv, const LaunchCfg &lc, int agent_i,
                                              const double *Q, int64_t n_cand, const double *q_to,
                                              uint8_t *out, unsigned long long *stats) {
    const int N = cv.n_agents;
    const int j_cap = N - 1 > 1 ? N - 1 : 1;
    cudaError_t e = cudaErrorNotSupported;
    MRMP_CHAIN_SWITCH(cv, {
        using Y = Lay<MODEL>;
        int seg = seg_smem_bytes(cv.col_len);
        const size_t warp_bytes = (size_t)Y::CAP * Y::IE * 8 + (size_t)((Y::QCAP + 1) & ~1) * 4;
        const size_t tables = (size_t)j_cap * Y::JE * 8 + kPairWarps * warp_bytes;
        if (seg + tables > 200 * 1024) seg = 0;
        const size_t sm = seg + tables;
        if (sm > 220 * 1024) return cudaErrorNotSupported;
        e = set_smem(k_chain_config_moves<MODEL>, (int)sm);
        if (e != cudaSuccess) return e;
        int per_sm = (int)((220 * 1024) / sm);
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 8) per_sm = 8;
        int64_t want = (n_cand + kPairWarps - 1) / kPairWarps;
        const int64_t cap = (int64_t)lc.sm_count * per_sm;
        if (want < 1) want = 1;
        const int grid = (int)(want < cap ? want : cap);
        k_chain_config_moves<MODEL><<<grid, kPairThreads, sm, lc.stream>>>(
            cv, agent_i, Q, n_cand, q_to, out, stats, seg > 0, seg, j_cap);
        e = cudaSuccess;
    });
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

}  // namespace mrmp
