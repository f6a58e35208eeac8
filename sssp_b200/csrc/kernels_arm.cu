// kernels_arm.cu -- connect / collide kernels of the StateArm22 model (src/models/arm22.jl).
//
// An arm check is a loop over interpolation steps (connect: S postures; collide: S_i x S_j
// posture pairs x 4 link-link distances), 10^2 .. 10^5 primitive tests per check, so the unit
// of scheduling is a WARP per check:
//
//   connect(q_from,q_to,i)   lanes stride over the postures; one ballot per 32 postures gives
//                            the early exit of arm22.jl:73-82.
//   collide (6-arity, (C,q_to,i), configuration pairs, cross products)
//                            the postures of both sides are built ONCE per check into
//                            per-warp shared-memory tables (the reference rebuilds agent j's
//                            joints inside the inner loop, arm22.jl:125-131); lanes own the
//                            entries of the second table, the first table is read by
//                            broadcast.  A conservative midpoint bound (7 flops per link pair,
//                            same construction as the point-model cross kernel, DESIGN.md
//                            section 3) discards link pairs that are provably farther apart
//                            than safety_dist + 2^-20; the survivors are compacted through a
//                            per-warp queue (ballot + popc) so that the exact reference
//                            arithmetic (segseg_lt) always runs on full warps.  A hit anywhere
//                            ends the check (`any(...) && return true`, arm22.jl:134-142).
//
// The verdict of a check is an OR/AND over its steps, so evaluation order is free; everything
// that decides a verdict is the reference's arithmetic (arm_model.cuh, trig.cuh, geom.cuh).
#include "arm_model.cuh"
#include "point_model.cuh"

namespace mrmp {

extern __shared__ __align__(16) unsigned char smem_raw[];

constexpr int kThreads = 256;      // thread-per-check and connect kernels
constexpr int kPairWarps = 8;      // warps per CTA in the table kernels
constexpr int kPairThreads = 32 * kPairWarps;
constexpr int kArmCap = 64;        // postures per table tile (8 KB of tables per warp)
constexpr int kQueueCap = 160;     // < 32 left over + up to 128 new per sweep iteration
constexpr unsigned kFull = 0xffffffffu;

struct IEntry {  // posture of the first agent: joints b, c and the midpoints of its two links
    double bx, by, cx, cy, m1x, m1y, m2x, m2y;
};
struct JEntry {  // posture of the second agent: base, joints, cull bound of the agent pair
    double ax, ay, bx, by, cx, cy, G, pad;
};
__host__ __device__ constexpr int warp_table_bytes(int cap) {
    return cap * (int)(sizeof(IEntry) + sizeof(JEntry)) + kQueueCap * 4;
}

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

__device__ __forceinline__ ArmObs arm_obs(const CtxView &cv, const double *con) {
    ArmObs o;
    o.s = con;
    o.obs = cv.obs_off - cv.con_off;
    o.t2 = cv.t2_obs_off - cv.con_off;
    o.mp = cv.mp;
    o.n_obs = cv.n_obs;
    return o;
}
__device__ __forceinline__ ArmSeg arm_con(const CtxView &cv, const double *con) {
    return ArmSeg(con, cv.rads_off - cv.con_off, cv.base_off - cv.con_off);
}
__device__ __forceinline__ ArmSeg arm_col(const CtxView &cv, const double *col) {
    return ArmSeg(col, cv.crads_off - cv.col_off, cv.cbase_off - cv.col_off);
}

// conservative cull bound of an agent pair: a link pair whose midpoints are farther apart than
// sqrt(G) has distance > safety + 2^-20 (triangle inequality + (x+y+z)^2 <= 3(x^2+y^2+z^2));
// decided in plain FP64 with a 1e-9 relative pad, far above any rounding in the test itself.
__device__ __forceinline__ double arm_pair_bound(double safety, double ri, double rj) {
    const double tp = (safety > 0.0 ? safety : 0.0) + 9.5367431640625e-07;
    return 3.0 * (tp * tp + 0.25 * ri * ri + 0.25 * rj * rj) * (1.0 + 1e-9);
}

// ---- warp-cooperative move setup ----------------------------------------------------------
// The four diff_angles of a move (two turns, two components of D) run on four lanes at once;
// the schedule (continued-fraction test of Base.(:)) is evaluated on one lane and broadcast.
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(kFull, v, src); }
__device__ __forceinline__ long long shfl_ll(long long v, int src) {
    return __shfl_sync(kFull, v, src);
}

// set up n_moves (1 or 2) moves at once; move m uses lanes 4m .. 4m+3 for the angles
__device__ __forceinline__ void warp_arm_moves(const ArmCtx &ac, int lane, const double *qf0,
                                               const double *qt0, const double *qf1,
                                               const double *qt1, ArmMove &m0, ArmMove &m1,
                                               bool two) {
    double val = 0.0;
    if (lane < (two ? 8 : 4)) {
        const double *qf = lane < 4 ? qf0 : qf1, *qt = lane < 4 ? qt0 : qt1;
        const int k = lane & 1;
        val = (lane & 2) ? ddiv(diff_angles(qf[k], qt[k]), kPi64)  // component k of dist()
                         : diff_angles(qt[k], qf[k]);              // signed turn of joint k
    }
    double v[2];
    m0.dt1 = shfl_d(val, 0);
    m0.dt2 = shfl_d(val, 1);
    v[0] = shfl_d(val, 2);
    v[1] = shfl_d(val, 3);
    const double D0 = norm_val<2>(v);
    m1.dt1 = shfl_d(val, 4);
    m1.dt2 = shfl_d(val, 5);
    v[0] = shfl_d(val, 6);
    v[1] = shfl_d(val, 7);
    const double D1 = two ? norm_val<2>(v) : 0.0;
    ArmSched s;
    s.len = 0;
    s.rational = 0;
    if (lane < (two ? 2 : 1)) s = arm_schedule(ac.sr, ac.step, lane ? D1 : D0);
    m0.sch.len = shfl_ll(s.len, 0);
    m0.sch.rational = __shfl_sync(kFull, s.rational, 0);
    m1.sch.len = shfl_ll(s.len, 1);
    m1.sch.rational = __shfl_sync(kFull, s.rational, 1);
    m0.sch.num = m1.sch.num = ac.sr.num;
    m0.sch.den = m1.sch.den = ac.sr.den;
    m0.sch.step = m1.sch.step = ac.step;
    m0.sch.D = D0;
    m1.sch.D = D1;
    m0.t1 = qf0[0];
    m0.t2 = qf0[1];
    if (two) {
        m1.t1 = qf1[0];
        m1.t2 = qf1[1];
    }
}

// ---- thread-per-check kernels -----------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_arm_connect_points(CtxView cv, int64_t n, const int32_t *__restrict__ agent,
                     const double *__restrict__ q, uint8_t *__restrict__ out, int staged) {
    const double *con = stage_segment(cv, cv.con_off, cv.con_len, smem_raw, staged);
    const ArmObs ob = arm_obs(cv, con);
    const ArmSeg sg = arm_con(cv, con);
    const ArmCtx ac = arm_ctx(cv);
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
         k += (int64_t)gridDim.x * blockDim.x) {
        double p[2];
        load_state<2>(q, k, p);
        const int i = __ldg(agent + k);
        out[k] = arm_connect_point(ac, ob, p, sg.bx(i), sg.by(i), sg.rad(i)) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(kThreads)
k_arm_collide_static_pairs(CtxView cv, int64_t n, const int32_t *__restrict__ ai,
                           const int32_t *__restrict__ aj, const double *__restrict__ qi,
                           const double *__restrict__ qj, uint8_t *__restrict__ out, int staged) {
    const ArmSeg sg = arm_col(cv, stage_segment(cv, cv.col_off, cv.col_len, smem_raw, staged));
    const ArmCtx ac = arm_ctx(cv);
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
         k += (int64_t)gridDim.x * blockDim.x) {
        double a[2], b[2];
        load_state<2>(qi, k, a);
        load_state<2>(qj, k, b);
        const int i = __ldg(ai + k), j = __ldg(aj + k);
        out[k] = arm_collide_static(ac, a, b, sg.bx(i), sg.by(i), sg.rad(i), sg.bx(j), sg.by(j),
                                    sg.rad(j))
                     ? 1
                     : 0;
    }
}

// ---- connect(q_from, q_to, i): warp per check ------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_arm_connect_edges(CtxView cv, int64_t n, const int32_t *__restrict__ agent,
                    const double *__restrict__ qf, const double *__restrict__ qt,
                    uint8_t *__restrict__ out, int staged) {
    const double *con = stage_segment(cv, cv.con_off, cv.con_len, smem_raw, staged);
    const ArmObs ob = arm_obs(cv, con);
    const ArmSeg sg = arm_con(cv, con);
    const ArmCtx ac = arm_ctx(cv);
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n; k += n_warps) {
        double from[2], to[2];
        load_state<2>(qf, k, from);
        load_state<2>(qt, k, to);
        const int i = __ldg(agent + k);
        ArmMove mv, unused;
        warp_arm_moves(ac, lane, from, to, from, to, mv, unused, false);
        mv.ax = sg.bx(i);
        mv.ay = sg.by(i);
        mv.rad = sg.rad(i);
        bool ok = !(ac.max_dist == ac.max_dist && mv.sch.D > ac.max_dist);  // arm22.jl:58
        const double a[2] = {mv.ax, mv.ay};
        const long long ns = mv.sch.n();
        for (long long s0 = 0; ok && s0 < ns; s0 += 32) {
            const long long s = s0 + lane;
            bool bad = false;
            if (s < ns) {
                double b[2], c[2];
                mv.posture(s, b, c);
                bad = !arm_posture_free(ob, a, b, c, ac.safety, ac.safety_t2);
            }
            ok = !__any_sync(kFull, bad);
        }
        if (lane == 0) out[k] = ok ? 1 : 0;
    }
}

// ---- the table sweep ---------------------------------------------------------------------------
__device__ __forceinline__ bool queue_entry_hit(unsigned ent, const IEntry *I, double aix,
                                                double aiy, const JEntry *J, const ArmCtx &ac) {
    const IEntry ei = I[ent & 0xfffu];
    const JEntry ej = J[(ent >> 12) & 0xfffu];
    const double ai[2] = {aix, aiy}, bi[2] = {ei.bx, ei.by}, ci[2] = {ei.cx, ei.cy};
    const double aj[2] = {ej.ax, ej.ay}, bj[2] = {ej.bx, ej.by}, cj[2] = {ej.cx, ej.cy};
    return arm_link_pair_hit((int)(ent >> 24), ai, bi, ci, aj, bj, cj, ac.safety, ac.safety_t2);
}

// All nI x nJ posture pairs x 4 link pairs; true when any link pair is closer than safety_dist.
// Warp-synchronous; the tables must be visible to the warp (__syncwarp before the call).
__device__ __forceinline__ bool warp_sweep(const IEntry *I, int nI, double aix, double aiy,
                                           const JEntry *J, int nJ, unsigned *queue,
                                           const ArmCtx &ac, int lane, unsigned &n_exact,
                                           unsigned long long &n_bound) {
    int qlen = 0;
    const unsigned lt = (1u << lane) - 1u;
    for (int j0 = 0; j0 < nJ; j0 += 32) {
        const int jj = j0 + lane;
        const bool jv = jj < nJ;
        const int live_j = nJ - j0 < 32 ? nJ - j0 : 32;
        const JEntry ej = J[jv ? jj : 0];
        const double n1x = dmul(dadd(ej.ax, ej.bx), 0.5), n1y = dmul(dadd(ej.ay, ej.by), 0.5);
        const double n2x = dmul(dadd(ej.bx, ej.cx), 0.5), n2y = dmul(dadd(ej.by, ej.cy), 0.5);
        for (int ii = 0; ii < nI; ++ii) {
            const IEntry ei = I[ii];  // broadcast read
            double dx, dy;
            unsigned mask = 0;
            n_bound += 4u * (unsigned)live_j;  // link-pair midpoint bounds of this round (diagnostics)
            dx = dsub(ei.m1x, n1x), dy = dsub(ei.m1y, n1y);
            mask |= !(dadd(dmul(dx, dx), dmul(dy, dy)) > ej.G) ? 1u : 0u;
            dx = dsub(ei.m1x, n2x), dy = dsub(ei.m1y, n2y);
            mask |= !(dadd(dmul(dx, dx), dmul(dy, dy)) > ej.G) ? 2u : 0u;
            dx = dsub(ei.m2x, n1x), dy = dsub(ei.m2y, n1y);
            mask |= !(dadd(dmul(dx, dx), dmul(dy, dy)) > ej.G) ? 4u : 0u;
            dx = dsub(ei.m2x, n2x), dy = dsub(ei.m2y, n2y);
            mask |= !(dadd(dmul(dx, dx), dmul(dy, dy)) > ej.G) ? 8u : 0u;
            if (!jv) mask = 0;
            if (!__any_sync(kFull, mask != 0)) continue;
#pragma unroll
            for (int lp = 0; lp < 4; ++lp) {
                const bool on = (mask >> lp) & 1u;
                const unsigned bal = __ballot_sync(kFull, on);
                if (on) queue[qlen + __popc(bal & lt)] = (unsigned)ii | ((unsigned)jj << 12) | ((unsigned)lp << 24);
                qlen += __popc(bal);
            }
            __syncwarp();
            while (qlen >= 32) {
                qlen -= 32;
                const bool hit = queue_entry_hit(queue[qlen + lane], I, aix, aiy, J, ac);
                n_exact += 32;
                if (__any_sync(kFull, hit)) return true;
            }
            __syncwarp();
        }
    }
    bool hit = false;
    if (lane < qlen) hit = queue_entry_hit(queue[lane], I, aix, aiy, J, ac);
    n_exact += qlen;
    __syncwarp();
    return __any_sync(kFull, hit);
}

__device__ __forceinline__ void fill_I(IEntry *I, const ArmMove &mv, long long s0, int cnt,
                                       int lane) {
    for (int s = lane; s < cnt; s += 32) {
        double b[2], c[2];
        mv.posture(s0 + s, b, c);
        IEntry e;
        e.bx = b[0], e.by = b[1], e.cx = c[0], e.cy = c[1];
        e.m1x = dmul(dadd(mv.ax, b[0]), 0.5), e.m1y = dmul(dadd(mv.ay, b[1]), 0.5);
        e.m2x = dmul(dadd(b[0], c[0]), 0.5), e.m2y = dmul(dadd(b[1], c[1]), 0.5);
        I[s] = e;
    }
}
__device__ __forceinline__ void fill_J(JEntry *J, const ArmMove &mv, long long s0, int cnt,
                                       double G, int lane) {
    for (int s = lane; s < cnt; s += 32) {
        double b[2], c[2];
        mv.posture(s0 + s, b, c);
        JEntry e;
        e.ax = mv.ax, e.ay = mv.ay, e.bx = b[0], e.by = b[1], e.cx = c[0], e.cy = c[1];
        e.G = G, e.pad = 0.0;
        J[s] = e;
    }
}

// collide(qi_from, qi_to, qj_from, qj_to, i, j), arm22.jl:100-146, one warp
__device__ __forceinline__ bool warp_arm_collide_pair(const ArmCtx &ac, const ArmSeg &sg,
                                                      const double *qif, const double *qit,
                                                      const double *qjf, const double *qjt, int i,
                                                      int j, IEntry *I, JEntry *J, unsigned *queue,
                                                      int cap, int lane, unsigned &n_exact,
                                                      unsigned long long &n_bound) {
    ArmMove mi, mj;
    warp_arm_moves(ac, lane, qif, qit, qjf, qjt, mi, mj, true);
    mi.ax = sg.bx(i), mi.ay = sg.by(i), mi.rad = sg.rad(i);
    mj.ax = sg.bx(j), mj.ay = sg.by(j), mj.rad = sg.rad(j);
    // arms whose bases are farther apart than both reaches + margin can never touch
    {
        const double reach = 2.0 * fabs(mi.rad) + 2.0 * fabs(mj.rad) +
                             (ac.safety > 0.0 ? ac.safety : 0.0) + 9.5367431640625e-07;
        const double dx = mi.ax - mj.ax, dy = mi.ay - mj.ay;
        if (dx * dx + dy * dy > reach * reach * (1.0 + 1e-9)) return false;
    }
    const double G = arm_pair_bound(ac.safety, mi.rad, mj.rad);
    const long long ni = mi.sch.n(), nj = mj.sch.n();
    for (long long i0 = 0; i0 < ni; i0 += cap) {
        const int ci = (int)(ni - i0 < cap ? ni - i0 : cap);
        fill_I(I, mi, i0, ci, lane);
        for (long long j0 = 0; j0 < nj; j0 += cap) {
            const int cj = (int)(nj - j0 < cap ? nj - j0 : cap);
            if (i0 == 0 || nj > cap) fill_J(J, mj, j0, cj, G, lane);
            __syncwarp();
            if (warp_sweep(I, ci, mi.ax, mi.ay, J, cj, queue, ac, lane, n_exact, n_bound)) return true;
            __syncwarp();
        }
    }
    return false;
}

// ---- sources of 6-arity checks ----------------------------------------------------------------
struct PairCheck {
    int i, j;
    double qif[2], qit[2], qjf[2], qjt[2];
};

struct ListSrc {  // explicit list (mrmp_collide_pairs)
    const int32_t *ai, *aj;
    const double *qif, *qit, *qjf, *qjt;
    uint8_t *out;
    __device__ bool fetch(int64_t k, PairCheck &p) const {
        p.i = __ldg(ai + k);
        p.j = __ldg(aj + k);
        load_state<2>(qif, k, p.qif);
        load_state<2>(qit, k, p.qit);
        load_state<2>(qjf, k, p.qjf);
        load_state<2>(qjt, k, p.qjt);
        return true;
    }
    __device__ void emit(int64_t k, bool ran, bool hit) const { out[k] = (ran && hit) ? 1 : 0; }
    __device__ bool cbs() const { return false; }
};

struct IndexedSrc {  // roadmap edges by id (mrmp_collide_edges_indexed)
    const int32_t *ai, *aj, *vif, *vit, *vjf, *vjt;
    const double *nodes;
    int agent0, nv_stride;
    uint8_t *out;
    __device__ bool fetch(int64_t k, PairCheck &p) const {
        p.i = __ldg(ai + k);
        p.j = __ldg(aj + k);
        const int64_t bi = (int64_t)(p.i - agent0) * nv_stride, bj = (int64_t)(p.j - agent0) * nv_stride;
        load_state<2>(nodes, bi + __ldg(vif + k), p.qif);
        load_state<2>(nodes, bi + __ldg(vit + k), p.qit);
        load_state<2>(nodes, bj + __ldg(vjf + k), p.qjf);
        load_state<2>(nodes, bj + __ldg(vjt + k), p.qjt);
        return true;
    }
    __device__ void emit(int64_t k, bool ran, bool hit) const { out[k] = (ran && hit) ? 1 : 0; }
    __device__ bool cbs() const { return false; }
};

struct CrossSrc {  // rows x columns -> bits / counts / first hits (mrmp_collide_cross*, edge_conflicts)
    const int32_t *row_agent;
    const double *row_from, *row_to;
    int64_t n_cols;
    const int32_t *col_agent;
    const double *col_from, *col_to;     // explicit column states, or
    const int32_t *col_vf, *col_vt;      // ids into nodes
    const double *nodes;
    int agent0, nv_stride;
    CrossReq rq;
    int64_t out_stride;
    uint32_t *out;  // initialised by the caller (abi.cu: cross_run_dev)
    __device__ int64_t slot_of(int64_t c) const {
        return rq.col_group ? __ldg(rq.col_group + c) : (rq.group_size > 0 ? c / rq.group_size : c);
    }
    // cbs.jl:343-346: the COLUMN agent j is the first argument of both tests
    __device__ bool cbs() const { return rq.cbs != 0; }
    __device__ bool fetch(int64_t k, PairCheck &p) const {
        const int64_t r = k / n_cols, c = k - r * n_cols;
        const int ri = __ldg(row_agent + r), cj = __ldg(col_agent + c);
        if (ri == cj) return false;
        if (rq.row_key && !(__ldg(rq.col_key + c) > __ldg(rq.row_key + r))) return false;
        if (rq.mode == kCrossOr) {  // OR-reduced: a set bit cannot change any more
            const int64_t b = slot_of(c);
            if ((*(volatile uint32_t *)(out + r * out_stride + (b >> 5)) >> (b & 31)) & 1u) return false;
        }
        double *rf = rq.cbs ? p.qjf : p.qif, *rt = rq.cbs ? p.qjt : p.qit;
        double *cf = rq.cbs ? p.qif : p.qjf, *ct = rq.cbs ? p.qit : p.qjt;
        p.i = rq.cbs ? cj : ri;
        p.j = rq.cbs ? ri : cj;
        load_state<2>(row_from, r, rf);
        load_state<2>(row_to, r, rt);
        if (col_from) {
            load_state<2>(col_from, c, cf);
            load_state<2>(col_to, c, ct);
        } else {
            const int64_t bj = (int64_t)(cj - agent0) * nv_stride;
            load_state<2>(nodes, bj + __ldg(col_vf + c), cf);
            load_state<2>(nodes, bj + __ldg(col_vt + c), ct);
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

struct ConfigSrc {  // collide(Q_from, Q_to): arm22.jl:164-171, pairs i < j in row-major order
    const double *Qf, *Qt;
    int N;
    int64_t pairs;                // N(N-1)/2
    unsigned long long *first;    // per configuration: min i*N+j over colliding pairs
    __device__ bool fetch(int64_t k, PairCheck &p) const {
        const int64_t cfg = k / pairs;
        int64_t r = k - cfg * pairs;
        int i = 0;
        while (r >= N - 1 - i) {
            r -= N - 1 - i;
            ++i;
        }
        p.i = i;
        p.j = i + 1 + (int)r;
        // an earlier pair already collided: this one cannot become the first hit
        if (*(volatile unsigned long long *)(first + cfg) < (unsigned long long)(p.i * N + p.j))
            return false;
        const int64_t base = cfg * N;
        load_state<2>(Qf, base + p.i, p.qif);
        load_state<2>(Qt, base + p.i, p.qit);
        load_state<2>(Qf, base + p.j, p.qjf);
        load_state<2>(Qt, base + p.j, p.qjt);
        return true;
    }
    __device__ void emit(int64_t k, bool ran, bool hit) const {
        if (ran && hit) {
            const int64_t cfg = k / pairs;
            int64_t r = k - cfg * pairs;
            int i = 0;
            while (r >= N - 1 - i) {
                r -= N - 1 - i;
                ++i;
            }
            atomicMin(first + cfg, (unsigned long long)(i * N + i + 1 + (int)r));
        }
    }
    __device__ bool cbs() const { return false; }
};

template <class Src>
__global__ void __launch_bounds__(kPairThreads, 3)
k_arm_collide(CtxView cv, int64_t n, Src src, unsigned long long *stats,
              unsigned long long *work, int staged, int seg_bytes, int cap) {
    const ArmSeg sg = arm_col(cv, stage_segment(cv, cv.col_off, cv.col_len, smem_raw, staged));
    const ArmCtx ac = arm_ctx(cv);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *wb = smem_raw + seg_bytes + (size_t)warp * warp_table_bytes(cap);
    IEntry *I = reinterpret_cast<IEntry *>(wb);
    JEntry *J = reinterpret_cast<JEntry *>(wb + cap * sizeof(IEntry));
    unsigned *queue = reinterpret_cast<unsigned *>(wb + cap * (sizeof(IEntry) + sizeof(JEntry)));
    unsigned n_exact = 0;
    unsigned long long n_bound = 0;
    // checks differ in cost by orders of magnitude (S_i x S_j posture pairs, early exits):
    // warps pull the next check from a counter instead of striding
    for (;;) {
        unsigned long long kk = 0;
        if (lane == 0) kk = atomicAdd(work, 1ull);
        const int64_t k = (int64_t)__shfl_sync(kFull, kk, 0);
        if (k >= n) break;
        PairCheck p;
        const bool run = src.fetch(k, p);  // warp-uniform
        bool hit = false;
        if (run && src.cbs())  // collide(q_j_to, q_i_to, j, i), cbs.jl:345 / arm22.jl:148-162
            hit = arm_collide_static(ac, p.qit, p.qjt, sg.bx(p.i), sg.by(p.i), sg.rad(p.i),
                                     sg.bx(p.j), sg.by(p.j), sg.rad(p.j));
        if (run && !hit)
            hit = warp_arm_collide_pair(ac, sg, p.qif, p.qit, p.qjf, p.qjt, p.i, p.j, I, J, queue,
                                        cap, lane, n_exact, n_bound);
        if (lane == 0) src.emit(k, run, hit);
        __syncwarp();
    }
    if (lane == 0 && stats && n_exact) atomicAdd(stats, (unsigned long long)n_exact);
    if (lane == 0 && stats && n_bound) atomicAdd(stats + 3, n_bound);  // ctx stats slot 4
}

// ---- collide(C, q_to, i): arm22.jl:173-205, warp per candidate --------------------------------
// The static postures of the other agents are the same for every candidate: built once per
// CTA into a shared J table (entry per agent j != i, in ascending j).
__global__ void __launch_bounds__(kPairThreads)
k_arm_config_moves(CtxView cv, int agent_i, const double *__restrict__ Q, int64_t n_cand,
                   const double *__restrict__ q_to, uint8_t *__restrict__ out,
                   unsigned long long *stats, int staged, int seg_bytes, int j_cap) {
    const ArmSeg sg = arm_col(cv, stage_segment(cv, cv.col_off, cv.col_len, smem_raw, staged));
    const ArmCtx ac = arm_ctx(cv);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = cv.n_agents;
    JEntry *J = reinterpret_cast<JEntry *>(smem_raw + seg_bytes);
    unsigned char *wb = smem_raw + seg_bytes + (size_t)j_cap * sizeof(JEntry) +
                        (size_t)warp * (kArmCap * sizeof(IEntry) + kQueueCap * 4);
    IEntry *I = reinterpret_cast<IEntry *>(wb);
    unsigned *queue = reinterpret_cast<unsigned *>(wb + kArmCap * sizeof(IEntry));
    const double ri = sg.rad(agent_i), aix = sg.bx(agent_i), aiy = sg.by(agent_i);
    for (int t = threadIdx.x; t < N - 1; t += blockDim.x) {
        const int j = t < agent_i ? t : t + 1;
        double qj[2], b[2], c[2];
        load_state<2>(Q, j, qj);
        arm_joints(qj[0], qj[1], sg.bx(j), sg.by(j), sg.rad(j), b, c);
        JEntry e;
        e.ax = sg.bx(j), e.ay = sg.by(j), e.bx = b[0], e.by = b[1], e.cx = c[0], e.cy = c[1];
        e.G = arm_pair_bound(ac.safety, ri, sg.rad(j)), e.pad = 0.0;
        J[t] = e;
    }
    __syncthreads();
    double from[2];
    load_state<2>(Q, agent_i, from);
    unsigned n_exact = 0;
    unsigned long long n_bound = 0;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n_cand;
         k += n_warps) {
        double to[2];
        load_state<2>(q_to, k, to);
        ArmMove mv, unused;
        warp_arm_moves(ac, lane, from, to, from, to, mv, unused, false);
        mv.ax = aix, mv.ay = aiy, mv.rad = ri;
        const long long ns = mv.sch.n();
        bool hit = false;
        for (long long s0 = 0; !hit && s0 < ns; s0 += kArmCap) {
            const int cnt = (int)(ns - s0 < kArmCap ? ns - s0 : kArmCap);
            fill_I(I, mv, s0, cnt, lane);
            __syncwarp();
            hit = warp_sweep(I, cnt, aix, aiy, J, N - 1, queue, ac, lane, n_exact, n_bound);
            __syncwarp();
        }
        if (lane == 0) out[k] = hit ? 1 : 0;
    }
    if (lane == 0 && stats && n_exact) atomicAdd(stats, (unsigned long long)n_exact);
}

// first[cfg] (min pair index or ~0) -> verdict byte and the ABI's first_pair (-1 = none)
__global__ void k_arm_configs_finish(int64_t n_cfg, const unsigned long long *__restrict__ first,
                                     uint8_t *__restrict__ out, int64_t *__restrict__ first_pair) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n_cfg;
         k += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long f = first[k];
        const bool hit = f != ~0ull;
        out[k] = hit ? 1 : 0;
        if (first_pair) first_pair[k] = hit ? (int64_t)f : -1;
    }
}

// ---- launchers ----------------------------------------------------------------------------------
static inline int grid_threads(const LaunchCfg &lc, int64_t n, int per_sm) {
    int64_t want = (n + kThreads - 1) / kThreads;
    int64_t cap = (int64_t)lc.sm_count * per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

cudaError_t launch_arm_connect_points(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                      const int32_t *agent, const double *q, uint8_t *out) {
    const int sm = seg_smem_bytes(cv.con_len);
    cudaError_t e = set_smem(k_arm_connect_points, sm);
    if (e != cudaSuccess) return e;
    k_arm_connect_points<<<grid_threads(lc, n, 4), kThreads, sm, lc.stream>>>(cv, n, agent, q, out,
                                                                              sm > 0);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_arm_collide_static_pairs(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                            const int32_t *ai, const int32_t *aj,
                                            const double *qi, const double *qj, uint8_t *out) {
    const int sm = seg_smem_bytes(cv.col_len);
    cudaError_t e = set_smem(k_arm_collide_static_pairs, sm);
    if (e != cudaSuccess) return e;
    k_arm_collide_static_pairs<<<grid_threads(lc, n, 4), kThreads, sm, lc.stream>>>(
        cv, n, ai, aj, qi, qj, out, sm > 0);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_arm_connect_edges(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                     const int32_t *agent, const double *qf, const double *qt,
                                     uint8_t *out) {
    const int sm = seg_smem_bytes(cv.con_len);
    cudaError_t e = set_smem(k_arm_connect_edges, sm);
    if (e != cudaSuccess) return e;
    // one warp per check: n*32 threads wanted
    k_arm_connect_edges<<<grid_threads(lc, n * 32, 6), kThreads, sm, lc.stream>>>(
        cv, n, agent, qf, qt, out, sm > 0);
    count_launch();
    return cudaGetLastError();
}

template <class Src>
static cudaError_t launch_arm_collide_src(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                          const Src &src, unsigned long long *stats) {
    // bulk batches: 8 warps per CTA with 64-posture table tiles (3 CTAs = 24 warps per SM);
    // small batches are bound by their slowest check: fewer warps per CTA spread them over more
    // SMs, and 128-posture tiles hold a whole move (no table refills)
    int warps = kPairWarps;
    while (warps > 1 && (n + warps - 1) / warps < (int64_t)lc.sm_count * 3) warps >>= 1;
    const int cap = warps <= 4 ? 2 * kArmCap : kArmCap;
    int seg = seg_smem_bytes(cv.col_len);
    const int tables = warps * warp_table_bytes(cap);
    if (seg + tables > 200 * 1024) seg = 0;
    const int sm = seg + tables;
    cudaError_t e = set_smem(k_arm_collide<Src>, sm);
    if (e != cudaSuccess) return e;
    int per_sm = (220 * 1024) / sm;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 12) per_sm = 12;
    int64_t want = (n + warps - 1) / warps;
    const int64_t cap_ctas = (int64_t)lc.sm_count * per_sm;
    if (want < 1) want = 1;
    const int grid = (int)(want < cap_ctas ? want : cap_ctas);
    // work counter: the ctx's reserved counter slot, zeroed in stream order before the launch
    unsigned long long *work = lc.stats + 3;
    e = cudaMemsetAsync(work, 0, sizeof(unsigned long long), lc.stream);
    if (e != cudaSuccess) return e;
    k_arm_collide<Src><<<grid, 32 * warps, sm, lc.stream>>>(cv, n, src, stats, work, seg > 0, seg, cap);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_arm_collide_pairs(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                     const int32_t *ai, const int32_t *aj, const double *qif,
                                     const double *qit, const double *qjf, const double *qjt,
                                     uint8_t *out, unsigned long long *stats) {
    ListSrc s{ai, aj, qif, qit, qjf, qjt, out};
    return launch_arm_collide_src(cv, lc, n, s, stats);
}

cudaError_t launch_arm_collide_edges_indexed(const CtxView &cv, const LaunchCfg &lc, int64_t n,
                                             const int32_t *ai, const int32_t *aj,
                                             const int32_t *vif, const int32_t *vit,
                                             const int32_t *vjf, const int32_t *vjt,
                                             const double *nodes, int agent0, int nv_stride,
                                             uint8_t *out, unsigned long long *stats) {
    IndexedSrc s{ai, aj, vif, vit, vjf, vjt, nodes, agent0, nv_stride, out};
    return launch_arm_collide_src(cv, lc, n, s, stats);
}

cudaError_t launch_arm_collide_cross(const CtxView &cv, const LaunchCfg &lc, int64_t n_rows,
                                     const int32_t *row_agent, const double *row_from,
                                     const double *row_to, int64_t n_cols,
                                     const int32_t *col_agent, const double *col_from,
                                     const double *col_to, const int32_t *col_vf,
                                     const int32_t *col_vt, const double *nodes, int agent0,
                                     int nv_stride, const CrossReq &rq, int out_stride,
                                     uint32_t *out, unsigned long long *stats) {
    CrossSrc s{row_agent, row_from, row_to, n_cols, col_agent, col_from, col_to, col_vf, col_vt,
               nodes, agent0, nv_stride, rq, (int64_t)out_stride, out};
    return launch_arm_collide_src(cv, lc, n_rows * n_cols, s, stats);
}

cudaError_t launch_arm_collide_configs(const CtxView &cv, const LaunchCfg &lc, int64_t n_cfg,
                                       const double *Qf, const double *Qt, uint8_t *out,
                                       int64_t *first_pair, unsigned long long *first_scratch,
                                       unsigned long long *stats) {
    cudaError_t e = cudaMemsetAsync(first_scratch, 0xff, sizeof(unsigned long long) * (size_t)n_cfg,
                                    lc.stream);
    if (e != cudaSuccess) return e;
    const int N = cv.n_agents;
    const int64_t pairs = (int64_t)N * (N - 1) / 2;
    if (pairs > 0) {
        ConfigSrc s{Qf, Qt, N, pairs, first_scratch};
        e = launch_arm_collide_src(cv, lc, n_cfg * pairs, s, stats);
        if (e != cudaSuccess) return e;
    }
    k_arm_configs_finish<<<grid_threads(lc, n_cfg, 4), kThreads, 0, lc.stream>>>(
        n_cfg, first_scratch, out, first_pair);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_arm_collide_config_moves(const CtxView &cv, const LaunchCfg &lc, int agent_i,
                                            const double *Q, int64_t n_cand, const double *q_to,
                                            uint8_t *out, unsigned long long *stats) {
    const int N = cv.n_agents;
    const int j_cap = N - 1 > 1 ? N - 1 : 1;
    if (j_cap > 0xfff) return cudaErrorNotSupported;  // queue entries carry 12-bit table indices
    int seg = seg_smem_bytes(cv.col_len);
    const int tables = j_cap * (int)sizeof(JEntry) +
                       kPairWarps * (kArmCap * (int)sizeof(IEntry) + kQueueCap * 4);
    if (seg + tables > 200 * 1024) seg = 0;
    const int sm = seg + tables;
    if (sm > 220 * 1024) return cudaErrorNotSupported;
    cudaError_t e = set_smem(k_arm_config_moves, sm);
    if (e != cudaSuccess) return e;
    int per_sm = (220 * 1024) / sm;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    int64_t want = (n_cand + kPairWarps - 1) / kPairWarps;
    const int64_t cap = (int64_t)lc.sm_count * per_sm;
    if (want < 1) want = 1;
    const int grid = (int)(want < cap ? want : cap);
    k_arm_config_moves<<<grid, kPairThreads, sm, lc.stream>>>(cv, agent_i, Q, n_cand, q_to, out,
                                                              stats, seg > 0, seg, j_cap);
    count_launch();
    return cudaGetLastError();
}

}  // namespace mrmp
