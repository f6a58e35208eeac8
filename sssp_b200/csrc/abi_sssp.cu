// abi_sssp.cu -- C ABI of the SSSP hot loops (include/mrmp_b200.h, "SSSP" section):
// device-resident growing roadmaps, the fused node expansion, and the small model helpers
// (dist, get_mid_status, nearest vertex) the host-side search loops call.
#include "abi_internal.cuh"

#include <chrono>

struct mrmp_sssp {
    mrmp_ctx *ctx = nullptr;
    int cap = 0;                 // vertices per agent the node table can hold
    double *nodes = nullptr;     // device [N][cap][d]
    int32_t *nv_d = nullptr;     // device [N]
    std::vector<int32_t> nv_h;   // host mirror
    unsigned char *io_h = nullptr, *io_d = nullptr;  // pinned / device staging of one expansion
    size_t io_bytes = 0;
    int seq = 0;                 // completion counter of the fused expansion (hdr[3])
    double last_total_us = 0.0, last_wait_us = 0.0;  // host time of the last mrmp_sssp_expand
};

static int sssp_io(mrmp_sssp *s, size_t bytes) {
    if (bytes <= s->io_bytes) return MRMP_OK;
    mrmp_ctx *c = s->ctx;
    CU(cudaStreamSynchronize(c->s_k));
    if (s->io_h) CU(cudaFreeHost(s->io_h));
    if (s->io_d) CU(cudaFree(s->io_d));
    s->io_h = s->io_d = nullptr;
    s->io_bytes = 0;
    size_t nb = 1 << 16;
    while (nb < bytes) nb *= 2;
    CU(cudaHostAlloc((void **)&s->io_h, nb, cudaHostAllocMapped));  // the fused kernel reads/writes it
    CU(cudaMalloc((void **)&s->io_d, nb));
    s->io_bytes = nb;
    return MRMP_OK;
}

// grow the node tables to hold at least `need` vertices per agent (keeps the contents)
static int sssp_reserve(mrmp_sssp *s, int need) {
    if (need <= s->cap) return MRMP_OK;
    mrmp_ctx *c = s->ctx;
    int ncap = s->cap;
    while (ncap < need) ncap *= 2;
    const size_t row_old = sizeof(double) * c->d * (size_t)s->cap;
    const size_t row_new = sizeof(double) * c->d * (size_t)ncap;
    double *nn = nullptr;
    CU(cudaMalloc((void **)&nn, row_new * c->n_agents));
    cudaError_t e = cudaMemcpy2DAsync(nn, row_new, s->nodes, row_old, row_old, c->n_agents,
                                      cudaMemcpyDeviceToDevice, c->s_k);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->s_k);
    if (e != cudaSuccess) {
        cudaFree(nn);
        return fail(MRMP_ERR_CUDA, "sssp_reserve: %s", cudaGetErrorString(e));
    }
    cudaFree(s->nodes);
    s->nodes = nn;
    s->cap = ncap;
    return MRMP_OK;
}

extern "C" int mrmp_sssp_create(mrmp_ctx *c, int nv_cap, mrmp_sssp **out) {
    NEED(c && out, "bad ctx / out");
    *out = nullptr;
    NEED(nv_cap >= 2 && nv_cap <= (1 << 24), "nv_cap out of range");
    CU(cudaSetDevice(c->device));
    mrmp_sssp *s = new (std::nothrow) mrmp_sssp();
    if (!s) return fail(MRMP_ERR_NOMEM, "out of host memory");
    s->ctx = c;
    s->cap = (nv_cap + 1) & ~1;
    s->nv_h.assign(c->n_agents, 0);
    cudaError_t e = cudaMalloc((void **)&s->nodes, sizeof(double) * c->d * (size_t)s->cap * c->n_agents);
    if (e == cudaSuccess) e = cudaMalloc((void **)&s->nv_d, sizeof(int32_t) * c->n_agents);
    if (e == cudaSuccess) e = cudaMemset(s->nv_d, 0, sizeof(int32_t) * c->n_agents);
    if (e != cudaSuccess) {
        mrmp_sssp_destroy(s);
        return fail(MRMP_ERR_CUDA, "sssp alloc: %s", cudaGetErrorString(e));
    }
    *out = s;
    return MRMP_OK;
}

extern "C" void mrmp_sssp_destroy(mrmp_sssp *s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->s_k);
    cudaFree(s->nodes);
    cudaFree(s->nv_d);
    if (s->io_h) cudaFreeHost(s->io_h);
    cudaFree(s->io_d);
    delete s;
}

extern "C" int mrmp_sssp_set_roadmap(mrmp_sssp *s, int agent, int nv, const double *nodes) {
    NEED(s && nodes, "bad sssp / nodes");
    mrmp_ctx *c = s->ctx;
    NEED((unsigned)agent < (unsigned)c->n_agents, "agent out of range");
    NEED(nv >= 1, "nv must be positive");
    CU(cudaSetDevice(c->device));
    int rc = sssp_reserve(s, nv);
    if (rc) return rc;
    const size_t bytes = sizeof(double) * c->d * (size_t)nv;
    rc = sssp_io(s, bytes + 16);
    if (rc) return rc;
    memcpy(s->io_h, nodes, bytes);
    int32_t *nvp = (int32_t *)(s->io_h + align16(bytes));
    *nvp = nv;
    CU(cudaMemcpyAsync(s->nodes + (size_t)agent * s->cap * c->d, s->io_h, bytes,
                       cudaMemcpyHostToDevice, c->s_k));
    CU(cudaMemcpyAsync(s->nv_d + agent, nvp, 4, cudaMemcpyHostToDevice, c->s_k));
    CU(cudaStreamSynchronize(c->s_k));
    s->nv_h[agent] = nv;
    return MRMP_OK;
}

extern "C" int mrmp_sssp_nv(mrmp_sssp *s, int agent, int *nv_out) {
    NEED(s && nv_out, "bad arguments");
    NEED((unsigned)agent < (unsigned)s->ctx->n_agents, "agent out of range");
    *nv_out = s->nv_h[agent];
    return MRMP_OK;
}

extern "C" int mrmp_sssp_get_nodes(mrmp_sssp *s, int agent, double *nodes_out, int cap) {
    NEED(s && nodes_out, "bad arguments");
    mrmp_ctx *c = s->ctx;
    NEED((unsigned)agent < (unsigned)c->n_agents, "agent out of range");
    const int nv = s->nv_h[agent];
    if (cap < nv) return fail(MRMP_ERR_CAPACITY, "nodes_out holds %d < %d vertices", cap, nv);
    if (nv == 0) return MRMP_OK;
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(nodes_out, s->nodes + (size_t)agent * s->cap * c->d,
                       sizeof(double) * c->d * (size_t)nv, cudaMemcpyDeviceToHost, c->s_k));
    CU(cudaStreamSynchronize(c->s_k));
    return MRMP_OK;
}

extern "C" int mrmp_sssp_expand(mrmp_sssp *s, int agent, int v_from, int n_samples,
                                const double *q_rand, uint64_t seed, uint64_t t0,
                                int steering_depth, double epsilon, double min_dist_thread,
                                const int32_t *Q_ids, int n_old_nbrs, const int32_t *old_nbrs,
                                int slow_collide, int32_t *n_new_out, double *q_new_out,
                                int32_t *words_out, uint32_t *out_bits, uint32_t *in_bits,
                                int64_t bits_cap_words, int32_t *n_succ_out, int32_t *succ_ids,
                                uint8_t *succ_collide) {
    NEED(s && Q_ids && n_new_out && q_new_out && words_out && n_succ_out && succ_ids && succ_collide,
         "NULL argument");
    mrmp_ctx *c = s->ctx;
    const int N = c->n_agents, d = c->d;
    NEED((unsigned)agent < (unsigned)N, "agent out of range");
    NEED(n_samples >= 1 && n_samples <= kMaxExpand, "n_samples must be in 1..64");
    NEED(steering_depth >= 0 && n_old_nbrs >= 0 && (n_old_nbrs == 0 || old_nbrs), "bad arguments");
    const int nv0 = s->nv_h[agent];
    NEED(nv0 >= 1 && (unsigned)v_from < (unsigned)nv0, "v_from out of range (roadmap not set?)");
    for (int j = 0; j < N; ++j)
        if ((unsigned)Q_ids[j] >= (unsigned)s->nv_h[j])
            return fail(MRMP_ERR_INVALID, "Q_ids[%d] = %d out of range", j, Q_ids[j]);
    if (Q_ids[agent] != v_from) return fail(MRMP_ERR_INVALID, "Q_ids[agent] must equal v_from");
    for (int k = 0; k < n_old_nbrs; ++k)
        if ((unsigned)old_nbrs[k] >= (unsigned)nv0)
            return fail(MRMP_ERR_INVALID, "old_nbrs[%d] out of range", k);
    const int K = n_samples;
    const int words = (nv0 + K + 31) / 32;
    *words_out = words;
    if (bits_cap_words < (int64_t)K * words || !out_bits || !in_bits)
        return fail(MRMP_ERR_CAPACITY, "out_bits / in_bits need %d x %d words each", K, words);
    const auto t_call = std::chrono::steady_clock::now();
    CU(cudaSetDevice(c->device));
    int rc = sssp_reserve(s, nv0 + K);
    if (rc) return rc;

    // ---- staging layout: [inputs | outputs | device-only scratch]
    const int cap_s = n_old_nbrs + K + 1;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t o = off;
        off = align16(off + bytes);
        return o;
    };
    const size_t o_qids = take(4 * (size_t)N), o_old = take(4 * (size_t)n_old_nbrs),
                 o_qrand = take(q_rand ? sizeof(double) * d * (size_t)K : 0);
    const size_t in_end = off;
    const size_t o_hdr = take(16), o_qnew = take(sizeof(double) * d * (size_t)K),
                 o_bits = take(4 * 2 * (size_t)K * words), o_succ = take(4 * (size_t)cap_s),
                 o_sout = take((size_t)cap_s);
    const size_t out_end = off;
    const size_t o_qbuf = take(sizeof(double) * d * (size_t)N),
                 o_cand = take(sizeof(double) * d * (size_t)cap_s);
    const size_t cfg_b = slow_collide ? sizeof(double) * d * (size_t)N * cap_s : 0;
    const size_t o_cf = take(cfg_b), o_ct = take(cfg_b), o_first = take(slow_collide ? 8 * (size_t)cap_s : 0);
    rc = sssp_io(s, off);
    if (rc) return rc;
    memcpy(s->io_h + o_qids, Q_ids, 4 * (size_t)N);
    if (n_old_nbrs) memcpy(s->io_h + o_old, old_nbrs, 4 * (size_t)n_old_nbrs);
    if (q_rand) memcpy(s->io_h + o_qrand, q_rand, sizeof(double) * d * (size_t)K);
    cudaStream_t st = c->s_k;

    SsspView view{s->nodes, s->nv_d, s->cap};
    SsspExpand a{};
    a.agent = agent;
    a.v_from = v_from;
    a.K = K;
    a.depth = steering_depth;
    a.eps = epsilon;
    a.eps_r2 = sq_le_threshold(epsilon);
    a.thr = min_dist_thread;
    a.thr_r2 = sq_le_threshold(min_dist_thread);
    a.seed = seed;
    a.t0 = t0;
    a.host_samples = q_rand ? 1 : 0;
    a.words = words;
    a.n_old = n_old_nbrs;
    a.slow_collide = slow_collide ? 1 : 0;
    SsspWs ws{};
    unsigned char *b = s->io_d;
    ws.q_rand = (const double *)(b + o_qrand);
    ws.Q_ids = (const int32_t *)(b + o_qids);
    ws.old_nbrs = (const int32_t *)(b + o_old);
    ws.hdr = (int32_t *)(b + o_hdr);
    ws.q_new = (double *)(b + o_qnew);
    ws.out_bits = (uint32_t *)(b + o_bits);
    ws.in_bits = ws.out_bits + (size_t)K * words;
    ws.succ_ids = (int32_t *)(b + o_succ);
    ws.succ_out = (uint8_t *)(b + o_sout);
    ws.Qbuf = (double *)(b + o_qbuf);
    ws.cand = (double *)(b + o_cand);
    ws.cfg_from = (double *)(b + o_cf);
    ws.cfg_to = (double *)(b + o_ct);
    ws.cfg_first = (int64_t *)(b + o_first);
    // search-sized roadmaps of the point models: one launch that works on the mapped pinned
    // block directly; otherwise the multi-kernel path with explicit copies
    const bool fused = c->model <= MRMP_MODEL_POINT3D && K <= kFusedMaxK && N <= kFusedMaxAgents &&
                       n_old_nbrs <= kFusedMaxOld &&
                       (int64_t)K * 2 * (nv0 + K) <= (1 << 16) &&
                       sssp_fused_smem(c->cv, a) <= 160 * 1024;
    bool polled = false;
    auto t_wait = std::chrono::steady_clock::now();
    if (fused) {
        unsigned char *hb = nullptr;
        CU(cudaHostGetDevicePointer((void **)&hb, s->io_h, 0));
        a.seq = ++s->seq;
        if (a.seq == 0) a.seq = s->seq = 1;
        volatile int32_t *flag = (volatile int32_t *)(s->io_h + o_hdr) + 3;
        *flag = 0;
        SsspFusedIn in;
        memcpy(in.q_ids, Q_ids, 4 * (size_t)N);
        if (n_old_nbrs) memcpy(in.old_nbrs, old_nbrs, 4 * (size_t)n_old_nbrs);
        if (q_rand) memcpy(in.q_rand, q_rand, sizeof(double) * d * (size_t)K);
        CU(launch_sssp_expand_fused(c->cv, lcfg(c, st), view, a, in, (int32_t *)(hb + o_hdr),
                                    (double *)(hb + o_qnew), (uint32_t *)(hb + o_bits),
                                    (int32_t *)(hb + o_succ), (uint8_t *)(hb + o_sout)));
        t_wait = std::chrono::steady_clock::now();
        // The kernel takes ~20 us and writes its results into this mapped block: poll its
        // completion flag (a stream wait costs several microseconds of wake-up latency on top);
        // after ~2 ms without it fall back to the stream wait, which also surfaces launch errors.
        for (int spin = 0; spin < 400000; ++spin) {
            if (*flag == a.seq) {
                polled = true;
                break;
            }
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
    } else {
        CU(cudaMemcpyAsync(s->io_d, s->io_h, in_end, cudaMemcpyHostToDevice, st));
        CU(launch_sssp_expand(c->cv, lcfg(c, st), view, a, ws, nv0));
        CU(cudaMemcpyAsync(s->io_h + in_end, s->io_d + in_end, out_end - in_end,
                           cudaMemcpyDeviceToHost, st));
    }
    if (!polled) {
        if (!fused) t_wait = std::chrono::steady_clock::now();
        CU(cudaStreamSynchronize(st));
    }

    const int32_t *hdr = (const int32_t *)(s->io_h + o_hdr);
    const int n_new = hdr[0], n_succ = hdr[1];
    s->nv_h[agent] = nv0 + n_new;
    *n_new_out = n_new;
    *n_succ_out = n_succ;
    memcpy(q_new_out, s->io_h + o_qnew, sizeof(double) * d * (size_t)n_new);
    memcpy(out_bits, s->io_h + o_bits, 4 * (size_t)K * words);
    memcpy(in_bits, s->io_h + o_bits + 4 * (size_t)K * words, 4 * (size_t)K * words);
    memcpy(succ_ids, s->io_h + o_succ, 4 * (size_t)n_succ);
    memcpy(succ_collide, s->io_h + o_sout, (size_t)n_succ);
    const auto t_end = std::chrono::steady_clock::now();
    s->last_total_us = std::chrono::duration<double, std::micro>(t_end - t_call).count();
    s->last_wait_us = std::chrono::duration<double, std::micro>(t_end - t_wait).count();
    return MRMP_OK;
}

/* diagnostics: host time of the last mrmp_sssp_expand inside the library (total, and the part
 * spent waiting for the kernel after the launch call returned), in microseconds */
extern "C" int mrmp_sssp_last_call_us(mrmp_sssp *s, double *total_us, double *wait_us) {
    NEED(s, "bad sssp");
    if (total_us) *total_us = s->last_total_us;
    if (wait_us) *wait_us = s->last_wait_us;
    return MRMP_OK;
}

// ---- model helpers -----------------------------------------------------------------------------
static int state_op(mrmp_ctx *c, int op, int64_t n, const double *q1, const double *q2,
                    double *out) {
    NEED(c && n >= 0, "bad ctx / n");
    if (n == 0) return MRMP_OK;
    NEED(q1 && q2 && out, "NULL buffer");
    CU(cudaSetDevice(c->device));
    const size_t sb = sizeof(double) * c->d * (size_t)n;
    const size_t ob = op == 0 ? sizeof(double) * (size_t)n : sb;
    const size_t o2 = align16(sb), o3 = o2 + align16(sb);
    int rc = ensure_small(c, o3 + align16(ob));
    if (rc) return rc;
    memcpy(c->small_h, q1, sb);
    memcpy(c->small_h + o2, q2, sb);
    CU(cudaMemcpyAsync(c->small_d, c->small_h, o3, cudaMemcpyHostToDevice, c->s_k));
    CU(launch_state_op(c->cv, lcfg(c, c->s_k), op, n, (const double *)c->small_d,
                       (const double *)(c->small_d + o2), (double *)(c->small_d + o3)));
    CU(cudaMemcpyAsync(c->small_h + o3, c->small_d + o3, ob, cudaMemcpyDeviceToHost, c->s_k));
    CU(cudaStreamSynchronize(c->s_k));
    memcpy(out, c->small_h + o3, ob);
    return MRMP_OK;
}

extern "C" int mrmp_state_dist(mrmp_ctx *c, int64_t n, const double *q1, const double *q2,
                               double *out) {
    return state_op(c, 0, n, q1, q2, out);
}

extern "C" int mrmp_mid_status(mrmp_ctx *c, int64_t n, const double *p, const double *q,
                               double *out) {
    return state_op(c, 1, n, p, q, out);
}

extern "C" int mrmp_nearest(mrmp_ctx *c, int n_v, const double *V, const double *q, int reversed,
                            int32_t *idx_out, double *dist_out) {
    NEED(c && n_v >= 1 && V && q && idx_out, "bad arguments");
    CU(cudaSetDevice(c->device));
    const size_t vb = sizeof(double) * c->d * (size_t)n_v, qb = sizeof(double) * c->d;
    const size_t o_q = align16(vb), o_out = o_q + align16(qb);
    int rc = ensure_small(c, o_out + 32);
    if (rc) return rc;
    memcpy(c->small_h, V, vb);
    memcpy(c->small_h + o_q, q, qb);
    CU(cudaMemcpyAsync(c->small_d, c->small_h, o_out, cudaMemcpyHostToDevice, c->s_k));
    CU(launch_nearest(c->cv, lcfg(c, c->s_k), n_v, (const double *)c->small_d,
                      (const double *)(c->small_d + o_q), reversed,
                      (int32_t *)(c->small_d + o_out + 16), (double *)(c->small_d + o_out)));
    CU(cudaMemcpyAsync(c->small_h + o_out, c->small_d + o_out, 32, cudaMemcpyDeviceToHost, c->s_k));
    CU(cudaStreamSynchronize(c->s_k));
    *idx_out = *(const int32_t *)(c->small_h + o_out + 16);
    if (dist_out) *dist_out = *(const double *)(c->small_h + o_out);
    return MRMP_OK;
}

/* extend! of RRT-Connect (sssp.jl:404-440) in one launch: see include/mrmp_b200.h */
extern "C" int mrmp_rrt_extend(mrmp_ctx *c, int agent, int n_v, const double *V,
                               const double *q_rand, int from_start, int steering_depth,
                               double epsilon, int32_t *flag_out, int32_t *near_out,
                               double *q_new_out) {
    NEED(c && n_v >= 1 && V && q_rand && flag_out && q_new_out, "bad arguments");
    NEED(agent >= 0 && agent < c->n_agents && steering_depth >= 0, "agent / depth out of range");
    CU(cudaSetDevice(c->device));
    const size_t vb = sizeof(double) * c->d * (size_t)n_v, qb = sizeof(double) * c->d;
    const size_t o_q = align16(vb), o_out = o_q + align16(qb);
    int rc = ensure_small(c, o_out + align16(qb) + 32);
    if (rc) return rc;
    memcpy(c->small_h, V, vb);
    memcpy(c->small_h + o_q, q_rand, qb);
    CU(cudaMemcpyAsync(c->small_d, c->small_h, o_out, cudaMemcpyHostToDevice, c->s_k));
    SsspExpand a{};
    a.agent = agent;
    a.depth = steering_depth;
    a.eps = epsilon;
    a.eps_r2 = sq_le_threshold(epsilon);
    CU(launch_rrt_extend(c->cv, lcfg(c, c->s_k), a, n_v, (const double *)c->small_d,
                         (const double *)(c->small_d + o_q), from_start ? 1 : 0,
                         (int32_t *)(c->small_d + o_out), (int32_t *)(c->small_d + o_out + 4),
                         (double *)(c->small_d + o_out + 16)));
    CU(cudaMemcpyAsync(c->small_h + o_out, c->small_d + o_out, 16 + qb, cudaMemcpyDeviceToHost,
                       c->s_k));
    CU(cudaStreamSynchronize(c->s_k));
    *flag_out = *(const int32_t *)(c->small_h + o_out);
    if (near_out) *near_out = *(const int32_t *)(c->small_h + o_out + 4);
    memcpy(q_new_out, c->small_h + o_out + 16, qb);
    return MRMP_OK;
}

// ---- distance tables -----------------------------------------------------------------------------
extern "C" int mrmp_distance_tables(mrmp_ctx *c, int n_graphs, const int32_t *v_off,
                                    const double *nodes, const int32_t *row_ptr,
                                    const int32_t *col_idx, const int32_t *goal,
                                    double *tables_out) {
    return mrmp_distance_tables_g(c, n_graphs, v_off, nodes, row_ptr, col_idx, goal, 0.0,
                                  tables_out);
}

extern "C" int mrmp_distance_tables_g(mrmp_ctx *c, int n_graphs, const int32_t *v_off,
                                      const double *nodes, const int32_t *row_ptr,
                                      const int32_t *col_idx, const int32_t *goal,
                                      double stay_penalty, double *tables_out) {
    NEED(c && n_graphs >= 0, "bad ctx / n_graphs");
    // the relaxation orders distances by their bit patterns (non-negative doubles) and a negative
    // cost on a self loop would decrease forever: gen_g_func's penalty is a non-negative number
    NEED(stay_penalty >= 0.0 && stay_penalty < INFINITY, "stay_penalty must be finite and >= 0");
    if (n_graphs == 0) return MRMP_OK;
    NEED(v_off && nodes && row_ptr && goal && tables_out, "NULL buffer");
    const int64_t V = v_off[n_graphs];
    NEED(v_off[0] == 0 && V >= 1, "v_off must start at 0");
    for (int g = 0; g < n_graphs; ++g) {
        const int nv = v_off[g + 1] - v_off[g];
        if (nv < 1 || (unsigned)goal[g] >= (unsigned)nv)
            return fail(MRMP_ERR_INVALID, "graph %d: empty or goal out of range", g);
    }
    const int64_t E = row_ptr[V];
    NEED(E >= 0 && (E == 0 || col_idx), "bad CSR");
    CU(cudaSetDevice(c->device));
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t o = off;
        off = align16(off + bytes);
        return o;
    };
    const size_t o_voff = take(4 * (size_t)(n_graphs + 1)), o_goal = take(4 * (size_t)n_graphs),
                 o_rp = take(4 * (size_t)(V + 1)), o_ci = take(4 * (size_t)E),
                 o_nodes = take(sizeof(double) * c->d * (size_t)V);
    const size_t in_end = off;
    const size_t o_tab = take(sizeof(double) * (size_t)V), o_w = take(sizeof(double) * (size_t)E);
    int rc = ensure_scratch(c, 5, off);
    if (rc) return rc;
    rc = ensure_small(c, in_end > sizeof(double) * (size_t)V ? in_end : sizeof(double) * (size_t)V);
    if (rc) return rc;
    unsigned char *h = c->small_h, *b = c->scratch[5];
    memcpy(h + o_voff, v_off, 4 * (size_t)(n_graphs + 1));
    memcpy(h + o_goal, goal, 4 * (size_t)n_graphs);
    memcpy(h + o_rp, row_ptr, 4 * (size_t)(V + 1));
    if (E) memcpy(h + o_ci, col_idx, 4 * (size_t)E);
    memcpy(h + o_nodes, nodes, sizeof(double) * c->d * (size_t)V);
    cudaStream_t st = c->s_k;
    CU(cudaMemcpyAsync(b, h, in_end, cudaMemcpyHostToDevice, st));
    CU(launch_distance_tables(c->cv, lcfg(c, st), n_graphs, (const int32_t *)(b + o_voff),
                              (const int32_t *)(b + o_rp), (const int32_t *)(b + o_ci),
                              (const double *)(b + o_nodes), (const int32_t *)(b + o_goal),
                              stay_penalty, (double *)(b + o_w), (double *)(b + o_tab)));
    CU(cudaMemcpyAsync(h, b + o_tab, sizeof(double) * (size_t)V, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    memcpy(tables_out, h, sizeof(double) * (size_t)V);
    return MRMP_OK;
}
