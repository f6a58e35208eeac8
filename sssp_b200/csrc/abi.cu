// abi.cu -- the C ABI (include/mrmp_b200.h): contexts, host<->device plumbing, roadmaps.
// Host-side C++ only orchestrates; all arithmetic of the hot path runs in the kernels.
#include <atomic>

#include <stdlib.h>

#include "abi_internal.cuh"
#include "arm_model.cuh"

// ------------------------------------------------------------------------------------------
// errors / bookkeeping
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

namespace mrmp {
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace mrmp

int mrmp_fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char *mrmp_last_error(void) { return g_err; }
extern "C" const char *mrmp_version(void) { return "mrmp_b200 0.1 (sm_100a, fp64)"; }
extern "C" int64_t mrmp_kernel_launch_count(void) { return g_launches.load(); }

extern "C" int mrmp_device_count(int *n_out) {
    if (!n_out) return fail(MRMP_ERR_INVALID, "n_out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *n_out = 0;
        return fail(MRMP_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *n_out = n;
    return MRMP_OK;
}

extern "C" int mrmp_host_alloc(void **ptr_out, int64_t bytes) {
    if (!ptr_out || bytes < 0) return fail(MRMP_ERR_INVALID, "bad arguments");
    CU(cudaHostAlloc(ptr_out, (size_t)(bytes > 0 ? bytes : 1), cudaHostAllocDefault));
    return MRMP_OK;
}

extern "C" int mrmp_host_free(void *ptr) {
    if (ptr) CU(cudaFreeHost(ptr));
    return MRMP_OK;
}

static int ensure_slots(mrmp_ctx *c, size_t bytes) {
    if (bytes <= c->slot_bytes) return MRMP_OK;
    for (int s = 0; s < kSlots; ++s) {
        if (c->slot_buf[s]) CU(cudaFree(c->slot_buf[s]));
        c->slot_buf[s] = nullptr;
    }
    c->slot_bytes = 0;
    for (int s = 0; s < kSlots; ++s) CU(cudaMalloc((void **)&c->slot_buf[s], bytes));
    c->slot_bytes = bytes;
    return MRMP_OK;
}

extern "C" int mrmp_ctx_create(mrmp_ctx **out, int model, int n_agents, const double *rads,
                               const double *base_pos, const double *aux, int n_obs,
                               const double *obstacles_aos, double step_dist,
                               double safety_dist, double max_dist, int device) {
    if (!out) return fail(MRMP_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (model < 0 || model > MRMP_MODEL_CAPSULE3D)
        return fail(MRMP_ERR_INVALID, "unknown model %d", model);
    if (n_agents <= 0 || !rads) return fail(MRMP_ERR_INVALID, "need n_agents > 0 and rads");
    if (n_obs < 0 || (n_obs > 0 && !obstacles_aos))
        return fail(MRMP_ERR_INVALID, "need obstacles_aos for n_obs > 0");
    if ((model == MRMP_MODEL_ARM22 || model == MRMP_MODEL_ARM33) && !base_pos)
        return fail(MRMP_ERR_INVALID, "arm models need base_pos (positions)");
    if (model == MRMP_MODEL_CAPSULE3D && !aux)
        return fail(MRMP_ERR_INVALID, "capsule3d needs aux (axises)");
    if (model == MRMP_MODEL_CAPSULE3D && n_agents > 1024)
        return fail(MRMP_ERR_UNSUPPORTED, "capsule3d supports at most 1024 agents per ctx");
    if ((int64_t)n_agents * (n_obs > 0 ? n_obs : 1) > (int64_t)1 << 24)
        return fail(MRMP_ERR_UNSUPPORTED, "n_agents * n_obs too large");
    int ndev = 0;
    cudaError_t e0 = cudaGetDeviceCount(&ndev);
    if (e0 != cudaSuccess || ndev == 0)
        return fail(MRMP_ERR_CUDA, "no CUDA device (%s); mrmp_b200 has no CPU fallback",
                    e0 != cudaSuccess ? cudaGetErrorString(e0) : "0 devices");
    if (device < 0 || device >= ndev)
        return fail(MRMP_ERR_INVALID, "device %d out of range (have %d)", device, ndev);
    CU(cudaSetDevice(device));

    mrmp_ctx *c = new (std::nothrow) mrmp_ctx();
    if (!c) return fail(MRMP_ERR_NOMEM, "out of host memory");
    c->device = device;
    c->model = model;
    static const int kStateDim[] = {2, 3, 2, 3, 6, 6, 6}, kWorkDim[] = {2, 3, 2, 2, 2, 3, 3};
    c->d = kStateDim[model];
    const int wd = kWorkDim[model];  // workspace dimension of obstacle centres / joint points
    c->n_agents = n_agents;
    c->n_obs = n_obs;
    c->h_rads.assign(rads, rads + n_agents);
    for (int i = 0; i < n_agents; ++i)
        if (rads[i] > c->max_rad) c->max_rad = rads[i];
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);

    const int N = n_agents, M = n_obs;
    const int mp = pad2(M > 0 ? M : 1), np = pad2(N);
    const bool arm = model == MRMP_MODEL_ARM22 || model == MRMP_MODEL_ARM33;  // fixed bases
    // obstacle threshold o.r + rads[i] (points, capsule3d.jl:76) or o.r alone (arm22.jl:46,
    // line2d.jl:39, snake2d.jl:78, arm33.jl:60); pair threshold rads[i] + rads[j] (point.jl:11,
    // capsule3d.jl:163) or safety_dist
    const bool rad_thr = model == MRMP_MODEL_POINT2D || model == MRMP_MODEL_POINT3D ||
                         model == MRMP_MODEL_CAPSULE3D;
    const int n_thr_rows = rad_thr ? N : 1;
    const bool pair_tab = rad_thr && N <= 1024;

    std::vector<double> h;
    auto section = [&](size_t len) {
        size_t off = h.size();
        h.resize(off + ((len + 1) & ~(size_t)1), 0.0);
        return (int)off;
    };
    CtxView &cv = c->cv;
    cv.model = model;
    cv.d = c->d;
    cv.wd = wd;
    cv.n_agents = N;
    cv.n_obs = M;
    cv.mp = mp;
    cv.np = np;
    // ---- connect segment
    cv.con_off = 0;
    cv.obs_off = section((size_t)(wd + 1) * mp);
    for (int m = 0; m < M; ++m)
        for (int k = 0; k <= wd; ++k) h[cv.obs_off + k * mp + m] = obstacles_aos[(wd + 1) * m + k];
    for (int m = M; m < mp; ++m) {  // padding obstacles: far away, zero radius
        for (int k = 0; k < wd; ++k) h[cv.obs_off + k * mp + m] = 1.0e300;
        h[cv.obs_off + wd * mp + m] = 0.0;
    }
    cv.rads_off = section(np);
    for (int i = 0; i < N; ++i) h[cv.rads_off + i] = rads[i];
    cv.base_off = section(arm ? wd * np : 0);
    if (arm)
        for (int i = 0; i < wd * N; ++i) h[cv.base_off + i] = base_pos[i];
    cv.aux_off = section(aux ? np : 0);
    if (aux)
        for (int i = 0; i < N; ++i) h[cv.aux_off + i] = aux[i];
    cv.t2_obs_off = section((size_t)n_thr_rows * mp);
    cv.g_obs_off = section((size_t)n_thr_rows * mp);
    for (int i = 0; i < n_thr_rows; ++i)
        for (int m = 0; m < mp; ++m) {
            const double orad = h[cv.obs_off + wd * mp + m];
            const double thr = rad_thr ? orad + rads[i] : orad;  // o.r + rads[i]  point2d.jl:54,63
            h[cv.t2_obs_off + i * mp + m] = m < M ? sq_lt_threshold(thr) : 0.0;
            const double tp = (thr > 0 ? thr : 0.0) + kFilterPad;
            h[cv.g_obs_off + i * mp + m] = 2.0 * tp * tp * (1.0 + 1e-9);
        }
    cv.con_len = (int)h.size() - cv.con_off;
    // ---- collide segment
    cv.col_off = (int)h.size();
    cv.crads_off = section(np);
    for (int i = 0; i < N; ++i) h[cv.crads_off + i] = rads[i];
    cv.cbase_off = section(arm ? wd * np : 0);
    if (arm)
        for (int i = 0; i < wd * N; ++i) h[cv.cbase_off + i] = base_pos[i];
    cv.caux_off = section(aux ? np : 0);
    if (aux)
        for (int i = 0; i < N; ++i) h[cv.caux_off + i] = aux[i];
    cv.has_pair_tab = pair_tab ? 1 : 0;
    cv.t2_pair_off = section(pair_tab ? (size_t)N * N : 0);
    cv.g_pair_off = section(pair_tab ? (size_t)N * N : 0);
    if (pair_tab)
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) {
                const double thr = rads[i] + rads[j];  // point.jl:11,15,30
                h[cv.t2_pair_off + i * N + j] = sq_lt_threshold(thr);
                const double tp = (thr > 0 ? thr : 0.0) + kFilterPad;
                h[cv.g_pair_off + i * N + j] = 3.0 * tp * tp * (1.0 + 1e-9);
            }
    cv.col_len = (int)h.size() - cv.col_off;
    cv.step_dist = step_dist;
    cv.safety_dist = safety_dist;
    cv.max_dist = max_dist;
    cv.safety_t2 = sq_lt_threshold(safety_dist);
    {
        const StepRat sr = step_rat(step_dist);  // D-independent half of Julia's 0:step:D lift
        cv.step_rat_ok = sr.ok;
        cv.step_rat_num = sr.num;
        cv.step_rat_den = sr.den;
    }

    cudaError_t e = cudaMalloc((void **)&c->blob, h.size() * sizeof(double));
    if (e == cudaSuccess)
        e = cudaMemcpy(c->blob, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->stats_d, 8 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(c->stats_d, 0, 8 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_k, cudaStreamNonBlocking);
    c->s_k_owned = c->s_k;
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_aux, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_aux, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_order, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_read, cudaEventDisableTiming);
    {
        const char *ov = getenv("MRMP_TABLE_OVERLAP");
        c->table_overlap = !(ov && ov[0] == '0');
    }
    for (int s = 0; s < kSlots && e == cudaSuccess; ++s) {
        e = cudaEventCreateWithFlags(&c->ev_in[s], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_k[s], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_out[s], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        mrmp_ctx_destroy(c);
        return fail(MRMP_ERR_CUDA, "ctx setup: %s", cudaGetErrorString(e));
    }
    cv.blob = c->blob;
    *out = c;
    return MRMP_OK;
}

extern "C" void mrmp_ctx_destroy(mrmp_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->s_k) cudaStreamSynchronize(c->s_k);
    for (int s = 0; s < kSlots; ++s) {
        if (c->ev_in[s]) cudaEventDestroy(c->ev_in[s]);
        if (c->ev_k[s]) cudaEventDestroy(c->ev_k[s]);
        if (c->ev_out[s]) cudaEventDestroy(c->ev_out[s]);
        if (c->slot_buf[s]) cudaFree(c->slot_buf[s]);
    }
    if (c->small_h) cudaFreeHost(c->small_h);
    if (c->small_d) cudaFree(c->small_d);
    for (int k = 0; k < mrmp_ctx::kScratch; ++k)
        if (c->scratch[k]) cudaFree(c->scratch[k]);
    if (c->s_k_owned) cudaStreamDestroy(c->s_k_owned);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    if (c->s_aux) cudaStreamDestroy(c->s_aux);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_aux) cudaEventDestroy(c->ev_aux);
    if (c->ev_order) cudaEventDestroy(c->ev_order);
    if (c->ev_read) cudaEventDestroy(c->ev_read);
    if (c->blob) cudaFree(c->blob);
    if (c->stats_d) cudaFree(c->stats_d);
    delete c;
}

extern "C" int mrmp_ctx_state_dim(const mrmp_ctx *c) { return c ? c->d : 0; }

extern "C" int mrmp_ctx_stats(mrmp_ctx *c, int64_t *out4, int reset) {
    if (!c || !out4) return fail(MRMP_ERR_INVALID, "bad arguments");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->s_k));
    unsigned long long h[4];
    CU(cudaMemcpy(h, c->stats_d, sizeof(h), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 4; ++k) out4[k] = (int64_t)h[k];
    if (reset) CU(cudaMemset(c->stats_d, 0, sizeof(h)));
    return MRMP_OK;
}

/* 8 counters: the four of mrmp_ctx_stats, then [4] arm22 link-pair midpoint bounds evaluated */
extern "C" int mrmp_ctx_stats8(mrmp_ctx *c, int64_t *out8, int reset) {
    if (!c || !out8) return fail(MRMP_ERR_INVALID, "bad arguments");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->s_k));
    unsigned long long h[8];
    CU(cudaMemcpy(h, c->stats_d, sizeof(h), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 8; ++k) out8[k] = (int64_t)h[k];
    if (reset) CU(cudaMemset(c->stats_d, 0, sizeof(h)));
    return MRMP_OK;
}

extern "C" int mrmp_ctx_set_stream(mrmp_ctx *c, void *stream) {
    if (!c) return fail(MRMP_ERR_INVALID, "ctx is NULL");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->s_k));
    c->s_k = stream ? (cudaStream_t)stream : c->s_k_owned;
    return MRMP_OK;
}

extern "C" int mrmp_ctx_sync(mrmp_ctx *c) {
    if (!c) return fail(MRMP_ERR_INVALID, "ctx is NULL");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->s_in));
    CU(cudaStreamSynchronize(c->s_k));
    CU(cudaStreamSynchronize(c->s_out));
    return MRMP_OK;
}

// ------------------------------------------------------------------------------------------
// host-buffer batches: small single-stream path, chunked 3-stream pipeline for large n
// ------------------------------------------------------------------------------------------
struct Arr {
    const void *in;  // host source (inputs) or NULL
    void *out;       // host destination (outputs) or NULL
    size_t stride;   // bytes per check
};

// launch(dev_ptrs[], n_chunk, stream) -> cudaError_t
template <class F>
static int run_batch(mrmp_ctx *c, int64_t n, std::vector<Arr> &arrs, F launch) {
    CU(cudaSetDevice(c->device));
    const size_t na = arrs.size();
    std::vector<void *> dptr(na);
    if (n <= kSmallBatch) {
        // pack everything into one pinned buffer: one H2D, one kernel, one D2H
        std::vector<size_t> off(na);
        size_t in_bytes = 0, tot = 0;
        for (size_t a = 0; a < na; ++a)
            if (arrs[a].in) {
                off[a] = tot;
                tot += align16(arrs[a].stride * (size_t)n);
            }
        in_bytes = tot;
        for (size_t a = 0; a < na; ++a)
            if (arrs[a].out) {
                off[a] = tot;
                tot += align16(arrs[a].stride * (size_t)n);
            }
        int rc = ensure_small(c, tot);
        if (rc) return rc;
        if (n <= kTinyBatch) {
            // closure-sized calls (one check .. a few hundred): latency is everything, so the
            // kernel works on the mapped pinned block itself: one launch + one sync
            unsigned char *hb = nullptr;
            CU(cudaHostGetDevicePointer((void **)&hb, c->small_h, 0));
            for (size_t a = 0; a < na; ++a) {
                dptr[a] = hb + off[a];
                if (arrs[a].in) memcpy(c->small_h + off[a], arrs[a].in, arrs[a].stride * (size_t)n);
            }
            CU(launch(dptr.data(), n, c->s_k));
            CU(cudaStreamSynchronize(c->s_k));
            for (size_t a = 0; a < na; ++a)
                if (arrs[a].out)
                    memcpy(arrs[a].out, c->small_h + off[a], arrs[a].stride * (size_t)n);
            return MRMP_OK;
        }
        for (size_t a = 0; a < na; ++a) {
            dptr[a] = c->small_d + off[a];
            if (arrs[a].in) memcpy(c->small_h + off[a], arrs[a].in, arrs[a].stride * (size_t)n);
        }
        if (in_bytes)
            CU(cudaMemcpyAsync(c->small_d, c->small_h, in_bytes, cudaMemcpyHostToDevice, c->s_k));
        CU(launch(dptr.data(), n, c->s_k));
        if (tot > in_bytes)
            CU(cudaMemcpyAsync(c->small_h + in_bytes, c->small_d + in_bytes, tot - in_bytes,
                               cudaMemcpyDeviceToHost, c->s_k));
        CU(cudaStreamSynchronize(c->s_k));
        for (size_t a = 0; a < na; ++a)
            if (arrs[a].out) memcpy(arrs[a].out, c->small_h + off[a], arrs[a].stride * (size_t)n);
        return MRMP_OK;
    }
    // chunked pipeline
    std::vector<size_t> off(na);
    size_t tot = 0;
    for (size_t a = 0; a < na; ++a) {
        off[a] = tot;
        tot += align16(arrs[a].stride * (size_t)kChunk);
    }
    int rc = ensure_slots(c, tot);
    if (rc) return rc;
    int64_t done = 0;
    for (int64_t ch = 0; done < n; ++ch) {
        const int s = (int)(ch % kSlots);
        const int64_t len = n - done < kChunk ? n - done : kChunk;
        if (ch >= kSlots) {
            CU(cudaStreamWaitEvent(c->s_in, c->ev_k[s], 0));   // inputs of this slot consumed
            CU(cudaStreamWaitEvent(c->s_k, c->ev_out[s], 0));  // outputs of this slot drained
        }
        for (size_t a = 0; a < na; ++a) {
            dptr[a] = c->slot_buf[s] + off[a];
            if (arrs[a].in)
                CU(cudaMemcpyAsync(dptr[a], (const char *)arrs[a].in + arrs[a].stride * (size_t)done,
                                   arrs[a].stride * (size_t)len, cudaMemcpyHostToDevice, c->s_in));
        }
        CU(cudaEventRecord(c->ev_in[s], c->s_in));
        CU(cudaStreamWaitEvent(c->s_k, c->ev_in[s], 0));
        CU(launch(dptr.data(), len, c->s_k));
        CU(cudaEventRecord(c->ev_k[s], c->s_k));
        CU(cudaStreamWaitEvent(c->s_out, c->ev_k[s], 0));
        for (size_t a = 0; a < na; ++a)
            if (arrs[a].out)
                CU(cudaMemcpyAsync((char *)arrs[a].out + arrs[a].stride * (size_t)done, dptr[a],
                                   arrs[a].stride * (size_t)len, cudaMemcpyDeviceToHost, c->s_out));
        CU(cudaEventRecord(c->ev_out[s], c->s_out));
        done += len;
    }
    CU(cudaStreamSynchronize(c->s_out));
    CU(cudaStreamSynchronize(c->s_k));
    return MRMP_OK;
}

static int check_agents_host(const mrmp_ctx *c, int64_t n, const int32_t *a) {
    for (int64_t k = 0; k < n; ++k)
        if ((unsigned)a[k] >= (unsigned)c->n_agents)
            return fail(MRMP_ERR_INVALID, "agent index %d out of range at %lld", a[k], (long long)k);
    return MRMP_OK;
}


// ---- connect ---------------------------------------------------------------------------------
extern "C" int mrmp_connect_points(mrmp_ctx *c, int64_t n, const int32_t *agent, const double *q,
                                   uint8_t *out) {
    NEED(c && n >= 0, "bad ctx / n");
    if (n == 0) return MRMP_OK;
    NEED(agent && q && out, "NULL buffer");
    int rc = check_agents_host(c, n, agent);
    if (rc) return rc;
    const size_t sb = sizeof(double) * c->d;
    std::vector<Arr> arrs = {{agent, nullptr, 4}, {q, nullptr, sb}, {nullptr, out, 1}};
    return run_batch(c, n, arrs, [&](void **d, int64_t len, cudaStream_t s) {
        return launch_connect_points(c->cv, lcfg(c, s), len, (const int32_t *)d[0],
                                     (const double *)d[1], (uint8_t *)d[2]);
    });
}

extern "C" int mrmp_connect_point(mrmp_ctx *c, int agent, const double *q, uint8_t *out) {
    int32_t a = agent;
    return mrmp_connect_points(c, 1, &a, q, out);
}

extern "C" int mrmp_connect_edges(mrmp_ctx *c, int64_t n, const int32_t *agent,
                                  const double *q_from, const double *q_to, uint8_t *out) {
    NEED(c && n >= 0, "bad ctx / n");
    if (n == 0) return MRMP_OK;
    NEED(agent && q_from && q_to && out, "NULL buffer");
    int rc = check_agents_host(c, n, agent);
    if (rc) return rc;
    const size_t sb = sizeof(double) * c->d;
    std::vector<Arr> arrs = {{agent, nullptr, 4}, {q_from, nullptr, sb}, {q_to, nullptr, sb},
                             {nullptr, out, 1}};
    return run_batch(c, n, arrs, [&](void **d, int64_t len, cudaStream_t s) {
        return launch_connect_edges(c->cv, lcfg(c, s), len, (const int32_t *)d[0],
                                    (const double *)d[1], (const double *)d[2], (uint8_t *)d[3]);
    });
}

extern "C" int mrmp_connect_edge(mrmp_ctx *c, int agent, const double *q_from, const double *q_to,
                                 uint8_t *out) {
    int32_t a = agent;
    return mrmp_connect_edges(c, 1, &a, q_from, q_to, out);
}

// ---- collide ---------------------------------------------------------------------------------
extern "C" int mrmp_collide_pairs(mrmp_ctx *c, int64_t n, const int32_t *i, const int32_t *j,
                                  const double *qi_from, const double *qi_to,
                                  const double *qj_from, const double *qj_to, uint8_t *out) {
    NEED(c && n >= 0, "bad ctx / n");
    if (n == 0) return MRMP_OK;
    NEED(i && j && qi_from && qi_to && qj_from && qj_to && out, "NULL buffer");
    int rc = check_agents_host(c, n, i);
    if (!rc) rc = check_agents_host(c, n, j);
    if (rc) return rc;
    const size_t sb = sizeof(double) * c->d;
    std::vector<Arr> arrs = {{i, nullptr, 4},        {j, nullptr, 4},        {qi_from, nullptr, sb},
                             {qi_to, nullptr, sb},   {qj_from, nullptr, sb}, {qj_to, nullptr, sb},
                             {nullptr, out, 1}};
    return run_batch(c, n, arrs, [&](void **d, int64_t len, cudaStream_t s) {
        return launch_collide_pairs(c->cv, lcfg(c, s), len, (const int32_t *)d[0],
                                    (const int32_t *)d[1], (const double *)d[2],
                                    (const double *)d[3], (const double *)d[4],
                                    (const double *)d[5], (uint8_t *)d[6]);
    });
}

extern "C" int mrmp_collide_pair(mrmp_ctx *c, int i, int j, const double *qi_from,
                                 const double *qi_to, const double *qj_from, const double *qj_to,
                                 uint8_t *out) {
    int32_t a = i, b = j;
    return mrmp_collide_pairs(c, 1, &a, &b, qi_from, qi_to, qj_from, qj_to, out);
}

extern "C" int mrmp_collide_static_pairs(mrmp_ctx *c, int64_t n, const int32_t *i, const int32_t *j,
                                         const double *qi, const double *qj, uint8_t *out) {
    NEED(c && n >= 0, "bad ctx / n");
    if (n == 0) return MRMP_OK;
    NEED(i && j && qi && qj && out, "NULL buffer");
    int rc = check_agents_host(c, n, i);
    if (!rc) rc = check_agents_host(c, n, j);
    if (rc) return rc;
    const size_t sb = sizeof(double) * c->d;
    std::vector<Arr> arrs = {{i, nullptr, 4}, {j, nullptr, 4}, {qi, nullptr, sb}, {qj, nullptr, sb},
                             {nullptr, out, 1}};
    return run_batch(c, n, arrs, [&](void **d, int64_t len, cudaStream_t s) {
        return launch_collide_static_pairs(c->cv, lcfg(c, s), len, (const int32_t *)d[0],
                                           (const int32_t *)d[1], (const double *)d[2],
                                           (const double *)d[3], (uint8_t *)d[4]);
    });
}

extern "C" int mrmp_collide_configs(mrmp_ctx *c, int64_t n_cfg, const double *Q_from,
                                    const double *Q_to, uint8_t *out, int64_t *first_pair_out) {
    NEED(c && n_cfg >= 0, "bad ctx / n_cfg");
    if (n_cfg == 0) return MRMP_OK;
    NEED(Q_from && Q_to && out, "NULL buffer");
    const size_t sb = sizeof(double) * c->d * c->n_agents;
    std::vector<int64_t> fp_tmp;
    int64_t *fp = first_pair_out;
    if (!fp) {
        fp_tmp.resize((size_t)n_cfg);
        fp = fp_tmp.data();
    }
    std::vector<Arr> arrs = {{Q_from, nullptr, sb}, {Q_to, nullptr, sb}, {nullptr, out, 1},
                             {nullptr, fp, 8}};
    return run_batch(c, n_cfg, arrs, [&](void **d, int64_t len, cudaStream_t s) {
        return launch_collide_configs(c->cv, lcfg(c, s), len, (const double *)d[0],
                                      (const double *)d[1], (uint8_t *)d[2], (int64_t *)d[3]);
    });
}

extern "C" int mrmp_collide_config_moves(mrmp_ctx *c, int agent_i, const double *Q, int64_t n_cand,
                                         const double *q_to, uint8_t *out) {
    NEED(c && n_cand >= 0, "bad ctx / n_cand");
    NEED((unsigned)agent_i < (unsigned)c->n_agents, "agent_i out of range");
    if (n_cand == 0) return MRMP_OK;
    NEED(Q && q_to && out, "NULL buffer");
    CU(cudaSetDevice(c->device));
    // Q (N states) rides in front of the candidate list of the first chunk; it is tiny, so it
    // is uploaded once into the small staging buffer and shared by every chunk.
    const size_t qb = align16(sizeof(double) * c->d * c->n_agents);
    int rc = ensure_small(c, qb + 64);
    if (rc) return rc;
    if (n_cand <= kSmallBatch) {
        const size_t sb = sizeof(double) * c->d;
        const size_t in_b = align16(sb * (size_t)n_cand), out_b = align16((size_t)n_cand);
        rc = ensure_small(c, qb + in_b + out_b);
        if (rc) return rc;
        memcpy(c->small_h, Q, sizeof(double) * c->d * c->n_agents);
        memcpy(c->small_h + qb, q_to, sb * (size_t)n_cand);
        CU(cudaMemcpyAsync(c->small_d, c->small_h, qb + in_b, cudaMemcpyHostToDevice, c->s_k));
        CU(launch_collide_config_moves(c->cv, lcfg(c, c->s_k), agent_i, (const double *)c->small_d,
                                       n_cand, (const double *)(c->small_d + qb),
                                       c->small_d + qb + in_b));
        CU(cudaMemcpyAsync(c->small_h + qb + in_b, c->small_d + qb + in_b, (size_t)n_cand,
                           cudaMemcpyDeviceToHost, c->s_k));
        CU(cudaStreamSynchronize(c->s_k));
        memcpy(out, c->small_h + qb + in_b, (size_t)n_cand);
        return MRMP_OK;
    }
    memcpy(c->small_h, Q, sizeof(double) * c->d * c->n_agents);
    CU(cudaMemcpyAsync(c->small_d, c->small_h, qb, cudaMemcpyHostToDevice, c->s_k));
    CU(cudaStreamSynchronize(c->s_k));
    const double *Qd = (const double *)c->small_d;
    const size_t sb = sizeof(double) * c->d;
    std::vector<Arr> arrs = {{q_to, nullptr, sb}, {nullptr, out, 1}};
    return run_batch(c, n_cand, arrs, [&](void **d, int64_t len, cudaStream_t s) {
        return launch_collide_config_moves(c->cv, lcfg(c, s), agent_i, Qd, len,
                                           (const double *)d[0], (uint8_t *)d[1]);
    });
}

// ---- validate (src/utils.jl:345-363): all transitions of a solution in two launches ----------
extern "C" int mrmp_validate_solution(mrmp_ctx *c, int64_t T, const double *solution,
                                      int32_t *ok_out, int64_t *t_out, int32_t *kind_out,
                                      int32_t *i_out, int32_t *j_out) {
    NEED(c && T >= 1 && solution && ok_out, "bad arguments");
    const int N = c->n_agents, d = c->d;
    *ok_out = 1;
    if (t_out) *t_out = -1;
    if (kind_out) *kind_out = 0;
    if (i_out) *i_out = -1;
    if (j_out) *j_out = -1;
    if (T == 1) return MRMP_OK;
    CU(cudaSetDevice(c->device));
    const int64_t nt = T - 1, n_con = nt * N;
    const size_t sol_b = sizeof(double) * d * (size_t)N * T;
    const size_t o_ag = align16(sol_b), o_con = align16(o_ag + 4 * (size_t)n_con),
                 o_col = align16(o_con + (size_t)n_con), o_fp = align16(o_col + (size_t)nt),
                 tot = o_fp + 8 * (size_t)nt;
    int rc = ensure_scratch(c, 1, tot);
    if (rc) return rc;
    std::vector<unsigned char> hbuf(tot);
    int32_t *ag = (int32_t *)(hbuf.data() + o_ag);
    for (int64_t k = 0; k < n_con; ++k) ag[k] = (int32_t)(k % N);
    unsigned char *b = c->scratch[1];
    cudaStream_t s = c->s_k;
    CU(cudaMemcpyAsync(b, solution, sol_b, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(b + o_ag, ag, 4 * (size_t)n_con, cudaMemcpyHostToDevice, s));
    const double *sol_d = (const double *)b;
    const double *next_d = sol_d + (size_t)N * d;  // configuration t+1 follows configuration t
    // connect(solution[t-1][i].q, solution[t][i].q, i)  for every t, i   (:347-358)
    CU(launch_connect_edges(c->cv, lcfg(c, s), n_con, (const int32_t *)(b + o_ag), sol_d, next_d,
                            b + o_con));
    // collide(solution[t-1], solution[t])  for every t, with the first colliding pair (:360)
    CU(launch_collide_configs(c->cv, lcfg(c, s), nt, sol_d, next_d, b + o_col,
                              (int64_t *)(b + o_fp)));
    CU(cudaMemcpyAsync(hbuf.data() + o_con, b + o_con, tot - o_con, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    const uint8_t *con = hbuf.data() + o_con, *col = hbuf.data() + o_col;
    const int64_t *fp = (const int64_t *)(hbuf.data() + o_fp);
    for (int64_t t = 0; t < nt; ++t) {  // the reference's order: per step, connectivity first
        for (int i = 0; i < N; ++i)
            if (!con[t * N + i]) {
                *ok_out = 0;
                if (t_out) *t_out = t + 1;
                if (kind_out) *kind_out = 1;
                if (i_out) *i_out = i;
                return MRMP_OK;
            }
        if (col[t]) {
            *ok_out = 0;
            if (t_out) *t_out = t + 1;
            if (kind_out) *kind_out = 2;
            if (i_out) *i_out = (int32_t)(fp[t] / N);
            if (j_out) *j_out = (int32_t)(fp[t] % N);
            return MRMP_OK;
        }
    }
    return MRMP_OK;
}

// ---- first conflict of a set of paths (CBS get_constraints, src/solvers/cbs.jl:374-412) ------
extern "C" int mrmp_first_conflict(mrmp_ctx *c, int64_t T, const double *configs, int64_t *t_out,
                                   int32_t *kind_out, int32_t *i_out, int32_t *j_out) {
    NEED(c && T >= 1 && configs && t_out && kind_out && i_out && j_out, "bad arguments");
    const int N = c->n_agents, d = c->d;
    *t_out = -1;
    *kind_out = 0;
    *i_out = *j_out = -1;
    if (T == 1 || N < 2) return MRMP_OK;
    CU(cudaSetDevice(c->device));
    const int64_t nt = T - 1;
    const size_t sol_b = sizeof(double) * d * (size_t)N * T;
    const size_t o_v = align16(sol_b), o_e = align16(o_v + (size_t)nt),
                 o_vfp = align16(o_e + (size_t)nt), o_efp = o_vfp + 8 * (size_t)nt,
                 tot = o_efp + 8 * (size_t)nt;
    int rc = ensure_scratch(c, 1, tot);
    if (rc) return rc;
    unsigned char *b = c->scratch[1];
    cudaStream_t s = c->s_k;
    CU(cudaMemcpyAsync(b, configs, sol_b, cudaMemcpyHostToDevice, s));
    const double *from = (const double *)b, *to = from + (size_t)N * d;
    // vertex conflicts collide(q_i_to, q_j_to, i, j) (:391): stationary 6-arity checks at the
    // target configuration -- for every model the reference's 6-arity arithmetic on two
    // zero-length moves reduces to its 4-arity test (point.jl:9-16, arm22.jl:100-162)
    CU(launch_collide_configs(c->cv, lcfg(c, s), nt, to, to, b + o_v, (int64_t *)(b + o_vfp)));
    // edge conflicts collide(v_i_from.q, v_i_to.q, v_j_from.q, v_j_to.q, i, j) (:401)
    CU(launch_collide_configs(c->cv, lcfg(c, s), nt, from, to, b + o_e, (int64_t *)(b + o_efp)));
    std::vector<int64_t> fp(2 * (size_t)nt);
    CU(cudaMemcpyAsync(fp.data(), b + o_vfp, 16 * (size_t)nt, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    for (int64_t t = 0; t < nt; ++t) {
        const int64_t v = fp[t], e = fp[nt + t];
        if (v < 0 && e < 0) continue;
        // per pair the vertex test comes first, so a tie on the pair goes to the vertex conflict
        const bool vertex = v >= 0 && (e < 0 || v <= e);
        const int64_t pr = vertex ? v : e;
        *t_out = t + 1;
        *kind_out = vertex ? 1 : 2;
        *i_out = (int32_t)(pr / N);
        *j_out = (int32_t)(pr % N);
        return MRMP_OK;
    }
    return MRMP_OK;
}

// ---- device-pointer forms --------------------------------------------------------------------
#define DEV_PROLOGUE(c, n)                          \
    NEED(c && n >= 0, "bad ctx / n");               \
    if (n == 0) return MRMP_OK;                     \
    CU(cudaSetDevice(c->device))

static int need_aligned(const mrmp_ctx *c, const void *p) {
    if (c->d == 2 && !aligned16(p))
        return fail(MRMP_ERR_INVALID, "device state arrays must be 16-byte aligned");
    return MRMP_OK;
}

extern "C" int mrmp_connect_points_dev(mrmp_ctx *c, int64_t n, const int32_t *agent,
                                       const double *q, uint8_t *out, void *stream) {
    DEV_PROLOGUE(c, n);
    if (need_aligned(c, q)) return MRMP_ERR_INVALID;
    CU(launch_connect_points(c->cv, lcfg(c, (cudaStream_t)stream), n, agent, q, out));
    return MRMP_OK;
}

extern "C" int mrmp_connect_edges_dev(mrmp_ctx *c, int64_t n, const int32_t *agent,
                                      const double *q_from, const double *q_to, uint8_t *out,
                                      void *stream) {
    DEV_PROLOGUE(c, n);
    if (need_aligned(c, q_from) || need_aligned(c, q_to)) return MRMP_ERR_INVALID;
    CU(launch_connect_edges(c->cv, lcfg(c, (cudaStream_t)stream), n, agent, q_from, q_to, out));
    return MRMP_OK;
}

extern "C" int mrmp_collide_pairs_dev(mrmp_ctx *c, int64_t n, const int32_t *i, const int32_t *j,
                                      const double *qi_from, const double *qi_to,
                                      const double *qj_from, const double *qj_to, uint8_t *out,
                                      void *stream) {
    DEV_PROLOGUE(c, n);
    if (need_aligned(c, qi_from) || need_aligned(c, qi_to) || need_aligned(c, qj_from) ||
        need_aligned(c, qj_to))
        return MRMP_ERR_INVALID;
    CU(launch_collide_pairs(c->cv, lcfg(c, (cudaStream_t)stream), n, i, j, qi_from, qi_to,
                            qj_from, qj_to, out));
    return MRMP_OK;
}

extern "C" int mrmp_collide_static_pairs_dev(mrmp_ctx *c, int64_t n, const int32_t *i,
                                             const int32_t *j, const double *qi, const double *qj,
                                             uint8_t *out, void *stream) {
    DEV_PROLOGUE(c, n);
    if (need_aligned(c, qi) || need_aligned(c, qj)) return MRMP_ERR_INVALID;
    CU(launch_collide_static_pairs(c->cv, lcfg(c, (cudaStream_t)stream), n, i, j, qi, qj, out));
    return MRMP_OK;
}

extern "C" int mrmp_collide_configs_dev(mrmp_ctx *c, int64_t n_cfg, const double *Q_from,
                                        const double *Q_to, uint8_t *out, int64_t *first_pair_out,
                                        void *stream) {
    DEV_PROLOGUE(c, n_cfg);
    if (need_aligned(c, Q_from) || need_aligned(c, Q_to)) return MRMP_ERR_INVALID;
    if (c->model >= MRMP_MODEL_ARM22 && !first_pair_out) {
        // the arm kernel keeps the running first hit per configuration in this array
        int rc = ensure_scratch(c, 3, sizeof(int64_t) * (size_t)n_cfg);
        if (rc) return rc;
        first_pair_out = (int64_t *)c->scratch[3];
    }
    CU(launch_collide_configs(c->cv, lcfg(c, (cudaStream_t)stream), n_cfg, Q_from, Q_to, out,
                              first_pair_out));
    return MRMP_OK;
}

extern "C" int mrmp_collide_config_moves_dev(mrmp_ctx *c, int agent_i, const double *Q,
                                             int64_t n_cand, const double *q_to, uint8_t *out,
                                             void *stream) {
    DEV_PROLOGUE(c, n_cand);
    NEED((unsigned)agent_i < (unsigned)c->n_agents, "agent_i out of range");
    if (need_aligned(c, Q) || need_aligned(c, q_to)) return MRMP_ERR_INVALID;
    CU(launch_collide_config_moves(c->cv, lcfg(c, (cudaStream_t)stream), agent_i, Q, n_cand, q_to,
                                   out));
    return MRMP_OK;
}

// ------------------------------------------------------------------------------------------
// sampler
// ------------------------------------------------------------------------------------------
extern "C" int mrmp_sample_states(mrmp_ctx *c, int agent, uint64_t seed, uint64_t t0, int64_t n,
                                  double *q_out) {
    NEED(c && n >= 0, "bad ctx / n");
    NEED((unsigned)agent < (unsigned)c->n_agents, "agent out of range");
    if (n == 0) return MRMP_OK;
    NEED(q_out, "NULL buffer");
    CU(cudaSetDevice(c->device));
    double *d = nullptr;
    const size_t bytes = sizeof(double) * c->d * (size_t)n;
    CU(cudaMalloc((void **)&d, bytes));
    cudaError_t e = launch_sample_states(c->cv, lcfg(c, c->s_k), agent, seed, t0, n, d);
    if (e == cudaSuccess) e = cudaMemcpyAsync(q_out, d, bytes, cudaMemcpyDeviceToHost, c->s_k);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->s_k);
    cudaFree(d);
    CU(e);
    return MRMP_OK;
}

static const int64_t kMaxTriesFactor = 1 << 14;  // give up after 16384 * n_wanted tries

extern "C" int mrmp_sample_free(mrmp_ctx *c, int agent, uint64_t seed, int64_t n_wanted,
                                double *q_out, int64_t *n_tried_out) {
    NEED(c && n_wanted >= 0 && n_wanted < (1 << 30), "bad ctx / n_wanted");
    NEED((unsigned)agent < (unsigned)c->n_agents, "agent out of range");
    if (n_wanted == 0) {
        if (n_tried_out) *n_tried_out = 0;
        return MRMP_OK;
    }
    NEED(q_out, "NULL buffer");
    CU(cudaSetDevice(c->device));
    double *d = nullptr;
    const size_t bytes = sizeof(double) * c->d * (size_t)n_wanted;
    CU(cudaMalloc((void **)&d, align16(bytes) + 16));
    int64_t *dt = (int64_t *)((char *)d + align16(bytes));
    int64_t tried = -1;
    cudaError_t e = launch_sample_free(c->cv, lcfg(c, c->s_k), agent, 1, 0, (int)n_wanted,
                                       (int)n_wanted, seed, d, dt, kMaxTriesFactor * n_wanted);
    if (e == cudaSuccess) e = cudaMemcpyAsync(q_out, d, bytes, cudaMemcpyDeviceToHost, c->s_k);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&tried, dt, 8, cudaMemcpyDeviceToHost, c->s_k);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->s_k);
    cudaFree(d);
    CU(e);
    if (n_tried_out) *n_tried_out = tried;
    if (tried < 0) return fail(MRMP_ERR_CAPACITY, "free-space sampling gave up (space too cluttered?)");
    return MRMP_OK;
}

// ------------------------------------------------------------------------------------------
// roadmaps
// ------------------------------------------------------------------------------------------
struct mrmp_roadmaps {
    mrmp_ctx *ctx = nullptr;
    int agent0 = 0, n_agents = 0, nv_cap = 0, nv = 0;
    int words = 0;
    double *nodes = nullptr;     // [n_agents][nv][d] (stride nv, repacked on set)
    double *nodes_owned = nullptr;  // == nodes unless bound to caller storage
    uint32_t *bitmap = nullptr;  // [n_agents*nv][words]
    int32_t *row_cnt = nullptr, *row_ptr = nullptr, *col_idx = nullptr;
    int64_t col_cap = 0;
    int64_t *nnz_d = nullptr;    // device scalar
    long long *chunk_ws = nullptr;  // row-scan workspace: one partial sum per 1024 rows
    int64_t chunk_ws_len = 0;
    int64_t *tries_d = nullptr;  // [n_agents]
    double *rad_d = nullptr;     // [2][n_agents]: rad, r2
    int64_t nnz = -1;            // host copy after build
    bool built = false;
    // asynchronous build (mrmp_roadmaps_build_async): launched, nnz still on the device (nnz_d and
    // its mapped mirror); mrmp_roadmaps_finish() waits and validates the capacities used meanwhile
    bool pending = false;
    int64_t pend_slot_nnz = -1;  // cap_nnz of a shard pushed while pending (-1: none)
    int64_t pend_rows_cap = -1;  // row capacity of a table computed while pending (-1: none)
    int64_t pend_read_cap = -1;  // col_cap of a read-back queued while pending (-1: none)
    cudaStream_t pend_stream = nullptr;  // stream of an asynchronous adoption (finish waits on it)
    // CSR edges expanded to explicit (agent, from, to) rows for the cross kernels
    int32_t *e_agent = nullptr;
    double *e_from = nullptr, *e_to = nullptr;
    int64_t e_cap = 0;
    bool e_valid = false;
    void *e_rows = nullptr;      // the same edges as Morton-sorted row records (point models)
    // uniform-grid spatial hash path (large nv): cell lists and the row staging buffer
    int32_t *g_cell_start = nullptr, *g_cursor = nullptr, *g_ids = nullptr;
    int64_t g_cells_cap = 0;
    int64_t *g_row_off = nullptr;
    int32_t *g_tmp = nullptr;
    int64_t g_tmp_cap = 0;
    unsigned long long *g_tmp_cursor = nullptr;
    bool used_grid = false;      // diagnostics: which neighbour search the last build took
    // pinned staging owned by the object: [2*d*n] init/goal, [2*n] rad/r2, [n] tries, [2] nnz
    unsigned char *stage_h = nullptr;
    unsigned char *stage_d = nullptr;  // device alias of stage_h (mapped: kernels read the tiny
                                       // inputs and write the verdict scalars in place)
    bool tries_pending = false;  // sampling verdicts copied to stage_h but not yet examined
    double *st_ends() { return (double *)stage_h; }
    double *st_rad() { return st_ends() + (size_t)2 * ctx->d * n_agents; }
    int64_t *st_tries() { return (int64_t *)(st_rad() + (size_t)2 * n_agents); }
    int64_t *st_nnz() { return st_tries() + n_agents; }
    template <class T>
    T *dev(T *host_ptr) { return (T *)(stage_d + ((unsigned char *)host_ptr - stage_h)); }
};

// Examine the sampling verdicts of a mrmp_roadmaps_sample call that deferred its synchronise.
static int roadmaps_settle(mrmp_roadmaps *rm, bool synced) {
    if (!rm->tries_pending) return MRMP_OK;
    if (!synced) CU(cudaStreamSynchronize(rm->ctx->s_k));
    rm->tries_pending = false;
    const int64_t *tries = rm->st_tries();
    for (int a = 0; a < rm->n_agents; ++a)
        if (tries[a] < 0) {
            rm->nv = 0;
            rm->built = false;
            return fail(MRMP_ERR_CAPACITY, "free-space sampling gave up for agent %d",
                        rm->agent0 + a);
        }
    return MRMP_OK;
}

extern "C" int mrmp_roadmaps_create(mrmp_ctx *c, int agent0, int n_agents, int nv_cap,
                                    mrmp_roadmaps **out) {
    NEED(c && out, "bad ctx / out");
    *out = nullptr;
    NEED(agent0 >= 0 && n_agents > 0 && agent0 + n_agents <= c->n_agents, "agent range");
    NEED(nv_cap >= 2 && nv_cap <= (1 << 20), "nv_cap out of range");
    CU(cudaSetDevice(c->device));
    mrmp_roadmaps *rm = new (std::nothrow) mrmp_roadmaps();
    if (!rm) return fail(MRMP_ERR_NOMEM, "out of host memory");
    rm->ctx = c;
    rm->agent0 = agent0;
    rm->n_agents = n_agents;
    rm->nv_cap = nv_cap;
    const int words_cap = (nv_cap + 31) / 32;
    const int64_t rows = (int64_t)n_agents * nv_cap;
    cudaError_t e = cudaMalloc((void **)&rm->nodes_owned, sizeof(double) * c->d * rows);
    rm->nodes = rm->nodes_owned;
    (void)words_cap;  // the adjacency bitmap of the all-pairs path is allocated on first use
    if (e == cudaSuccess) e = cudaMalloc((void **)&rm->row_cnt, sizeof(int32_t) * (rows + 1));
    if (e == cudaSuccess) e = cudaMalloc((void **)&rm->row_ptr, sizeof(int32_t) * (rows + 1));
    if (e == cudaSuccess) e = cudaMalloc((void **)&rm->nnz_d, 16);
    rm->chunk_ws_len = (rows + 1023) / 1024 + 1;
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&rm->chunk_ws, sizeof(long long) * rm->chunk_ws_len);
    if (e == cudaSuccess) e = cudaMalloc((void **)&rm->tries_d, sizeof(int64_t) * n_agents);
    if (e == cudaSuccess) e = cudaMalloc((void **)&rm->rad_d, sizeof(double) * 2 * n_agents);
    if (e == cudaSuccess)
        e = cudaHostAlloc((void **)&rm->stage_h,
                          sizeof(double) * ((size_t)2 * c->d + 3) * n_agents + 64,
                          cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void **)&rm->stage_d, rm->stage_h, 0);
    if (e != cudaSuccess) {
        mrmp_roadmaps_destroy(rm);
        return fail(MRMP_ERR_CUDA, "roadmaps alloc: %s", cudaGetErrorString(e));
    }
    *out = rm;
    return MRMP_OK;
}

extern "C" void mrmp_roadmaps_destroy(mrmp_roadmaps *rm) {
    if (!rm) return;
    cudaSetDevice(rm->ctx->device);
    cudaStreamSynchronize(rm->ctx->s_k);
    cudaFree(rm->nodes_owned);
    cudaFree(rm->bitmap);
    cudaFree(rm->row_cnt);
    cudaFree(rm->row_ptr);
    cudaFree(rm->col_idx);
    cudaFree(rm->nnz_d);
    cudaFree(rm->chunk_ws);
    cudaFree(rm->tries_d);
    cudaFree(rm->rad_d);
    cudaFree(rm->e_agent);
    cudaFree(rm->e_from);
    cudaFree(rm->e_to);
    cudaFree(rm->e_rows);
    cudaFree(rm->g_cell_start);
    cudaFree(rm->g_cursor);
    cudaFree(rm->g_ids);
    cudaFree(rm->g_row_off);
    cudaFree(rm->g_tmp);
    cudaFree(rm->g_tmp_cursor);
    cudaFreeHost(rm->stage_h);
    delete rm;
}

extern "C" int mrmp_roadmaps_set_nodes(mrmp_roadmaps *rm, int nv, const double *nodes) {
    NEED(rm && nodes, "bad roadmaps / nodes");
    NEED(nv >= 1 && nv <= rm->nv_cap, "nv out of range");
    mrmp_ctx *c = rm->ctx;
    CU(cudaSetDevice(c->device));
    rm->nv = nv;
    rm->tries_pending = false;
    rm->built = false;
    CU(cudaMemcpyAsync(rm->nodes, nodes, sizeof(double) * c->d * (size_t)nv * rm->n_agents,
                       cudaMemcpyHostToDevice, c->s_k));
    CU(cudaStreamSynchronize(c->s_k));
    return MRMP_OK;
}

extern "C" int mrmp_roadmaps_set_nodes_dev(mrmp_roadmaps *rm, int nv, const double *nodes_dev,
                                           void *stream) {
    NEED(rm && nodes_dev, "bad roadmaps / nodes");
    NEED(nv >= 1 && nv <= rm->nv_cap, "nv out of range");
    mrmp_ctx *c = rm->ctx;
    CU(cudaSetDevice(c->device));
    rm->nv = nv;
    rm->tries_pending = false;
    rm->built = false;
    CU(cudaMemcpyAsync(rm->nodes, nodes_dev, sizeof(double) * c->d * (size_t)nv * rm->n_agents,
                       cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return MRMP_OK;
}

extern "C" int mrmp_roadmaps_bind_nodes(mrmp_roadmaps *rm, int nv, double *nodes_dev) {
    NEED(rm, "bad roadmaps");
    if (!nodes_dev) {  // back to the set's own storage
        rm->nodes = rm->nodes_owned;
        rm->tries_pending = false;
    rm->built = false;
        return MRMP_OK;
    }
    NEED(nv >= 1 && nv <= rm->nv_cap, "nv out of range");
    NEED(aligned16(nodes_dev), "nodes_dev must be 16-byte aligned");
    rm->nodes = nodes_dev;
    rm->nv = nv;
    rm->built = false;
    return MRMP_OK;
}

extern "C" int mrmp_roadmaps_sample(mrmp_roadmaps *rm, int nv, uint64_t seed, const double *q_init,
                                    const double *q_goal, int64_t *tries_out) {
    NEED(rm && q_init && q_goal, "bad roadmaps / init / goal");
    NEED(nv >= 2 && nv <= rm->nv_cap, "nv out of range");
    mrmp_ctx *c = rm->ctx;
    CU(cudaSetDevice(c->device));
    rm->nv = nv;
    rm->built = false;
    const int d = c->d, n = rm->n_agents;
    // vertices 0 / 1 = init / goal (prm.jl:232)
    const size_t sb = sizeof(double) * d;
    if (rm->tries_pending) CU(cudaStreamSynchronize(c->s_k));  // stage_h is about to be rewritten
    rm->tries_pending = false;
    unsigned char *ends = rm->stage_h;
    for (int a = 0; a < n; ++a) {
        memcpy(ends + (2 * a) * sb, q_init + (size_t)a * d, sb);
        memcpy(ends + (2 * a + 1) * sb, q_goal + (size_t)a * d, sb);
    }
    // the kernel reads init / goal from, and writes the per-agent verdicts to, mapped memory
    CU(launch_sample_free(c->cv, lcfg(c, c->s_k), rm->agent0, n, 2, nv, nv, seed, rm->nodes,
                          rm->dev(rm->st_tries()), kMaxTriesFactor * (int64_t)nv,
                          rm->dev(rm->st_ends())));
    rm->tries_pending = true;
    // Without a tries_out the verdicts are examined at the next call that synchronises anyway
    // (build / get_nodes / device_ptrs), which saves one host round trip per roadmap set.
    if (!tries_out) return MRMP_OK;
    CU(cudaStreamSynchronize(c->s_k));
    for (int a = 0; a < n; ++a) tries_out[a] = rm->st_tries()[a];
    return roadmaps_settle(rm, true);
}

static int roadmaps_emit(mrmp_roadmaps *rm, cudaStream_t s) {
    mrmp_ctx *c = rm->ctx;
    const int64_t rows = (int64_t)rm->n_agents * rm->nv;
    if (rm->col_cap == 0) {
        rm->col_cap = rows * 32;
        CU(cudaMalloc((void **)&rm->col_idx, sizeof(int32_t) * rm->col_cap));
    }
    CU(launch_prm_emit(lcfg(c, s), rows, rm->nv, rm->bitmap, rm->words, rm->row_ptr, rm->col_idx,
                       rm->col_cap));
    return MRMP_OK;
}

static const int kGridMinNv = 2048;   // below this the all-pairs kernel is faster
static const int kGridMinCells = 4;   // cells per dimension needed for the grid to prune

// neighbour pass through the uniform-grid spatial hash; leaves row_cnt / g_row_off / g_tmp
static int roadmaps_build_grid(mrmp_roadmaps *rm, int g, cudaStream_t s) {
    mrmp_ctx *c = rm->ctx;
    const int n = rm->n_agents, nv = rm->nv;
    const int64_t rows = (int64_t)n * nv;
    int64_t cells = 1;
    for (int k = 0; k < c->d; ++k) cells *= g;
    if ((int64_t)n * (cells + 1) > rm->g_cells_cap) {
        CU(cudaStreamSynchronize(s));
        cudaFree(rm->g_cell_start);
        cudaFree(rm->g_cursor);
        rm->g_cell_start = rm->g_cursor = nullptr;
        rm->g_cells_cap = (int64_t)n * (cells + 1);
        CU(cudaMalloc((void **)&rm->g_cell_start, sizeof(int32_t) * rm->g_cells_cap));
        CU(cudaMalloc((void **)&rm->g_cursor, sizeof(int32_t) * rm->g_cells_cap));
    }
    if (!rm->g_ids) {
        const int64_t rows_cap = (int64_t)n * rm->nv_cap;
        CU(cudaMalloc((void **)&rm->g_ids, sizeof(int32_t) * rows_cap));
        CU(cudaMalloc((void **)&rm->g_row_off, sizeof(int64_t) * rows_cap));
        CU(cudaMalloc((void **)&rm->g_tmp_cursor, 16));
    }
    LaunchCfg lc = lcfg(c, s);
    CU(launch_grid_bin(c->cv, lc, n, nv, g, (int)cells, rm->nodes, rm->g_cell_start, rm->g_cursor,
                       rm->g_ids));
    if (rm->g_tmp_cap == 0) {
        // expected degree: nv * volume of the rad-ball, with head room
        double rmax = 1.0 / g;
        double ball = c->d == 2 ? 3.1416 * rmax * rmax : 4.19 * rmax * rmax * rmax;
        double deg = (double)nv * (ball < 1.0 ? ball : 1.0) * 1.25 + 16.0;
        rm->g_tmp_cap = (int64_t)(deg * (double)rows);
        if (rm->g_tmp_cap > rows * (int64_t)(nv - 1)) rm->g_tmp_cap = rows * (int64_t)(nv - 1);
        CU(cudaMalloc((void **)&rm->g_tmp, sizeof(int32_t) * rm->g_tmp_cap));
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
        CU(launch_prm_grid_adjacency(c->cv, lc, rm->agent0, n, nv, g, (int)cells, rm->g_cell_start,
                                     rm->g_ids, rm->rad_d + n, rm->rad_d, rm->nodes, rm->row_cnt,
                                     rm->g_row_off, rm->g_tmp, rm->g_tmp_cap, rm->g_tmp_cursor));
        unsigned long long need = 0;
        CU(cudaMemcpyAsync(&need, rm->g_tmp_cursor, 8, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        if ((int64_t)need <= rm->g_tmp_cap) return MRMP_OK;
        // the staging buffer was too small: now the exact size is known
        CU(cudaFree(rm->g_tmp));
        rm->g_tmp = nullptr;
        rm->g_tmp_cap = (int64_t)need + 64;
        CU(cudaMalloc((void **)&rm->g_tmp, sizeof(int32_t) * rm->g_tmp_cap));
    }
    return fail(MRMP_ERR_CUDA, "grid neighbour pass did not converge");
}

static int roadmaps_build_impl(mrmp_roadmaps *rm, const double *rad, bool async) {
    NEED(rm && rad, "bad roadmaps / rad");
    NEED(rm->nv >= 1, "no nodes set");
    rm->pending = false;
    rm->pend_slot_nnz = rm->pend_rows_cap = rm->pend_read_cap = -1;
    rm->pend_stream = nullptr;
    mrmp_ctx *c = rm->ctx;
    CU(cudaSetDevice(c->device));
    const int n = rm->n_agents, nv = rm->nv;
    const int64_t rows = (int64_t)n * nv;
    rm->words = (nv + 31) / 32;
    int rc = MRMP_OK;
    double *hr = rm->st_rad();
    double rad_max = 0.0;
    bool finite = true;
    for (int a = 0; a < n; ++a) {
        hr[a] = rad[a];
        hr[n + a] = sq_le_threshold(rad[a]);
        finite = finite && rad[a] > 0.0 && rad[a] < 1e300;
        if (rad[a] > rad_max) rad_max = rad[a];
    }
    // neighbour search: uniform-grid spatial hash for large point roadmaps, all pairs otherwise
    int g = 0;
    if (finite && c->model <= MRMP_MODEL_POINT3D && nv >= kGridMinNv) {
        const double gf = floor(1.0 / rad_max);
        const double gmax = c->d == 2 ? 1024.0 : 100.0;   // <= ~10^6 cells per agent
        g = (int)(gf < gmax ? gf : gmax);
    }
    rm->used_grid = g >= kGridMinCells;
    int64_t *h_nnz = rm->st_nnz();
    if (rm->used_grid) {
        CU(cudaMemcpyAsync(rm->rad_d, hr, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, c->s_k));
        rc = roadmaps_build_grid(rm, g, c->s_k);
        if (rc) return rc;
        CU(launch_row_scan(lcfg(c, c->s_k), rows, rm->row_cnt, rm->row_ptr, rm->nnz_d,
                           rm->chunk_ws, rm->chunk_ws_len));
        CU(cudaMemcpyAsync(h_nnz, rm->nnz_d, 8, cudaMemcpyDeviceToHost, c->s_k));
        CU(cudaStreamSynchronize(c->s_k));
        rc = roadmaps_settle(rm, true);
        if (rc) return rc;
        rm->nnz = *h_nnz;
        if (rm->nnz >= ((int64_t)1 << 31))
            return fail(MRMP_ERR_UNSUPPORTED, "nnz exceeds int32 CSR");
        if (rm->nnz > rm->col_cap) {
            if (rm->col_idx) CU(cudaFree(rm->col_idx));
            rm->col_idx = nullptr;
            rm->col_cap = rm->nnz + rm->nnz / 4 + 64;
            CU(cudaMalloc((void **)&rm->col_idx, sizeof(int32_t) * rm->col_cap));
        }
        CU(launch_grid_gather_rows(lcfg(c, c->s_k), rows, rm->row_ptr, rm->g_row_off, rm->g_tmp,
                                   rm->col_idx, rm->col_cap));
        CU(cudaStreamSynchronize(c->s_k));
        rm->built = true;
        rm->e_valid = false;
        return MRMP_OK;
    }
    if (!rm->bitmap) {
        const int64_t rows_cap = (int64_t)n * rm->nv_cap;
        const int words_cap = (rm->nv_cap + 31) / 32;
        CU(cudaMalloc((void **)&rm->bitmap, sizeof(uint32_t) * rows_cap * words_cap));
    }
    // rad / r2 are read from, and nnz is written to, the mapped staging block
    CU(launch_prm_adjacency(c->cv, lcfg(c, c->s_k), rm->agent0, n, nv, rm->dev(hr) + n,
                            rm->dev(hr), rm->nodes, rm->bitmap, rm->words, rm->row_cnt, hr, hr + n));
    CU(launch_row_scan(lcfg(c, c->s_k), rows, rm->row_cnt, rm->row_ptr, rm->nnz_d,
                       rm->chunk_ws, rm->chunk_ws_len, rm->dev(h_nnz)));
    rc = roadmaps_emit(rm, c->s_k);
    if (rc) return rc;
    if (async) {  // no wait: nnz is examined by mrmp_roadmaps_finish()
        rm->pending = true;
        rm->built = true;
        rm->nnz = -1;
        rm->e_valid = false;
        return MRMP_OK;
    }
    CU(cudaStreamSynchronize(c->s_k));
    rc = roadmaps_settle(rm, true);
    if (rc) return rc;
    rm->nnz = *h_nnz;
    if (rm->nnz >= ((int64_t)1 << 31)) return fail(MRMP_ERR_UNSUPPORTED, "nnz exceeds int32 CSR");
    if (rm->nnz > rm->col_cap) {  // rare: grow and re-emit
        CU(cudaFree(rm->col_idx));
        rm->col_idx = nullptr;
        rm->col_cap = rm->nnz + rm->nnz / 4;
        CU(cudaMalloc((void **)&rm->col_idx, sizeof(int32_t) * rm->col_cap));
        rc = roadmaps_emit(rm, c->s_k);
        if (rc) return rc;
        CU(cudaStreamSynchronize(c->s_k));
    }
    rm->built = true;
    rm->e_valid = false;
    return MRMP_OK;
}

extern "C" int mrmp_roadmaps_build(mrmp_roadmaps *rm, const double *rad) {
    return roadmaps_build_impl(rm, rad, false);
}

/* The same launches without the wait for nnz (the uniform-grid path of large roadmaps needs nnz on
 * the host to size its gather and builds synchronously). */
extern "C" int mrmp_roadmaps_build_async(mrmp_roadmaps *rm, const double *rad) {
    return roadmaps_build_impl(rm, rad, true);
}

/* Wait for an asynchronous build (or adoption) and for everything launched on it since, then
 * validate: the CSR fitted its buffer, a shard pushed meanwhile fitted its slot, a table computed
 * meanwhile had room for every row.  MRMP_ERR_CAPACITY means those consumers worked on a truncated
 * edge list: the buffers have been grown, build again. */
extern "C" int mrmp_roadmaps_finish(mrmp_roadmaps *rm, int64_t *nnz_out) {
    NEED(rm, "bad roadmaps");
    mrmp_ctx *c = rm->ctx;
    if (!rm->pending) {
        NEED(rm->built, "roadmaps not built");
        if (nnz_out) *nnz_out = rm->nnz;
        return MRMP_OK;
    }
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->s_k));
    if (rm->pend_stream && rm->pend_stream != c->s_k) CU(cudaStreamSynchronize(rm->pend_stream));
    rm->pend_stream = nullptr;
    rm->pending = false;
    int rc = roadmaps_settle(rm, true);
    if (rc) {
        rm->built = false;
        return rc;
    }
    rm->nnz = *rm->st_nnz();
    if (nnz_out) *nnz_out = rm->nnz;
    const int64_t slot = rm->pend_slot_nnz, rows_cap = rm->pend_rows_cap, read_cap = rm->pend_read_cap;
    rm->pend_slot_nnz = rm->pend_rows_cap = rm->pend_read_cap = -1;
    if (rm->nnz >= ((int64_t)1 << 31)) return fail(MRMP_ERR_UNSUPPORTED, "nnz exceeds int32 CSR");
    if (rm->nnz > rm->col_cap) {
        CU(cudaFree(rm->col_idx));
        rm->col_idx = nullptr;
        rm->col_cap = rm->nnz + rm->nnz / 4;
        CU(cudaMalloc((void **)&rm->col_idx, sizeof(int32_t) * rm->col_cap));
        rm->built = false;
        return fail(MRMP_ERR_CAPACITY, "asynchronous build: %lld edges overflowed the CSR buffer "
                    "(grown now); build again", (long long)rm->nnz);
    }
    if (slot >= 0 && rm->nnz > slot)
        return fail(MRMP_ERR_CAPACITY, "asynchronous build: shard of %lld edges pushed into a slot "
                    "of %lld", (long long)rm->nnz, (long long)slot);
    if (rows_cap >= 0 && rm->nnz > rows_cap)
        return fail(MRMP_ERR_CAPACITY, "asynchronous build: table of %lld rows computed into room "
                    "for %lld", (long long)rm->nnz, (long long)rows_cap);
    if (read_cap >= 0 && rm->nnz > read_cap)
        return fail(MRMP_ERR_CAPACITY, "asynchronous build: %lld edges read back into col_cap %lld",
                    (long long)rm->nnz, (long long)read_cap);
    return MRMP_OK;
}

/* 1 when the last mrmp_roadmaps_build used the uniform-grid spatial hash, 0 for all pairs */
extern "C" int mrmp_roadmaps_used_grid(mrmp_roadmaps *rm) { return rm && rm->used_grid ? 1 : 0; }

extern "C" int mrmp_roadmaps_set_csr_dev(mrmp_roadmaps *rm, const int32_t *row_ptr_dev,
                                         const int32_t *col_idx_dev, int64_t nnz, void *stream) {
    NEED(rm && row_ptr_dev && nnz >= 0 && nnz < ((int64_t)1 << 31), "bad arguments");
    NEED(rm->nv >= 1, "no nodes set");
    mrmp_ctx *c = rm->ctx;
    CU(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t rows = (int64_t)rm->n_agents * rm->nv;
    if (nnz > rm->col_cap) {
        CU(cudaStreamSynchronize(s));
        if (rm->col_idx) CU(cudaFree(rm->col_idx));
        rm->col_idx = nullptr;
        rm->col_cap = nnz + nnz / 4 + 64;
        CU(cudaMalloc((void **)&rm->col_idx, sizeof(int32_t) * rm->col_cap));
    }
    CU(cudaMemcpyAsync(rm->row_ptr, row_ptr_dev, sizeof(int32_t) * (rows + 1),
                       cudaMemcpyDeviceToDevice, s));
    if (nnz > 0) {
        NEED(col_idx_dev, "col_idx_dev is NULL");
        CU(cudaMemcpyAsync(rm->col_idx, col_idx_dev, sizeof(int32_t) * nnz,
                           cudaMemcpyDeviceToDevice, s));
    }
    rm->nnz = nnz;
    rm->built = true;
    rm->e_valid = false;
    return MRMP_OK;
}

// ---- CSR shards of a multi-GPU build ------------------------------------------------------------
extern "C" int mrmp_roadmaps_pack_csr_shard_dev(mrmp_roadmaps *rm, int cap_agents, int64_t cap_nnz,
                                                int32_t *slot_dev, void *stream) {
    NEED(rm && slot_dev && cap_agents >= rm->n_agents && cap_nnz >= 0, "bad arguments");
    NEED(rm->built, "roadmaps not built");
    NEED(!rm->pending, "asynchronous build pending: call mrmp_roadmaps_finish first");
    if (rm->nnz > cap_nnz)
        return fail(MRMP_ERR_CAPACITY, "shard holds %lld edges, slot only %lld", (long long)rm->nnz,
                    (long long)cap_nnz);
    mrmp_ctx *c = rm->ctx;
    CU(cudaSetDevice(c->device));
    CU(order_after_ctx(c, (cudaStream_t)stream));
    CU(launch_csr_pack_shard(lcfg(c, (cudaStream_t)stream), (int64_t)rm->n_agents * rm->nv,
                             (int64_t)cap_agents * rm->nv, rm->row_ptr, rm->col_idx, slot_dev));
    return MRMP_OK;
}

static int set_csr_shards_impl(mrmp_roadmaps *rm, int world, int cap_agents, int64_t cap_nnz,
                               const int32_t *slots_dev, void *stream, int64_t *nnz_out, bool async);

// room for `need` column indices (contents are not kept: the caller rewrites the CSR)
static int grow_col_idx(mrmp_roadmaps *rm, int64_t need, cudaStream_t s) {
    if (need <= rm->col_cap) return MRMP_OK;
    CU(cudaStreamSynchronize(s));
    if (s != rm->ctx->s_k) CU(cudaStreamSynchronize(rm->ctx->s_k));
    if (rm->col_idx) CU(cudaFree(rm->col_idx));
    rm->col_idx = nullptr;
    rm->col_cap = need + 64;
    CU(cudaMalloc((void **)&rm->col_idx, sizeof(int32_t) * rm->col_cap));
    return MRMP_OK;
}

extern "C" int mrmp_roadmaps_set_csr_shards_dev(mrmp_roadmaps *rm, int world, int cap_agents,
                                                int64_t cap_nnz, const int32_t *slots_dev,
                                                void *stream, int64_t *nnz_out) {
    return set_csr_shards_impl(rm, world, cap_agents, cap_nnz, slots_dev, stream, nnz_out, false);
}

static int set_csr_shards_impl(mrmp_roadmaps *rm, int world, int cap_agents, int64_t cap_nnz,
                               const int32_t *slots_dev, void *stream, int64_t *nnz_out, bool async) {
    NEED(rm && slots_dev && world >= 1 && cap_agents >= 1 && cap_nnz >= 0, "bad arguments");
    NEED(rm->nv >= 1, "no nodes set");
    NEED((int64_t)world * cap_agents >= rm->n_agents, "shards do not cover the agents of this set");
    mrmp_ctx *c = rm->ctx;
    CU(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t rows = (int64_t)rm->n_agents * rm->nv, rows_cap = (int64_t)cap_agents * rm->nv;
    const int64_t stride = rows_cap + cap_nnz;
    LaunchCfg lc = lcfg(c, s);
    CU(order_after_ctx(c, s));  // earlier consumers of this set's CSR on the ctx's stream
    CU(launch_csr_shard_counts(lc, rows, rows_cap, stride, slots_dev, rm->row_cnt));
    rm->pending = false;
    rm->pend_slot_nnz = rm->pend_rows_cap = rm->pend_read_cap = -1;
    CU(launch_row_scan(lc, rows, rm->row_cnt, rm->row_ptr, rm->nnz_d,
                           rm->chunk_ws, rm->chunk_ws_len, async ? rm->dev(rm->st_nnz()) : nullptr));
    int rc0 = grow_col_idx(rm, (int64_t)world * cap_nnz, s);  // upper bound of the stitched size
    if (rc0) return rc0;
    CU(launch_csr_stitch(lc, rows, rows_cap, stride, slots_dev, rm->row_ptr, rm->col_idx,
                         rm->col_cap));
    if (async) {  // nnz stays on the device (and in the mapped mirror) until mrmp_roadmaps_finish
        rm->pend_stream = s;
        rm->pending = true;
        rm->built = true;
        rm->nnz = -1;
        rm->e_valid = false;
        if (nnz_out) *nnz_out = -1;
        return MRMP_OK;
    }
    int rc = ensure_small(c, 64);
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->small_h, rm->nnz_d, 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    rm->nnz = *(int64_t *)c->small_h;
    rm->built = true;
    rm->e_valid = false;
    if (nnz_out) *nnz_out = rm->nnz;
    return roadmaps_settle(rm, false);  // a deferred sampling verdict of this set, if any
}

// ---- the same exchange without a collective: shards pushed into the peers' memory ------------
extern "C" int mrmp_roadmaps_push_csr_shard(mrmp_roadmaps *rm, mrmp_exchange *x, int ch,
                                            int world, int cap_agents, int64_t cap_nnz,
                                            void *stream) {
    NEED(rm && x && world >= 1 && world <= 32 && cap_agents >= rm->n_agents && cap_nnz >= 0,
         "bad arguments");
    NEED(rm->built, "roadmaps not built");
    if (rm->pending)
        rm->pend_slot_nnz = cap_nnz;  // validated by mrmp_roadmaps_finish (the kernel truncates)
    else if (rm->nnz > cap_nnz)
        return fail(MRMP_ERR_CAPACITY, "shard holds %lld edges, slot only %lld", (long long)rm->nnz,
                    (long long)cap_nnz);
    const int64_t rows_cap = (int64_t)cap_agents * rm->nv;
    NEED(mrmp_exchange_slot_bytes(x, ch) >= (int64_t)sizeof(int32_t) * (rows_cap + cap_nnz),
         "exchange slot too small for the shard");
    mrmp_ctx *c = rm->ctx;
    CU(cudaSetDevice(c->device));
    CU(order_after_ctx(c, (cudaStream_t)stream));
    LaunchCfg lc = lcfg(c, (cudaStream_t)stream);
    for (int d0 = 0; d0 < world; d0 += 8) {
        CsrPushDst dst;
        dst.n = world - d0 < 8 ? world - d0 : 8;
        for (int k = 0; k < 8; ++k) {
            dst.slot[k] = k < dst.n ? (int32_t *)mrmp_exchange_ptr(x, ch, d0 + k) : nullptr;
            NEED(k >= dst.n || dst.slot[k], "exchange is not connected to every rank");
        }
        CU(launch_csr_pack_push(lc, (int64_t)rm->n_agents * rm->nv, rows_cap, cap_nnz, rm->row_ptr,
                                rm->col_idx, dst));
    }
    return mrmp_exchange_signal(x, ch, world >= 32 ? 0xffffffffu : ((1u << world) - 1u), stream);
}

static int adopt_impl(mrmp_roadmaps *rm, mrmp_exchange *x, int ch, int world, int cap_agents,
                      int64_t cap_nnz, void *stream, int64_t *nnz_out, bool async);

extern "C" int mrmp_roadmaps_adopt_csr_shards(mrmp_roadmaps *rm, mrmp_exchange *x, int ch,
                                              int world, int cap_agents, int64_t cap_nnz,
                                              void *stream, int64_t *nnz_out) {
    return adopt_impl(rm, x, ch, world, cap_agents, cap_nnz, stream, nnz_out, false);
}

/* no wait: the stitched set stays pending until mrmp_roadmaps_finish (which also reports a peer
 * that never published, through mrmp_exchange_error) */
extern "C" int mrmp_roadmaps_adopt_csr_shards_async(mrmp_roadmaps *rm, mrmp_exchange *x, int ch,
                                                    int world, int cap_agents, int64_t cap_nnz,
                                                    void *stream) {
    return adopt_impl(rm, x, ch, world, cap_agents, cap_nnz, stream, nullptr, true);
}

static int adopt_impl(mrmp_roadmaps *rm, mrmp_exchange *x, int ch, int world, int cap_agents,
                      int64_t cap_nnz, void *stream, int64_t *nnz_out, bool async) {
    NEED(rm && x && world >= 1 && world <= 32 && cap_nnz >= 0, "bad arguments");
    // (a reallocation synchronises the stream: before the wait is queued, not behind it)
    CU(cudaSetDevice(rm->ctx->device));
    int rc = grow_col_idx(rm, (int64_t)world * cap_nnz, (cudaStream_t)stream);
    if (rc) return rc;
    rc = mrmp_exchange_wait(x, ch, world >= 32 ? 0xffffffffu : ((1u << world) - 1u), stream);
    if (rc) return rc;
    NEED(((int64_t)cap_agents * rm->nv + cap_nnz) % 4 == 0 &&
             mrmp_exchange_slot_bytes(x, ch) ==
                 (int64_t)sizeof(int32_t) * ((int64_t)cap_agents * rm->nv + cap_nnz),
         "exchange slot must be exactly 4 * (cap_agents * nv + cap_nnz) bytes, a multiple of 16");
    rc = set_csr_shards_impl(rm, world, cap_agents, cap_nnz,
                             (const int32_t *)mrmp_exchange_slots(x, ch), stream, nnz_out, async);
    if (rc || async) return rc;
    const int bad = mrmp_exchange_error(x);  // the stream has been synchronised by the stitch
    if (bad) return fail(MRMP_ERR_CUDA, "exchange: rank %d never published its shard", bad - 1);
    return MRMP_OK;
}

/* table rows of this set (out_bits of mrmp_roadmaps_edge_conflicts_dev / _async, words_per_row
 * words each, in device memory of this rank) -> this rank's slot of channel ch in rank dst's
 * buffer; the number of rows is the set's edge count, taken on the device while a build is
 * pending.  max_rows = room of the source buffer (and of the slot) in rows. */
extern "C" int mrmp_roadmaps_push_table_rows(mrmp_roadmaps *rm, mrmp_exchange *x, int ch, int dst,
                                             const uint32_t *bits_dev, int words_per_row,
                                             int64_t max_rows, void *stream) {
    NEED(rm && x && bits_dev && words_per_row >= 1 && max_rows >= 0, "bad arguments");
    NEED(rm->built, "roadmaps not built");
    int64_t max_bytes = max_rows * words_per_row * 4;
    max_bytes -= max_bytes % 16;
    if (!rm->pending) {
        NEED(rm->nnz <= max_rows, "more rows than the source buffer holds");
        int64_t bytes = rm->nnz * words_per_row * 4;
        bytes += (16 - bytes % 16) % 16;
        if (bytes > max_bytes) bytes = max_bytes;
        return mrmp_exchange_push(x, ch, dst, bits_dev, bytes, stream);
    }
    if (rm->pend_rows_cap < 0 || rm->pend_rows_cap > max_rows) rm->pend_rows_cap = max_rows;
    return mrmp_exchange_push_counted(x, ch, dst, bits_dev, (const long long *)rm->nnz_d,
                                      (int64_t)words_per_row * 4, max_bytes, stream);
}

extern "C" int mrmp_roadmaps_nnz(mrmp_roadmaps *rm, int64_t *nnz_total_out) {
    NEED(rm && nnz_total_out, "bad arguments");
    NEED(rm->built, "roadmaps not built");
    NEED(!rm->pending, "asynchronous build pending: call mrmp_roadmaps_finish first");
    *nnz_total_out = rm->nnz;
    return MRMP_OK;
}

extern "C" int mrmp_roadmaps_get_nodes(mrmp_roadmaps *rm, double *nodes_out) {
    NEED(rm && nodes_out, "bad arguments");
    NEED(rm->nv >= 1, "no nodes set");
    mrmp_ctx *c = rm->ctx;
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(nodes_out, rm->nodes, sizeof(double) * c->d * (size_t)rm->nv * rm->n_agents,
                       cudaMemcpyDeviceToHost, c->s_k));
    CU(cudaStreamSynchronize(c->s_k));
    return roadmaps_settle(rm, true);
}

extern "C" int mrmp_roadmaps_get_csr(mrmp_roadmaps *rm, int32_t *row_ptr, int32_t *col_idx,
                                     int64_t col_cap, int64_t *nnz_out) {
    NEED(rm && row_ptr && nnz_out, "bad arguments");
    NEED(rm->built, "roadmaps not built");
    NEED(!rm->pending, "asynchronous build pending: call mrmp_roadmaps_finish first");
    mrmp_ctx *c = rm->ctx;
    CU(cudaSetDevice(c->device));
    *nnz_out = rm->nnz;
    const int64_t rows = (int64_t)rm->n_agents * rm->nv;
    CU(cudaMemcpyAsync(row_ptr, rm->row_ptr, sizeof(int32_t) * (rows + 1), cudaMemcpyDeviceToHost,
                       c->s_k));
    if (col_cap < rm->nnz) {
        CU(cudaStreamSynchronize(c->s_k));
        return fail(MRMP_ERR_CAPACITY, "col_cap %lld < nnz %lld", (long long)col_cap,
                    (long long)rm->nnz);
    }
    NEED(col_idx || rm->nnz == 0, "col_idx is NULL");
    if (rm->nnz > 0)
        CU(cudaMemcpyAsync(col_idx, rm->col_idx, sizeof(int32_t) * rm->nnz, cudaMemcpyDeviceToHost,
                           c->s_out));
    CU(cudaStreamSynchronize(c->s_k));
    CU(cudaStreamSynchronize(c->s_out));
    return MRMP_OK;
}

/* Pin (page-lock and map) / unpin a host buffer of the caller, so that host code without a CUDA
 * binding of its own (the reference's Julia) gets the asynchronous read-back and the in-place
 * table output for its ordinary arrays. */
extern "C" int mrmp_host_register(void *ptr, size_t bytes) {
    NEED(ptr && bytes > 0, "bad buffer");
    CU(cudaHostRegister(ptr, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
    return MRMP_OK;
}

extern "C" int mrmp_host_unregister(void *ptr) {
    NEED(ptr, "bad buffer");
    CU(cudaHostUnregister(ptr));
    return MRMP_OK;
}

extern "C" int mrmp_roadmaps_read_async(mrmp_roadmaps *rm, double *nodes_out, int32_t *row_ptr,
                                        int32_t *col_idx, int64_t col_cap, int64_t *nnz_out) {
    NEED(rm && nodes_out && row_ptr && nnz_out && col_cap >= 0, "bad arguments");
    NEED(rm->built, "roadmaps not built");
    mrmp_ctx *c = rm->ctx;
    CU(cudaSetDevice(c->device));
    if (rm->pending) {
        // Behind an asynchronous build (call it right after mrmp_roadmaps_build_async): the copies
        // wait for what the ctx's stream holds now and then run beside whatever is launched next
        // (the conflict table), instead of after the step's host wait.  The edge count is not on
        // the host yet: all col_cap entries (bounded by the CSR buffer) are copied, *nnz_out is
        // -1 and mrmp_roadmaps_finish reports MRMP_ERR_CAPACITY if nnz exceeded col_cap.
        NEED(col_idx || col_cap == 0, "col_idx is NULL");
        const int64_t rows = (int64_t)rm->n_agents * rm->nv;
        const int64_t n_col = col_cap < rm->col_cap ? col_cap : rm->col_cap;
        CU(cudaEventRecord(c->ev_read, c->s_k));
        CU(cudaStreamWaitEvent(c->s_out, c->ev_read, 0));
        CU(cudaMemcpyAsync(nodes_out, rm->nodes, sizeof(double) * c->d * (size_t)rows,
                           cudaMemcpyDeviceToHost, c->s_out));
        CU(cudaMemcpyAsync(row_ptr, rm->row_ptr, sizeof(int32_t) * (rows + 1),
                           cudaMemcpyDeviceToHost, c->s_out));
        if (n_col > 0)
            CU(cudaMemcpyAsync(col_idx, rm->col_idx, sizeof(int32_t) * n_col, cudaMemcpyDeviceToHost,
                               c->s_out));
        rm->pend_read_cap = col_cap;
        *nnz_out = -1;
        return MRMP_OK;
    }
    *nnz_out = rm->nnz;
    if (col_cap < rm->nnz)
        return fail(MRMP_ERR_CAPACITY, "col_cap %lld < nnz %lld", (long long)col_cap,
                    (long long)rm->nnz);
    NEED(col_idx || rm->nnz == 0, "col_idx is NULL");
    // build() has synchronised the kernel stream, so the copy stream may start at once and
    // runs beside whatever the caller launches next
    const int64_t rows = (int64_t)rm->n_agents * rm->nv;
    CU(cudaMemcpyAsync(nodes_out, rm->nodes, sizeof(double) * c->d * (size_t)rows,
                       cudaMemcpyDeviceToHost, c->s_out));
    CU(cudaMemcpyAsync(row_ptr, rm->row_ptr, sizeof(int32_t) * (rows + 1), cudaMemcpyDeviceToHost,
                       c->s_out));
    if (rm->nnz > 0)
        CU(cudaMemcpyAsync(col_idx, rm->col_idx, sizeof(int32_t) * rm->nnz, cudaMemcpyDeviceToHost,
                           c->s_out));
    return MRMP_OK;
}

extern "C" int mrmp_roadmaps_read_wait(mrmp_roadmaps *rm) {
    NEED(rm, "bad roadmaps");
    CU(cudaSetDevice(rm->ctx->device));
    CU(cudaStreamSynchronize(rm->ctx->s_out));
    return roadmaps_settle(rm, false);
}

extern "C" int mrmp_roadmaps_device_ptrs(mrmp_roadmaps *rm, int *nv_out, double **nodes_dev,
                                         int32_t **row_ptr_dev, int32_t **col_idx_dev,
                                         int64_t *nnz_out) {
    NEED(rm, "bad roadmaps");
    {
        int rc = roadmaps_settle(rm, false);
        if (rc) return rc;
    }
    if (nv_out) *nv_out = rm->nv;
    if (nodes_dev) *nodes_dev = rm->nodes;
    if (row_ptr_dev) *row_ptr_dev = rm->row_ptr;
    if (col_idx_dev) *col_idx_dev = rm->col_idx;
    if (nnz_out) *nnz_out = rm->built ? rm->nnz : -1;
    return MRMP_OK;
}

extern "C" int mrmp_prm_build(mrmp_ctx *c, int agent0, int n_agents, int nv, const double *rad,
                              const double *nodes, int32_t *row_ptr, int32_t *col_idx,
                              int64_t col_cap, int64_t *nnz_out) {
    NEED(c && rad && nodes && row_ptr && nnz_out, "bad arguments");
    mrmp_roadmaps *rm = nullptr;
    int rc = mrmp_roadmaps_create(c, agent0, n_agents, nv < 2 ? 2 : nv, &rm);
    if (rc) return rc;
    rc = mrmp_roadmaps_set_nodes(rm, nv, nodes);
    if (!rc) rc = mrmp_roadmaps_build(rm, rad);
    if (!rc) rc = mrmp_roadmaps_get_csr(rm, row_ptr, col_idx, col_cap, nnz_out);
    mrmp_roadmaps_destroy(rm);
    return rc;
}

extern "C" int mrmp_prms(mrmp_ctx *c, int agent0, int n_agents, int nv, const double *rad,
                         uint64_t seed, const double *q_init, const double *q_goal,
                         double *nodes_out, int32_t *row_ptr, int32_t *col_idx, int64_t col_cap,
                         int64_t *nnz_out) {
    NEED(c && rad && q_init && q_goal && nodes_out && row_ptr && nnz_out, "bad arguments");
    mrmp_roadmaps *rm = nullptr;
    int rc = mrmp_roadmaps_create(c, agent0, n_agents, nv, &rm);
    if (rc) return rc;
    rc = mrmp_roadmaps_sample(rm, nv, seed, q_init, q_goal, nullptr);
    if (!rc) rc = mrmp_roadmaps_build(rm, rad);
    if (!rc) rc = mrmp_roadmaps_get_nodes(rm, nodes_out);
    if (!rc) rc = mrmp_roadmaps_get_csr(rm, row_ptr, col_idx, col_cap, nnz_out);
    mrmp_roadmaps_destroy(rm);
    return rc;
}

// ---- collide over roadmap edges by id ----------------------------------------------------------
static int check_rm_agents(const mrmp_roadmaps *rm, int64_t n, const int32_t *a) {
    for (int64_t k = 0; k < n; ++k)
        if (a[k] < rm->agent0 || a[k] >= rm->agent0 + rm->n_agents)
            return fail(MRMP_ERR_INVALID, "agent %d not held by this roadmap set", a[k]);
    return MRMP_OK;
}
static int check_vertices(const mrmp_roadmaps *rm, int64_t n, const int32_t *v) {
    for (int64_t k = 0; k < n; ++k)
        if ((unsigned)v[k] >= (unsigned)rm->nv)
            return fail(MRMP_ERR_INVALID, "vertex id %d out of range at %lld", v[k], (long long)k);
    return MRMP_OK;
}

extern "C" int mrmp_collide_edges_indexed(mrmp_roadmaps *rm, int64_t n, const int32_t *i,
                                          const int32_t *j, const int32_t *vi_from,
                                          const int32_t *vi_to, const int32_t *vj_from,
                                          const int32_t *vj_to, uint8_t *out) {
    NEED(rm && n >= 0, "bad roadmaps / n");
    if (n == 0) return MRMP_OK;
    NEED(i && j && vi_from && vi_to && vj_from && vj_to && out, "NULL buffer");
    NEED(rm->nv >= 1, "no nodes set");
    int rc = check_rm_agents(rm, n, i);
    if (!rc) rc = check_rm_agents(rm, n, j);
    if (!rc) rc = check_vertices(rm, n, vi_from);
    if (!rc) rc = check_vertices(rm, n, vi_to);
    if (!rc) rc = check_vertices(rm, n, vj_from);
    if (!rc) rc = check_vertices(rm, n, vj_to);
    if (rc) return rc;
    mrmp_ctx *c = rm->ctx;
    std::vector<Arr> arrs = {{i, nullptr, 4},       {j, nullptr, 4},     {vi_from, nullptr, 4},
                             {vi_to, nullptr, 4},   {vj_from, nullptr, 4}, {vj_to, nullptr, 4},
                             {nullptr, out, 1}};
    return run_batch(c, n, arrs, [&](void **d, int64_t len, cudaStream_t s) {
        return launch_collide_edges_indexed(
            c->cv, lcfg(c, s), len, (const int32_t *)d[0], (const int32_t *)d[1],
            (const int32_t *)d[2], (const int32_t *)d[3], (const int32_t *)d[4],
            (const int32_t *)d[5], rm->nodes, rm->agent0, rm->n_agents, rm->nv, (uint8_t *)d[6]);
    });
}

extern "C" int mrmp_collide_edges_indexed_dev(mrmp_roadmaps *rm, int64_t n, const int32_t *i,
                                              const int32_t *j, const int32_t *vi_from,
                                              const int32_t *vi_to, const int32_t *vj_from,
                                              const int32_t *vj_to, uint8_t *out, void *stream) {
    NEED(rm && n >= 0, "bad roadmaps / n");
    if (n == 0) return MRMP_OK;
    mrmp_ctx *c = rm->ctx;
    CU(cudaSetDevice(c->device));
    CU(order_after_ctx(c, (cudaStream_t)stream));
    CU(launch_collide_edges_indexed(c->cv, lcfg(c, (cudaStream_t)stream), n, i, j, vi_from, vi_to,
                                    vj_from, vj_to, rm->nodes, rm->agent0, rm->n_agents, rm->nv,
                                    out));
    return MRMP_OK;
}

// ------------------------------------------------------------------------------------------
// cross-product collide (rows x columns -> bits)
// ------------------------------------------------------------------------------------------
static int cross_words(int64_t n_cols, int group_size) {
    const int64_t bits = group_size > 0 ? (n_cols + group_size - 1) / group_size : n_cols;
    return (int)((bits + 31) / 32);
}

// rows / cols / out are device pointers.  rows_sorted: Morton-sorted row records prepared
// earlier (roadmap edges cache them) or NULL (sorted here into scratch slot 4).  Column tiles,
// their headers and the sort workspace live in scratch slot 0.
// rq.mode: bits / or -> `out` is uint32 [n_rows][wpr]; count / first -> int32 [n_rows][n_groups].
static int cross_run_dev(mrmp_ctx *c, int64_t n_rows, const int32_t *row_agent,
                         const double *row_from, const double *row_to, const void *rows_sorted,
                         int64_t n_cols, const int32_t *col_agent, const double *col_from,
                         const double *col_to, const int32_t *col_vf, const int32_t *col_vt,
                         const double *nodes, int agent0, int nv_stride, const CrossReq &rq,
                         uint32_t *out, cudaStream_t s, const long long *n_rows_dev = nullptr,
                         bool cols_forked = false) {
    // cols_forked: ev_fork was recorded on s before the caller queued its row sort there; the
    // column tiles (which depend only on what precedes the fork) are prepared on s_aux meanwhile
    const int group_size = rq.group_size;
    const bool table = rq.mode == kCrossBits || rq.mode == kCrossOr;
    const int wpr = table ? cross_words(n_cols, group_size) : rq.n_groups;
    LaunchCfg lc = lcfg(c, s);
    // An OR table may be written with plain stores straight into mapped host memory or a peer's
    // memory (one work item per row).  When the output is ordinary memory of this GPU, a small
    // table is cut into finer work items that OR into it (kernels_cross.cu: cross_split).
    bool local_out = false;
    {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, out) == cudaSuccess)
            local_out = at.type == cudaMemoryTypeDevice && at.device == c->device;
        else
            cudaGetLastError();
    }
    const int n_tiles_all = (int)((n_cols + kCrossCols - 1) / kCrossCols);
    const bool split_or = rq.mode == kCrossOr && c->model <= MRMP_MODEL_POINT3D && local_out &&
                          cross_split(lc, n_rows, n_tiles_all, 1) > 1;
    if (rq.mode == kCrossFirst)
        CU(launch_fill_i32(lc, (int32_t *)out, n_rows * (int64_t)wpr, 0x7fffffff));
    else if (rq.mode != kCrossOr || c->model >= MRMP_MODEL_ARM22 || split_or)
        // bits / count: hits are OR-ed / added into a zeroed output (the un-split OR mode of the
        // point models writes every word itself)
        CU(cudaMemsetAsync(out, 0, sizeof(uint32_t) * (size_t)n_rows * wpr, s));
    if (c->model >= MRMP_MODEL_ARM22) {  // discretised models: warp per (row, column) check
        if (c->model == MRMP_MODEL_ARM22)
            CU(launch_arm_collide_cross(c->cv, lc, n_rows, row_agent, row_from, row_to, n_cols,
                                        col_agent, col_from, col_to, col_vf, col_vt, nodes, agent0,
                                        nv_stride, rq, wpr, out, c->stats_d + 1));
        else
            CU(launch_chain_collide_cross(c->cv, lc, n_rows, row_agent, row_from, row_to, n_cols,
                                          col_agent, col_from, col_to, col_vf, col_vt, nodes,
                                          agent0, nv_stride, rq, wpr, out, c->stats_d + 1));
        if (rq.mode == kCrossFirst)
            CU(launch_cross_first_finalize(lc, (int32_t *)out, n_rows * (int64_t)wpr));
        return MRMP_OK;
    }
    // columns per launch: everything, or kCrossMaxGroups groups in OR-reduced mode
    const int64_t chunk_cols =
        rq.mode == kCrossOr ? (int64_t)kCrossMaxGroups * group_size : n_cols;
    const int64_t max_cols = n_cols < chunk_cols ? n_cols : chunk_cols;
    const int64_t max_tiles = (max_cols + kCrossCols - 1) / kCrossCols;
    const size_t o_cells = 0, o_work = o_cells + 4096 * 4, o_order = o_work + 16,
                 o_hdr = align16(o_order + 4 * (size_t)max_cols),
                 o_tiles = align16(o_hdr + cross_hdr_bytes(c->d) * (size_t)max_tiles),
                 tot = o_tiles + cross_tile_bytes(c->d) * (size_t)max_tiles;
    int rc = ensure_scratch(c, 0, tot);
    if (rc) return rc;
    unsigned char *b = c->scratch[0];
    if (!rows_sorted) {
        rc = ensure_scratch(c, 4, cross_row_bytes(c->d) * (size_t)n_rows);
        if (rc) return rc;
        CU(launch_cross_sort_rows(c->cv, lc, n_rows, row_agent, row_from, row_to,
                                  (int32_t *)(b + o_cells), c->scratch[4]));
        rows_sorted = c->scratch[4];
    }
    const int d = c->d;
    for (int64_t c0 = 0; c0 < n_cols; c0 += chunk_cols) {
        const int64_t nc = n_cols - c0 < chunk_cols ? n_cols - c0 : chunk_cols;
        CrossCols cc;
        cc.group_size = rq.mode == kCrossBits ? 0 : group_size;
        cc.col_group = rq.col_group ? rq.col_group + c0 : nullptr;
        cc.col_key = rq.col_key ? rq.col_key + c0 : nullptr;
        cc.orig_base = c0;
        const bool forked = cols_forked && nc == n_cols;  // one chunk only: the tiles are not reused
        LaunchCfg lcp = lc;
        if (forked) {
            lcp.stream = c->s_aux;
            CU(cudaStreamWaitEvent(c->s_aux, c->ev_fork, 0));
        }
        CU(launch_cross_prep_cols(c->cv, lcp, nc, col_agent + c0,
                                  col_from ? col_from + c0 * d : nullptr,
                                  col_to ? col_to + c0 * d : nullptr, col_vf ? col_vf + c0 : nullptr,
                                  col_vt ? col_vt + c0 : nullptr, nodes, agent0, nv_stride, cc,
                                  (int32_t *)(b + o_cells), (int32_t *)(b + o_order), b + o_tiles,
                                  b + o_hdr));
        if (forked) {
            CU(cudaEventRecord(c->ev_join, c->s_aux));
            CU(cudaStreamWaitEvent(s, c->ev_join, 0));
        }
        const int64_t g0 = rq.mode == kCrossOr ? c0 / group_size : 0;
        CrossOut co;
        co.mode = rq.mode;
        co.words_per_row = rq.mode == kCrossOr ? cross_words(nc, group_size) : wpr;
        co.out_stride = wpr;
        co.out = out + g0 / 32;
        co.row_key = rq.row_key;
        co.cbs = rq.cbs;
        co.may_split = rq.mode == kCrossOr ? (split_or ? 1 : 0) : (local_out ? 1 : 0);
        co.n_rows_dev = n_rows_dev;
        CU(launch_cross_collide(c->cv, lc, n_rows, rows_sorted, nc, b + o_tiles, b + o_hdr, co,
                                c->stats_d, (unsigned int *)(b + o_work)));
    }
    if (rq.mode == kCrossFirst)
        CU(launch_cross_first_finalize(lc, (int32_t *)out, n_rows * (int64_t)wpr));
    return MRMP_OK;
}

static CrossReq table_req(int group_size) {
    CrossReq rq;
    rq.mode = group_size > 0 ? kCrossOr : kCrossBits;
    rq.group_size = group_size;
    rq.n_groups = 0;
    rq.col_group = rq.col_key = rq.row_key = nullptr;
    rq.cbs = 0;
    return rq;
}

extern "C" int mrmp_collide_cross_dev(mrmp_ctx *c, int64_t n_rows, const int32_t *row_agent,
                                      const double *row_from, const double *row_to,
                                      int64_t n_cols, const int32_t *col_agent,
                                      const double *col_from, const double *col_to,
                                      int group_size, uint32_t *out_bits, void *stream) {
    NEED(c && n_rows >= 0 && n_cols >= 0 && group_size >= 0, "bad arguments");
    if (n_rows == 0 || n_cols == 0) return MRMP_OK;
    CU(cudaSetDevice(c->device));
    if (need_aligned(c, row_from) || need_aligned(c, row_to) || need_aligned(c, col_from) ||
        need_aligned(c, col_to))
        return MRMP_ERR_INVALID;
    return cross_run_dev(c, n_rows, row_agent, row_from, row_to, nullptr, n_cols, col_agent,
                         col_from, col_to, nullptr, nullptr, nullptr, 0, 0, table_req(group_size),
                         out_bits, (cudaStream_t)stream);
}

extern "C" int mrmp_collide_cross(mrmp_ctx *c, int64_t n_rows, const int32_t *row_agent,
                                  const double *row_from, const double *row_to, int64_t n_cols,
                                  const int32_t *col_agent, const double *col_from,
                                  const double *col_to, int group_size, uint32_t *out_bits) {
    NEED(c && n_rows >= 0 && n_cols >= 0 && group_size >= 0, "bad arguments");
    if (n_rows == 0 || n_cols == 0) return MRMP_OK;
    NEED(row_agent && row_from && row_to && col_agent && col_from && col_to && out_bits,
         "NULL buffer");
    int rc = check_agents_host(c, n_rows, row_agent);
    if (!rc) rc = check_agents_host(c, n_cols, col_agent);
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    const size_t sb = sizeof(double) * c->d;
    const int wpr = cross_words(n_cols, group_size);
    // scratch slots: 1 rows+cols staging, 2 output bits
    const size_t o_ra = 0, o_rf = align16(o_ra + 4 * (size_t)n_rows),
                 o_rt = align16(o_rf + sb * (size_t)n_rows),
                 o_ca = align16(o_rt + sb * (size_t)n_rows),
                 o_cf = align16(o_ca + 4 * (size_t)n_cols),
                 o_ct = align16(o_cf + sb * (size_t)n_cols),
                 tot = align16(o_ct + sb * (size_t)n_cols);
    rc = ensure_scratch(c, 1, tot);
    if (!rc) rc = ensure_scratch(c, 2, sizeof(uint32_t) * (size_t)n_rows * wpr);
    if (rc) return rc;
    unsigned char *b = c->scratch[1];
    cudaStream_t s = c->s_k;
    CU(cudaMemcpyAsync(b + o_ra, row_agent, 4 * (size_t)n_rows, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(b + o_rf, row_from, sb * (size_t)n_rows, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(b + o_rt, row_to, sb * (size_t)n_rows, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(b + o_ca, col_agent, 4 * (size_t)n_cols, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(b + o_cf, col_from, sb * (size_t)n_cols, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(b + o_ct, col_to, sb * (size_t)n_cols, cudaMemcpyHostToDevice, s));
    rc = cross_run_dev(c, n_rows, (const int32_t *)(b + o_ra), (const double *)(b + o_rf),
                       (const double *)(b + o_rt), nullptr, n_cols, (const int32_t *)(b + o_ca),
                       (const double *)(b + o_cf), (const double *)(b + o_ct), nullptr, nullptr,
                       nullptr, 0, 0, table_req(group_size), (uint32_t *)c->scratch[2], s);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out_bits, c->scratch[2], sizeof(uint32_t) * (size_t)n_rows * wpr,
                       cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return MRMP_OK;
}

// ---- count / first reductions (cbs.jl:336-347, smoother.jl:78-106) ---------------------------
static int reduce_req(mrmp_ctx *c, int64_t n_cols, const int32_t *col_group, const int32_t *col_key,
                      const int32_t *row_key, int group_size, int n_groups, int mode, int flags,
                      CrossReq &rq) {
    NEED(mode == MRMP_CROSS_COUNT || mode == MRMP_CROSS_FIRST, "mode must be COUNT or FIRST");
    NEED((row_key == nullptr) == (col_key == nullptr), "row_key and col_key go together");
    NEED(n_groups >= 1 && n_groups <= (1 << 24), "n_groups out of range");
    NEED(col_group || (group_size > 0 && (n_cols + group_size - 1) / group_size <= n_groups),
         "group_size / n_groups do not cover the columns");
    NEED((flags & ~MRMP_CROSS_FLAG_CBS) == 0, "unknown flags");
    (void)c;
    rq.mode = mode == MRMP_CROSS_COUNT ? kCrossCount : kCrossFirst;
    rq.group_size = col_group ? 0 : group_size;
    rq.n_groups = n_groups;
    rq.col_group = col_group;
    rq.col_key = col_key;
    rq.row_key = row_key;
    rq.cbs = (flags & MRMP_CROSS_FLAG_CBS) ? 1 : 0;
    return MRMP_OK;
}

extern "C" int mrmp_collide_cross_reduce_dev(mrmp_ctx *c, int64_t n_rows, const int32_t *row_agent,
                                             const double *row_from, const double *row_to,
                                             const int32_t *row_key, int64_t n_cols,
                                             const int32_t *col_agent, const double *col_from,
                                             const double *col_to, const int32_t *col_group,
                                             const int32_t *col_key, int group_size, int n_groups,
                                             int mode, int flags, int32_t *out, void *stream) {
    NEED(c && n_rows >= 0 && n_cols >= 0 && out, "bad arguments");
    CrossReq rq;
    int rc = reduce_req(c, n_cols, col_group, col_key, row_key, group_size, n_groups, mode, flags, rq);
    if (rc) return rc;
    if (n_rows == 0) return MRMP_OK;
    CU(cudaSetDevice(c->device));
    if (n_cols == 0) {  // nothing collides: 0 counts / -1 first hits
        CU(launch_fill_i32(lcfg(c, (cudaStream_t)stream), out, n_rows * (int64_t)n_groups,
                           mode == MRMP_CROSS_FIRST ? -1 : 0));
        return MRMP_OK;
    }
    if (need_aligned(c, row_from) || need_aligned(c, row_to) || need_aligned(c, col_from) ||
        need_aligned(c, col_to))
        return MRMP_ERR_INVALID;
    return cross_run_dev(c, n_rows, row_agent, row_from, row_to, nullptr, n_cols, col_agent,
                         col_from, col_to, nullptr, nullptr, nullptr, 0, 0, rq, (uint32_t *)out,
                         (cudaStream_t)stream);
}

extern "C" int mrmp_collide_cross_reduce(mrmp_ctx *c, int64_t n_rows, const int32_t *row_agent,
                                         const double *row_from, const double *row_to,
                                         const int32_t *row_key, int64_t n_cols,
                                         const int32_t *col_agent, const double *col_from,
                                         const double *col_to, const int32_t *col_group,
                                         const int32_t *col_key, int group_size, int n_groups,
                                         int mode, int flags, int32_t *out) {
    NEED(c && n_rows >= 0 && n_cols >= 0, "bad arguments");
    if (n_rows == 0) return MRMP_OK;
    NEED(out && row_agent && row_from && row_to, "NULL buffer");
    NEED(n_cols == 0 || (col_agent && col_from && col_to), "NULL buffer");
    CrossReq rq;
    int rc = reduce_req(c, n_cols, col_group, col_key, row_key, group_size, n_groups, mode, flags, rq);
    if (rc) return rc;
    rc = check_agents_host(c, n_rows, row_agent);
    if (!rc) rc = check_agents_host(c, n_cols, col_agent);
    if (rc) return rc;
    if (col_group)
        for (int64_t k = 0; k < n_cols; ++k)
            NEED(col_group[k] >= 0 && col_group[k] < n_groups, "col_group out of range");
    CU(cudaSetDevice(c->device));
    const size_t sb = sizeof(double) * c->d;
    const size_t o_ra = 0, o_rf = align16(o_ra + 4 * (size_t)n_rows),
                 o_rt = align16(o_rf + sb * (size_t)n_rows),
                 o_rk = align16(o_rt + sb * (size_t)n_rows),
                 o_ca = align16(o_rk + 4 * (size_t)n_rows),
                 o_cf = align16(o_ca + 4 * (size_t)n_cols),
                 o_ct = align16(o_cf + sb * (size_t)n_cols),
                 o_cg = align16(o_ct + sb * (size_t)n_cols),
                 o_ck = align16(o_cg + 4 * (size_t)n_cols),
                 tot = align16(o_ck + 4 * (size_t)n_cols);
    const size_t ob = sizeof(int32_t) * (size_t)n_rows * n_groups;
    rc = ensure_scratch(c, 1, tot);
    if (!rc) rc = ensure_scratch(c, 2, ob);
    if (rc) return rc;
    unsigned char *b = c->scratch[1];
    cudaStream_t s = c->s_k;
    const auto up = [&](size_t off, const void *src, size_t bytes) {
        return bytes ? cudaMemcpyAsync(b + off, src, bytes, cudaMemcpyHostToDevice, s) : cudaSuccess;
    };
    CU(up(o_ra, row_agent, 4 * (size_t)n_rows));
    CU(up(o_rf, row_from, sb * (size_t)n_rows));
    CU(up(o_rt, row_to, sb * (size_t)n_rows));
    if (row_key) CU(up(o_rk, row_key, 4 * (size_t)n_rows));
    CU(up(o_ca, col_agent, 4 * (size_t)n_cols));
    CU(up(o_cf, col_from, sb * (size_t)n_cols));
    CU(up(o_ct, col_to, sb * (size_t)n_cols));
    if (col_group) CU(up(o_cg, col_group, 4 * (size_t)n_cols));
    if (col_key) CU(up(o_ck, col_key, 4 * (size_t)n_cols));
    rc = mrmp_collide_cross_reduce_dev(
        c, n_rows, (const int32_t *)(b + o_ra), (const double *)(b + o_rf),
        (const double *)(b + o_rt), row_key ? (const int32_t *)(b + o_rk) : nullptr, n_cols,
        (const int32_t *)(b + o_ca), (const double *)(b + o_cf), (const double *)(b + o_ct),
        col_group ? (const int32_t *)(b + o_cg) : nullptr,
        col_key ? (const int32_t *)(b + o_ck) : nullptr, group_size, n_groups, mode, flags,
        (int32_t *)c->scratch[2], s);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out, c->scratch[2], ob, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return MRMP_OK;
}

// rows = every CSR edge of the roadmap set (packed CSR order); columns = edges of roadmap
// agents given by (agent, v_from, v_to) ids
static int roadmaps_expand_edges(mrmp_roadmaps *rm, cudaStream_t s) {
    if (rm->e_valid) return MRMP_OK;
    mrmp_ctx *c = rm->ctx;
    const bool point = c->model <= MRMP_MODEL_POINT3D;
    if (rm->pending) NEED(point, "an asynchronous build of this model must be finished first");
    const int64_t need = rm->pending ? rm->col_cap : rm->nnz;  // pending: bounded by the CSR buffer
    if (need > rm->e_cap) {
        CU(cudaStreamSynchronize(s));
        cudaFree(rm->e_agent);
        cudaFree(rm->e_from);
        cudaFree(rm->e_to);
        cudaFree(rm->e_rows);
        rm->e_agent = nullptr;
        rm->e_from = rm->e_to = nullptr;
        rm->e_rows = nullptr;
        rm->e_cap = need + need / 8 + 64;
        if (point) {
            // point models: the cross kernel reads Morton-sorted row records only
            CU(cudaMalloc(&rm->e_rows, cross_row_bytes(c->d) * (size_t)rm->e_cap));
        } else {
            CU(cudaMalloc((void **)&rm->e_agent, sizeof(int32_t) * rm->e_cap));
            CU(cudaMalloc((void **)&rm->e_from, sizeof(double) * c->d * rm->e_cap));
            CU(cudaMalloc((void **)&rm->e_to, sizeof(double) * c->d * rm->e_cap));
        }
    }
    if (point) {
        int rc = ensure_scratch(c, 6, 2 * 4096 * 4 + 64);
        if (rc) return rc;
        CU(launch_cross_sort_csr_rows(c->cv, lcfg(c, s), (int64_t)rm->n_agents * rm->nv, rm->nv,
                                      rm->agent0, rm->row_ptr, rm->col_idx, rm->nodes,
                                      (int32_t *)c->scratch[6], rm->e_rows,
                                      rm->col_cap < rm->e_cap ? rm->col_cap : rm->e_cap));
    } else {
        CU(launch_csr_expand_edges(c->cv, lcfg(c, s), (int64_t)rm->n_agents * rm->nv, rm->nv,
                                   rm->agent0, rm->row_ptr, rm->col_idx, rm->nodes, rm->e_agent,
                                   rm->e_from, rm->e_to));
    }
    rm->e_valid = true;
    return MRMP_OK;
}

// out_cap_rows < 0: the set is built (nnz known); >= 0: room of `out` in rows, the set may be pending
static int edge_conflicts_core(mrmp_roadmaps *rm, mrmp_roadmaps *rm_cols, int64_t n_cols,
                               const int32_t *col_agent, const int32_t *col_vfrom,
                               const int32_t *col_vto, int group_size, uint32_t *out_bits,
                               int64_t out_cap_rows, cudaStream_t s) {
    NEED(rm && n_cols >= 0 && group_size >= 0, "bad arguments");
    NEED(rm->built, "roadmaps not built");
    if (!rm_cols) rm_cols = rm;
    NEED(rm_cols->ctx == rm->ctx && rm_cols->nv >= 1, "rm_cols must belong to the same ctx");
    mrmp_ctx *c = rm->ctx;
    int64_t n_rows = rm->nnz;
    const long long *n_rows_dev = nullptr;
    if (rm->pending) {
        NEED(out_cap_rows >= 0, "the set has a pending asynchronous build: finish it first");
        NEED(c->model <= MRMP_MODEL_POINT3D, "asynchronous tables are built for the point models");
        n_rows = out_cap_rows < rm->col_cap ? out_cap_rows : rm->col_cap;
        n_rows_dev = (const long long *)rm->nnz_d;
        rm->pend_rows_cap = out_cap_rows;
    } else if (out_cap_rows >= 0 && rm->nnz > out_cap_rows) {
        return fail(MRMP_ERR_CAPACITY, "out_bits has room for %lld rows, the set has %lld edges",
                    (long long)out_cap_rows, (long long)rm->nnz);
    }
    if (n_cols == 0 || n_rows == 0) return MRMP_OK;
    CU(cudaSetDevice(c->device));
    CU(order_after_ctx(c, s));  // the set (and rm_cols) were sampled / built on the ctx's stream
    // the row sort (this stream) and the column tiles (s_aux) both read the finished roadmaps only
    const bool fork = c->table_overlap && !rm->e_valid && c->model <= MRMP_MODEL_POINT3D;
    if (fork) CU(cudaEventRecord(c->ev_fork, s));
    int rc = roadmaps_expand_edges(rm, s);
    if (rc) return rc;
    return cross_run_dev(c, n_rows, rm->e_agent, rm->e_from, rm->e_to, rm->e_rows, n_cols,
                         col_agent, nullptr, nullptr, col_vfrom, col_vto, rm_cols->nodes,
                         rm_cols->agent0, rm_cols->nv, table_req(group_size), out_bits, s,
                         n_rows_dev, fork);
}

extern "C" int mrmp_roadmaps_edge_conflicts_dev(mrmp_roadmaps *rm, mrmp_roadmaps *rm_cols,
                                                int64_t n_cols, const int32_t *col_agent,
                                                const int32_t *col_vfrom, const int32_t *col_vto,
                                                int group_size, uint32_t *out_bits,
                                                void *stream) {
    return edge_conflicts_core(rm, rm_cols, n_cols, col_agent, col_vfrom, col_vto, group_size,
                               out_bits, -1, (cudaStream_t)stream);
}

/* the same behind an asynchronous build: the row count stays on the device; out_bits has room for
 * out_cap_rows rows (checked by mrmp_roadmaps_finish) */
extern "C" int mrmp_roadmaps_edge_conflicts_async(mrmp_roadmaps *rm, mrmp_roadmaps *rm_cols,
                                                  int64_t n_cols, const int32_t *col_agent,
                                                  const int32_t *col_vfrom,
                                                  const int32_t *col_vto, int group_size,
                                                  uint32_t *out_bits, int64_t out_cap_rows,
                                                  void *stream) {
    NEED(out_cap_rows >= 0, "out_cap_rows must be >= 0");
    return edge_conflicts_core(rm, rm_cols, n_cols, col_agent, col_vfrom, col_vto, group_size,
                               out_bits, out_cap_rows, (cudaStream_t)stream);
}

/* One step of the strong-scaled path on this rank, enqueued by ONE host call (a rank's share of a
 * small scenario is ~0.2 ms of GPU work: a dozen separate calls from an interpreted host cost as
 * much).  Order: [next epoch] -> draw all vertices -> neighbour pass of this rank's agents ->
 * shard into every rank's memory + flags -> table rows of this rank's agents -> rows into rank
 * table_dst's memory + flag -> [destination: wait for every rank's rows]; the adoption of the
 * peers' shards is queued on the ctx's second stream, so it does not sit between the table kernel
 * and the row transfer.  Everything stays pending until mrmp_roadmaps_finish(rm_shard) and
 * mrmp_roadmaps_finish(rm_all). */
extern "C" int mrmp_roadmaps_strong_step(mrmp_roadmaps *rm_all, mrmp_roadmaps *rm_shard,
                                         mrmp_exchange *x, int world, int advance, int nv,
                                         uint64_t seed, const double *q_init, const double *q_goal,
                                         const double *rad_shard, int cap_agents, int64_t cap_nnz,
                                         int64_t n_cols, const int32_t *col_agent,
                                         const int32_t *col_vfrom, const int32_t *col_vto,
                                         int group_size, uint32_t *tab_local, int words_per_row,
                                         int64_t tab_rows_cap, int table_dst, int wait_table) {
    NEED(rm_all && rm_shard && x && rad_shard && tab_local, "bad arguments");
    NEED(rm_all->ctx == rm_shard->ctx, "both sets must belong to one ctx");
    NEED(world >= 1 && world <= 32 && table_dst >= 0 && table_dst < world, "bad world / table_dst");
    mrmp_ctx *c = rm_all->ctx;
    CU(cudaSetDevice(c->device));
    cudaStream_t s = c->s_k;
    const unsigned all = world >= 32 ? 0xffffffffu : ((1u << world) - 1u);
    int rc = MRMP_OK;
    if (advance) {
        rc = mrmp_exchange_advance(x, 0);
        if (!rc) rc = mrmp_exchange_advance(x, 1);
    }
    if (!rc) rc = mrmp_roadmaps_sample(rm_all, nv, seed, q_init, q_goal, nullptr);
    if (!rc) rc = mrmp_roadmaps_build_async(rm_shard, rad_shard);
    if (!rc) rc = mrmp_roadmaps_push_csr_shard(rm_shard, x, 0, world, cap_agents, cap_nnz, s);
    if (rc) return rc;
    CU(cudaEventRecord(c->ev_aux, s));  // this rank's own slot and flag are on their way
    rc = mrmp_roadmaps_edge_conflicts_async(rm_shard, rm_all, n_cols, col_agent, col_vfrom, col_vto,
                                            group_size, tab_local, tab_rows_cap, s);
    if (!rc)
        rc = mrmp_roadmaps_push_table_rows(rm_shard, x, 1, table_dst, tab_local, words_per_row,
                                           tab_rows_cap, s);
    if (!rc) rc = mrmp_exchange_signal(x, 1, 1u << table_dst, s);
    if (rc) return rc;
    static const bool adopt_aux = [] {
        const char *e = getenv("MRMP_STRONG_ADOPT_AUX");
        return !(e && e[0] == '0');
    }();
    cudaStream_t s_adopt = s;
    if (adopt_aux) {  // (queued after the column tiles of the table, which use the same stream)
        CU(cudaStreamWaitEvent(c->s_aux, c->ev_aux, 0));
        s_adopt = c->s_aux;
    }
    rc = mrmp_roadmaps_adopt_csr_shards_async(rm_all, x, 0, world, cap_agents, cap_nnz, s_adopt);
    if (!rc && wait_table) rc = mrmp_exchange_wait(x, 1, all, s);
    return rc;
}

extern "C" int mrmp_roadmaps_edge_conflicts(mrmp_roadmaps *rm, mrmp_roadmaps *rm_cols,
                                            int64_t n_cols, const int32_t *col_agent,
                                            const int32_t *col_vfrom,
                                            const int32_t *col_vto, int group_size,
                                            uint32_t *out_bits, int64_t out_cap_words) {
    NEED(rm && n_cols >= 0 && group_size >= 0, "bad arguments");
    NEED(rm->built, "roadmaps not built");
    NEED(!rm->pending, "asynchronous build pending: call mrmp_roadmaps_finish first");
    if (n_cols == 0 || rm->nnz == 0) return MRMP_OK;
    NEED(col_agent && col_vfrom && col_vto && out_bits, "NULL buffer");
    if (!rm_cols) rm_cols = rm;
    int rc = check_rm_agents(rm_cols, n_cols, col_agent);
    if (!rc) rc = check_vertices(rm_cols, n_cols, col_vfrom);
    if (!rc) rc = check_vertices(rm_cols, n_cols, col_vto);
    if (rc) return rc;
    mrmp_ctx *c = rm->ctx;
    CU(cudaSetDevice(c->device));
    const int wpr = cross_words(n_cols, group_size);
    if (out_cap_words < rm->nnz * wpr)
        return fail(MRMP_ERR_CAPACITY, "out_bits needs %lld words", (long long)(rm->nnz * wpr));
    const size_t ib = align16(4 * (size_t)n_cols);
    rc = ensure_scratch(c, 1, 3 * ib);
    if (!rc) rc = ensure_scratch(c, 2, sizeof(uint32_t) * (size_t)rm->nnz * wpr);
    if (rc) return rc;
    unsigned char *b = c->scratch[1];
    cudaStream_t s = c->s_k;
    CU(cudaMemcpyAsync(b, col_agent, 4 * (size_t)n_cols, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(b + ib, col_vfrom, 4 * (size_t)n_cols, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(b + 2 * ib, col_vto, 4 * (size_t)n_cols, cudaMemcpyHostToDevice, s));
    // OR-reduced tables of the point models write each output word exactly once with plain
    // stores: when the caller's buffer is pinned (mapped) host memory the kernel writes it in
    // place, the result travels while the kernel runs and no copy follows
    uint32_t *direct = nullptr;
    static const bool allow_direct = [] {   // MRMP_TABLE_DIRECT=0: always stage in device memory
        const char *e = getenv("MRMP_TABLE_DIRECT");
        return !(e && e[0] == '0');
    }();
    if (allow_direct && group_size > 0 && c->model <= MRMP_MODEL_POINT3D) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, out_bits) == cudaSuccess &&
            at.type == cudaMemoryTypeHost && at.devicePointer)
            direct = (uint32_t *)at.devicePointer;
        else
            cudaGetLastError();  // clear the sticky error of an unregistered pointer
    }
    rc = mrmp_roadmaps_edge_conflicts_dev(rm, rm_cols, n_cols, (const int32_t *)b,
                                          (const int32_t *)(b + ib),
                                          (const int32_t *)(b + 2 * ib), group_size,
                                          direct ? direct : (uint32_t *)c->scratch[2], s);
    if (rc) return rc;
    if (!direct)
        CU(cudaMemcpyAsync(out_bits, c->scratch[2], sizeof(uint32_t) * (size_t)rm->nnz * wpr,
                           cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return MRMP_OK;
}
