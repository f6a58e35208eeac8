// abi_internal.cuh -- host-side state shared by the ABI translation units (abi.cu, abi_sssp.cu).
#pragma once
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <vector>

#include "../../include/mrmp_b200.h"
#include "mrmp_internal.cuh"

using namespace mrmp;

// records the thread-local error message and returns `code` (defined in abi.cu)
int mrmp_fail(int code, const char *fmt, ...);
#define fail mrmp_fail

#define CU(expr)                                                                              \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return fail(MRMP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                                  \
    } while (0)

#define NEED(cond, msg) \
    if (!(cond)) return fail(MRMP_ERR_INVALID, msg)

static inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }
static inline bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

// ------------------------------------------------------------------------------------------
// exact squared-threshold helpers (see geom.cuh header comment)
// ------------------------------------------------------------------------------------------
// min{ s : RN(sqrt(s)) >= thr }: then  sqrt(s) < thr  <=>  s < T2  for every double s >= 0.
inline double sq_lt_threshold(double thr) {
    if (!(thr > 0.0)) return 0.0;  // dist < thr never holds for thr <= 0 or NaN
    if (isinf(thr)) return INFINITY;
    double s = thr * thr;
    while (s > 0.0 && sqrt(s) >= thr) s = nextafter(s, -INFINITY);
    while (sqrt(s) < thr) s = nextafter(s, INFINITY);
    return s;
}
// max{ s : RN(sqrt(s)) <= rad }: then  sqrt(s) <= rad  <=>  s <= R2.
inline double sq_le_threshold(double rad) {
    if (isnan(rad)) return NAN;
    if (rad < 0.0) return -1.0;
    if (isinf(rad)) return INFINITY;
    double s = rad * rad;
    while (sqrt(s) <= rad) s = nextafter(s, INFINITY);
    while (s > 0.0 && sqrt(s) > rad) s = nextafter(s, -INFINITY);
    return s;
}
// min{ s : RN(sqrt(s)) > thr }: then  sqrt(s) > thr  <=>  s >= G2.  (unused helper kept tiny)

static const double kFilterPad = 9.5367431640625e-07;  // 2^-20, absolute slack of the broad phase

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
static const int kSlots = 3;
static const int64_t kChunk = 1 << 20;     // checks per pipeline chunk
static const int64_t kSmallBatch = 1 << 16;  // single-stream fast path below this
static const int64_t kTinyBatch = 256;       // zero-copy (mapped pinned memory) below this

struct mrmp_ctx {
    int device = 0;
    int model = 0, d = 2, n_agents = 0, n_obs = 0;
    int sm_count = 148;
    CtxView cv{};
    double *blob = nullptr;
    std::vector<double> h_rads;
    cudaStream_t s_k = nullptr, s_in = nullptr, s_out = nullptr;
    cudaStream_t s_k_owned = nullptr;  // s_k unless the caller installed its own stream
    cudaEvent_t ev_in[kSlots]{}, ev_k[kSlots]{}, ev_out[kSlots]{};
    // conflict table of a roadmap set: the column tiles are prepared on s_aux while the rows are
    // sorted on the caller's stream (edge_conflicts_core)
    cudaStream_t s_aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_aux = nullptr, ev_order = nullptr, ev_read = nullptr;
    bool table_overlap = true;  // MRMP_TABLE_OVERLAP=0 keeps everything on one stream
    // chunk slots (device) and small-batch staging (pinned host + device)
    unsigned char *slot_buf[kSlots]{};
    size_t slot_bytes = 0;
    unsigned char *small_h = nullptr, *small_d = nullptr;
    size_t small_bytes = 0;
    // named device scratch buffers (column tiles, staged rows/cols, output bits)
    static const int kScratch = 8;
    unsigned char *scratch[kScratch]{};
    size_t scratch_bytes[kScratch]{};
    double max_rad = 0.0;
    unsigned long long *stats_d = nullptr;  // [4]: cross-kernel survivors, reserved
};

static inline int pad2(int n) { return (n + 1) & ~1; }

inline int ensure_small(mrmp_ctx *c, size_t bytes) {
    if (bytes <= c->small_bytes) return MRMP_OK;
    size_t nb = c->small_bytes ? c->small_bytes : (1 << 16);
    while (nb < bytes) nb *= 2;
    if (c->small_h) CU(cudaFreeHost(c->small_h));
    if (c->small_d) CU(cudaFree(c->small_d));
    c->small_h = c->small_d = nullptr;
    c->small_bytes = 0;
    // mapped: tiny batches let the kernel read / write this block directly (no copy nodes)
    CU(cudaHostAlloc((void **)&c->small_h, nb, cudaHostAllocMapped));
    CU(cudaMalloc((void **)&c->small_d, nb));
    c->small_bytes = nb;
    return MRMP_OK;
}

inline int ensure_scratch(mrmp_ctx *c, int k, size_t bytes) {
    if (bytes <= c->scratch_bytes[k]) return MRMP_OK;
    if (c->scratch[k]) {
        CU(cudaStreamSynchronize(c->s_k));
        CU(cudaFree(c->scratch[k]));
    }
    c->scratch[k] = nullptr;
    c->scratch_bytes[k] = 0;
    size_t nb = bytes + bytes / 4 + 256;
    CU(cudaMalloc((void **)&c->scratch[k], nb));
    c->scratch_bytes[k] = nb;
    return MRMP_OK;
}


// A *_dev / *_async entry point that consumes a roadmap set on a stream other than the ctx's: the
// set's sampling / build kernels run on the ctx's stream, so the consumer is ordered behind them.
// (stream 0 and cudaStreamLegacy are the same stream)
static inline bool same_stream(cudaStream_t a, cudaStream_t b) {
    if (a == (cudaStream_t)0) a = cudaStreamLegacy;
    if (b == (cudaStream_t)0) b = cudaStreamLegacy;
    return a == b;
}
static inline cudaError_t order_after_ctx(mrmp_ctx *c, cudaStream_t s) {
    if (same_stream(s, c->s_k) || s == c->s_aux) return cudaSuccess;  // (s_aux: ordered by its user)
    cudaError_t e = cudaEventRecord(c->ev_order, c->s_k);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s, c->ev_order, 0);
    return e;
}

static inline LaunchCfg lcfg(const mrmp_ctx *c, cudaStream_t s) {
    return LaunchCfg{c->sm_count, s, c->stats_d};
}
