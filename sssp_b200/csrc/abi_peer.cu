// abi_peer.cu -- peer-memory exchange between the GPUs of one box (one process per GPU).
//
// The reference has no distributed code; what shards here is the agent axis
// (prm.jl:265-275 builds the agents' roadmaps independently, pp.jl:179-189 tests one agent's
// candidate edges against the others' paths).  The only data that has to move between GPUs is
//   (1) the CSR adjacency of each rank's agent block, so that every GPU holds all roadmaps, and
//   (2) the rows of a conflict table that a rank computed for its agents.
// Both are produced by kernels of this library, so they are WRITTEN STRAIGHT INTO THE PEERS'
// MEMORY over NVLink by the producing kernel (plain coalesced stores through a peer mapping),
// followed by one flag store per peer; the consumer spins on its local flag.  No collective
// library call, no staging copy, nothing on the host between producer and consumer.
//
// An exchange owns, on every rank, one device buffer
//     [channel][parity 0/1][source rank][slot bytes]   +   flags [channel][source rank]
// Rank s writes its slot of channel ch for the current epoch into the buffer of every
// destination rank d (mrmp_exchange_ptr(x, ch, d) -> address of slot s in d's buffer), then
// mrmp_exchange_signal stores the epoch number into flags[ch][s] of d.  mrmp_exchange_wait
// makes the stream wait until flags[ch][*] of the local buffer have all reached the epoch.
// Epochs only grow and slots are double buffered by epoch parity: a rank can be at most one
// epoch ahead of the slowest peer it waits for, so a slot is never overwritten while it is read.
//
// Peer mappings come from CUDA IPC handles (cudaIpcGetMemHandle / cudaIpcOpenMemHandle) that the
// host framework exchanges (torch.distributed in sssp_b200/peer.py); inside one process (tests:
// several emulated ranks on one GPU) raw pointers are connected instead.
#include "abi_internal.cuh"

struct mrmp_exchange {
    mrmp_ctx *ctx = nullptr;
    int world = 1, rank = 0, n_chan = 0;
    std::vector<int64_t> slot_bytes;   // per channel (16-byte multiples)
    std::vector<int64_t> chan_off;     // byte offset of the channel's [2][world][slot] block
    int64_t flags_off = 0, total = 0;
    unsigned char *buf = nullptr;      // this rank's buffer
    std::vector<unsigned char *> peer; // [world] base of every rank's buffer (own at [rank])
    std::vector<bool> opened;          // peer[r] came from cudaIpcOpenMemHandle
    std::vector<unsigned int> epoch;   // per channel, starts at 1
    int *err_h = nullptr, *err_d = nullptr;  // mapped: a wait that timed out reports here
    bool connected = false;
};

namespace {

struct PeerFlags {
    unsigned int *p[32];
};

// lane r stores `epoch` into flag `slot` of destination r (release: everything this stream wrote
// before the launch is visible to a peer that observes the flag)
__global__ void k_exchange_signal(PeerFlags dst, unsigned int mask, unsigned int epoch) {
    const int r = threadIdx.x;
    if (r < 32 && ((mask >> r) & 1u) && dst.p[r]) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned int *>(dst.p[r]) = epoch;
        __threadfence_system();
    }
}

// lane r spins until flags[r] >= epoch (acquire).  A peer that never arrives must not hang the
// GPU: after ~4 s (clock64 at ~2 GHz) the wait gives up and reports through *err.
__global__ void k_exchange_wait(const unsigned int *flags, unsigned int mask, unsigned int epoch,
                                int *err) {
    const int r = threadIdx.x;
    if (r < 32 && ((mask >> r) & 1u)) {
        const volatile unsigned int *f = flags + r;
        const long long t0 = clock64();
        while ((int)(*f - epoch) < 0) {
            __nanosleep(200);
            if (clock64() - t0 > 8000000000LL) {
                *err = 1 + r;
                break;
            }
        }
        __threadfence_system();
    }
}

// coalesced 16-byte copy of a finished block into a peer's slot (NVLink stores)
__global__ void __launch_bounds__(256) k_exchange_push(int4 *__restrict__ dst,
                                                       const int4 *__restrict__ src, int64_t n16) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n16;
         k += (int64_t)gridDim.x * blockDim.x)
        dst[k] = src[k];
}

// the same with the length on the device: n16 = ceil(min(*count * unit, max_bytes) / 16)
__global__ void __launch_bounds__(256) k_exchange_push_n(int4 *__restrict__ dst,
                                                         const int4 *__restrict__ src,
                                                         const long long *__restrict__ count,
                                                         long long unit, long long max_bytes) {
    long long bytes = *count * unit;
    if (bytes > max_bytes) bytes = max_bytes;
    const int64_t n16 = (bytes + 15) / 16;
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n16;
         k += (int64_t)gridDim.x * blockDim.x)
        dst[k] = src[k];
}

}  // namespace

extern "C" int mrmp_exchange_create(mrmp_ctx *c, int world, int rank, int n_chan,
                                    const int64_t *slot_bytes, mrmp_exchange **out,
                                    unsigned char *handle_out) {
    NEED(c && out && slot_bytes, "bad arguments");
    *out = nullptr;
    NEED(world >= 1 && world <= 32 && rank >= 0 && rank < world, "world / rank out of range");
    NEED(n_chan >= 1 && n_chan <= 8, "n_chan out of range");
    CU(cudaSetDevice(c->device));
    mrmp_exchange *x = new (std::nothrow) mrmp_exchange();
    if (!x) return fail(MRMP_ERR_NOMEM, "out of host memory");
    x->ctx = c;
    x->world = world;
    x->rank = rank;
    x->n_chan = n_chan;
    int64_t off = 0;
    for (int ch = 0; ch < n_chan; ++ch) {
        if (slot_bytes[ch] < 0) {
            delete x;
            return fail(MRMP_ERR_INVALID, "negative slot size");
        }
        const int64_t sb = (int64_t)align16((size_t)slot_bytes[ch]);
        x->slot_bytes.push_back(sb);
        x->chan_off.push_back(off);
        off += 2 * (int64_t)world * sb;
    }
    x->flags_off = off;
    x->total = off + (int64_t)align16(sizeof(unsigned int) * 32 * (size_t)n_chan);
    x->epoch.assign(n_chan, 1u);
    x->peer.assign(world, nullptr);
    x->opened.assign(world, false);
    cudaError_t e = cudaMalloc((void **)&x->buf, (size_t)x->total);
    if (e == cudaSuccess) e = cudaMemset(x->buf, 0, (size_t)x->total);
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&x->err_h, 64, cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void **)&x->err_d, x->err_h, 0);
    if (e == cudaSuccess) {
        *x->err_h = 0;
        if (handle_out) {
            cudaIpcMemHandle_t h;
            e = cudaIpcGetMemHandle(&h, x->buf);
            if (e == cudaSuccess) memcpy(handle_out, &h, sizeof(h));
        }
    }
    if (e != cudaSuccess) {
        cudaFree(x->buf);
        if (x->err_h) cudaFreeHost(x->err_h);
        delete x;
        return fail(MRMP_ERR_CUDA, "exchange alloc / IPC handle: %s", cudaGetErrorString(e));
    }
    x->peer[rank] = x->buf;
    *out = x;
    return MRMP_OK;
}

extern "C" int mrmp_exchange_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

extern "C" void mrmp_exchange_destroy(mrmp_exchange *x) {
    if (!x) return;
    cudaSetDevice(x->ctx->device);
    cudaStreamSynchronize(x->ctx->s_k);
    for (int r = 0; r < x->world; ++r)
        if (x->opened[r] && x->peer[r]) cudaIpcCloseMemHandle(x->peer[r]);
    cudaFree(x->buf);
    if (x->err_h) cudaFreeHost(x->err_h);
    delete x;
}

/* handles[world][mrmp_exchange_handle_bytes()], gathered from all ranks (own entry ignored) */
extern "C" int mrmp_exchange_connect(mrmp_exchange *x, const unsigned char *handles) {
    NEED(x && handles, "bad arguments");
    CU(cudaSetDevice(x->ctx->device));
    for (int r = 0; r < x->world; ++r) {
        if (r == x->rank || x->peer[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * sizeof(h), sizeof(h));
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(MRMP_ERR_UNSUPPORTED, "cudaIpcOpenMemHandle(rank %d): %s", r,
                        cudaGetErrorString(e));
        }
        x->peer[r] = (unsigned char *)p;
        x->opened[r] = true;
    }
    x->connected = true;
    return MRMP_OK;
}

/* same process (emulated ranks on one device, or several devices of one process with peer
 * access enabled by the caller): bufs[world] = mrmp_exchange_local_buffer of every rank */
extern "C" int mrmp_exchange_connect_ptrs(mrmp_exchange *x, void *const *bufs) {
    NEED(x && bufs, "bad arguments");
    for (int r = 0; r < x->world; ++r) {
        NEED(bufs[r], "NULL peer buffer");
        if (r != x->rank) x->peer[r] = (unsigned char *)bufs[r];
    }
    x->connected = true;
    return MRMP_OK;
}

extern "C" void *mrmp_exchange_local_buffer(mrmp_exchange *x) { return x ? x->buf : nullptr; }

static inline unsigned char *slot_addr(const mrmp_exchange *x, int ch, int dst, int src,
                                       unsigned int epoch) {
    return x->peer[dst] + x->chan_off[ch] +
           ((int64_t)(epoch & 1u) * x->world + src) * x->slot_bytes[ch];
}

/* where THIS rank writes its slot of channel ch (current epoch) in the buffer of rank dst */
extern "C" void *mrmp_exchange_ptr(mrmp_exchange *x, int ch, int dst) {
    if (!x || ch < 0 || ch >= x->n_chan || dst < 0 || dst >= x->world || !x->peer[dst])
        return nullptr;
    return slot_addr(x, ch, dst, x->rank, x->epoch[ch]);
}

/* base of the local [world][slot] block of channel ch for the current epoch (slot s = what rank
 * s wrote); valid to read after mrmp_exchange_wait */
extern "C" void *mrmp_exchange_slots(mrmp_exchange *x, int ch) {
    if (!x || ch < 0 || ch >= x->n_chan) return nullptr;
    return slot_addr(x, ch, x->rank, 0, x->epoch[ch]);
}

extern "C" int64_t mrmp_exchange_slot_bytes(mrmp_exchange *x, int ch) {
    return (x && ch >= 0 && ch < x->n_chan) ? x->slot_bytes[ch] : -1;
}

/* after the kernels that filled this rank's slots on `stream`: publish the current epoch of
 * channel ch to the ranks in dst_mask (bit r = rank r; include the own rank if it waits too) */
extern "C" int mrmp_exchange_signal(mrmp_exchange *x, int ch, unsigned int dst_mask, void *stream) {
    NEED(x && ch >= 0 && ch < x->n_chan, "bad exchange / channel");
    NEED(x->connected || x->world == 1, "exchange not connected");
    CU(cudaSetDevice(x->ctx->device));
    PeerFlags pf;
    for (int r = 0; r < 32; ++r)
        pf.p[r] = (r < x->world && x->peer[r])
                      ? (unsigned int *)(x->peer[r] + x->flags_off) + 32 * ch + x->rank
                      : nullptr;
    k_exchange_signal<<<1, 32, 0, (cudaStream_t)stream>>>(pf, dst_mask, x->epoch[ch]);
    count_launch();
    CU(cudaGetLastError());
    return MRMP_OK;
}

/* make `stream` wait until the ranks in src_mask have published the current epoch of ch */
extern "C" int mrmp_exchange_wait(mrmp_exchange *x, int ch, unsigned int src_mask, void *stream) {
    NEED(x && ch >= 0 && ch < x->n_chan, "bad exchange / channel");
    CU(cudaSetDevice(x->ctx->device));
    k_exchange_wait<<<1, 32, 0, (cudaStream_t)stream>>>(
        (const unsigned int *)(x->buf + x->flags_off) + 32 * ch, src_mask, x->epoch[ch], x->err_d);
    count_launch();
    CU(cudaGetLastError());
    return MRMP_OK;
}

/* next epoch of channel ch (every rank calls it once per exchange round, in the same order) */
extern "C" int mrmp_exchange_advance(mrmp_exchange *x, int ch) {
    NEED(x && ch >= 0 && ch < x->n_chan, "bad exchange / channel");
    x->epoch[ch] += 1;
    return MRMP_OK;
}

/* 0, or 1 + rank of a peer a wait gave up on (after the stream has been synchronised) */
extern "C" int mrmp_exchange_error(mrmp_exchange *x) { return x ? *x->err_h : -1; }

/* copy `bytes` (a multiple of 16, at most the slot size) of device memory of this rank into its
 * slot of channel ch in rank dst's buffer; follow with mrmp_exchange_signal */
extern "C" int mrmp_exchange_push(mrmp_exchange *x, int ch, int dst, const void *src_dev,
                                  int64_t bytes, void *stream) {
    NEED(x && ch >= 0 && ch < x->n_chan && dst >= 0 && dst < x->world && src_dev, "bad arguments");
    NEED(bytes >= 0 && bytes % 16 == 0 && bytes <= x->slot_bytes[ch] && aligned16(src_dev),
         "bytes must be a multiple of 16 within the slot, src 16-byte aligned");
    void *p = mrmp_exchange_ptr(x, ch, dst);
    NEED(p, "exchange is not connected to the destination rank");
    if (bytes == 0) return MRMP_OK;
    CU(cudaSetDevice(x->ctx->device));
    const int64_t n16 = bytes / 16;
    int64_t grid = (n16 + 255) / 256;
    const int64_t cap = (int64_t)x->ctx->sm_count * 4;
    if (grid > cap) grid = cap;
    k_exchange_push<<<(int)grid, 256, 0, (cudaStream_t)stream>>>((int4 *)p, (const int4 *)src_dev, n16);
    count_launch();
    CU(cudaGetLastError());
    return MRMP_OK;
}

/* as mrmp_exchange_push with the length taken from device memory at run time:
 * min(*count_dev * unit_bytes, max_bytes) bytes, rounded up to 16 (max_bytes a multiple of 16
 * within the slot; the source buffer must be readable up to that rounding) */
extern "C" int mrmp_exchange_push_counted(mrmp_exchange *x, int ch, int dst, const void *src_dev,
                                          const long long *count_dev, int64_t unit_bytes,
                                          int64_t max_bytes, void *stream) {
    NEED(x && ch >= 0 && ch < x->n_chan && dst >= 0 && dst < x->world && src_dev && count_dev,
         "bad arguments");
    NEED(unit_bytes > 0 && max_bytes >= 0 && max_bytes % 16 == 0 && max_bytes <= x->slot_bytes[ch] &&
             aligned16(src_dev),
         "max_bytes must be a multiple of 16 within the slot, src 16-byte aligned");
    void *p = mrmp_exchange_ptr(x, ch, dst);
    NEED(p, "exchange is not connected to the destination rank");
    if (max_bytes == 0) return MRMP_OK;
    CU(cudaSetDevice(x->ctx->device));
    int64_t grid = (max_bytes / 16 + 255) / 256;
    const int64_t cap = (int64_t)x->ctx->sm_count * 4;
    if (grid > cap) grid = cap;
    k_exchange_push_n<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(
        (int4 *)p, (const int4 *)src_dev, count_dev, (long long)unit_bytes, (long long)max_bytes);
    count_launch();
    CU(cudaGetLastError());
    return MRMP_OK;
}
