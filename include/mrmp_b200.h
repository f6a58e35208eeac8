/*
 * mrmp_b200.h -- C ABI of the B200-native connect / collide / roadmap hot path.
 *
 * Drop-in boundary for Kei18/sssp (MRMP.jl).  The reference has no FFI: its seam is
 * the closure API  gen_connect / gen_collide / gen_check_goal  (src/MRMP.jl:98).  Each
 * entry point below names the reference closure method (file:line under
 * /root/reference) it replaces; julia/MRMPB200.jl and INTEGRATION.md show the `ccall`
 * glue that rebuilds the closures on top of this library.
 *
 * Conventions
 *  - plain C, no torch / C++ types; every pointer is a HOST pointer to caller-owned,
 *    caller-sized memory unless the function name ends in _dev (then: device pointers
 *    on the ctx's device plus a cudaStream_t passed as void*).
 *  - states are the reference's isbits layout: StatePoint2D = double[2] (point2d.jl:3-6),
 *    StatePoint3D = double[3] (point3d.jl:3-7), StateArm22 = double[2] (arm22.jl:3-6);
 *    obstacles are CircleObstacle2D/3D = double[3]/double[4] (obstacle.jl:7-19);
 *    verdicts are Julia Bool = uint8_t 0/1.
 *  - agent indices and vertex ids are 0-BASED here (the Julia glue subtracts 1).
 *  - every function returns MRMP_OK (0) or a negative error; the message is available
 *    from mrmp_last_error() (thread-local).  Nothing throws or aborts.
 *  - a ctx is confined to one host thread at a time (one CUDA stream set); distinct
 *    ctxs may be used concurrently (scripts/eval.jl:152 creates closures per thread).
 *  - there is NO CPU fallback: without a CUDA device every compute call fails with
 *    MRMP_ERR_CUDA.
 */
#ifndef MRMP_B200_H
#define MRMP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRMP_OK 0
#define MRMP_ERR_INVALID (-1)     /* bad argument */
#define MRMP_ERR_CUDA (-2)        /* CUDA runtime / no device */
#define MRMP_ERR_CAPACITY (-3)    /* caller-provided output too small (see *_out sizes) */
#define MRMP_ERR_NOMEM (-4)
#define MRMP_ERR_UNSUPPORTED (-5)

#define MRMP_MODEL_POINT2D 0 /* src/models/point2d.jl */
#define MRMP_MODEL_POINT3D 1 /* src/models/point3d.jl */
#define MRMP_MODEL_ARM22 2   /* src/models/arm22.jl   */
#define MRMP_MODEL_LINE2D 3    /* src/models/line2d.jl:    state (x, y, theta) */
#define MRMP_MODEL_SNAKE2D 4   /* src/models/snake2d.jl:   state (x, y, theta1..4) */
#define MRMP_MODEL_ARM33 5     /* src/models/arm33.jl:     state (theta1, phi1, ..., theta3, phi3) */
#define MRMP_MODEL_CAPSULE3D 6 /* src/models/capsule3d.jl: state (x, y, z, roll, pitch, yaw) */

typedef struct mrmp_ctx mrmp_ctx;           /* one problem instance (obstacles, radii) */
typedef struct mrmp_roadmaps mrmp_roadmaps; /* device-resident per-agent roadmaps */

/* ---- library ----------------------------------------------------------------- */
const char *mrmp_last_error(void);
const char *mrmp_version(void);
/* number of kernels this library has launched in the calling process (diagnostics,
 * bench.py's gpu_launches) */
int64_t mrmp_kernel_launch_count(void);
int mrmp_device_count(int *n_out);

/* pinned host memory for the batch entry points (plain pageable pointers also work) */
int mrmp_host_alloc(void **ptr_out, int64_t bytes);
int mrmp_host_free(void *ptr);

/* ---- context: replaces the closure constructors ------------------------------- */
/* gen_connect(q, obstacles, ins_params...; step_dist, max_dist, safety_dist) and
 * gen_collide(q, ins_params...; step_dist, safety_dist):
 *   point2d.jl:42-69, point3d.jl:47-80, point.jl:5-36, arm22.jl:29-37, arm22.jl:90-96.
 * and the same constructors of line2d.jl:22-28,72-77, snake2d.jl:61-68,127-132,
 * arm33.jl:42-50,116-122, capsule3d.jl:61-68,107-112.
 * rads[n_agents]; base_pos = positions, [n_agents*2] for ARM22 (arm22.jl:32), [n_agents*3]
 * for ARM33 (arm33.jl:45), else NULL; aux = axises[n_agents] for CAPSULE3D (capsule3d.jl:65),
 * else NULL; obstacles_aos[n_obs*(w+1)] = x,y,(z),r with w the workspace dimension;
 * step_dist / safety_dist are used by the discretised models (STEP_DIST / SAFETY_DIST_LINE,
 * src/MRMP.jl:65-66; snake2d.jl:3 and capsule3d.jl:3 default to step 0.05); max_dist NaN ==
 * nothing; device = CUDA ordinal.  State layouts: the reference's isbits structs. */
int mrmp_ctx_create(mrmp_ctx **out, int model, int n_agents, const double *rads,
                    const double *base_pos, const double *aux, int n_obs,
                    const double *obstacles_aos, double step_dist, double safety_dist,
                    double max_dist, int device);
/* Objects created on a ctx (mrmp_roadmaps, mrmp_sssp, mrmp_exchange) use it in their own destroy
 * call: destroy them first.  (A binding whose finalisers run in no particular order -- Python's
 * cyclic GC, Julia's finalizers -- must defer this call until its dependents are gone, as
 * sssp_b200/closures.py: Context.close does.) */
void mrmp_ctx_destroy(mrmp_ctx *ctx);
int mrmp_ctx_state_dim(const mrmp_ctx *ctx);
int mrmp_ctx_sync(mrmp_ctx *ctx);
/* run the ctx's kernels and small copies on a caller-owned cudaStream_t (NULL restores the
 * ctx's own, non-blocking stream); lets a host framework order and time the work on its stream.
 * A framework whose stream is the default stream (handle 0) passes cudaStreamLegacy
 * ((cudaStream_t)0x1): otherwise the ctx's launches (own stream) and the `stream` arguments of
 * the *_dev / *_async entry points (stream 0) would be two unordered streams. */
int mrmp_ctx_set_stream(mrmp_ctx *ctx, void *stream);
/* diagnostics since the last reset: out4[0] = point-model pairs that survived the cross
 * kernel's broad phase (ran the exact reference arithmetic); out4[1] = arm22 link pairs that ran
 * the exact test; out4[2] = point-model pairs whose column tile survived the box test (ran the
 * broad-phase bound); out4[3] = point-model pairs evaluated by the verdict filter (out4[0] of them
 * undecided -> reference arithmetic) */
int mrmp_ctx_stats(mrmp_ctx *ctx, int64_t *out4, int reset);
/* the same four, then out8[4] = arm22 link-pair midpoint bounds evaluated (4 per posture pair),
 * out8[5..7] reserved */
int mrmp_ctx_stats8(mrmp_ctx *ctx, int64_t *out8, int reset);

/* ---- connect -------------------------------------------------------------------- */
/* connect(q, i): point2d.jl:49-57, point3d.jl:54-64, arm22.jl:40-54 */
int mrmp_connect_point(mrmp_ctx *ctx, int agent, const double *q, uint8_t *out);
int mrmp_connect_points(mrmp_ctx *ctx, int64_t n, const int32_t *agent, const double *q,
                        uint8_t *out);
/* connect(q_from, q_to, i): point2d.jl:59-66, point3d.jl:66-77, arm22.jl:56-86 */
int mrmp_connect_edge(mrmp_ctx *ctx, int agent, const double *q_from, const double *q_to,
                      uint8_t *out);
int mrmp_connect_edges(mrmp_ctx *ctx, int64_t n, const int32_t *agent, const double *q_from,
                       const double *q_to, uint8_t *out);

/* ---- collide -------------------------------------------------------------------- */
/* collide(qi_from, qi_to, qj_from, qj_to, i, j): point.jl:9-12, arm22.jl:100-146 */
int mrmp_collide_pair(mrmp_ctx *ctx, int i, int j, const double *qi_from, const double *qi_to,
                      const double *qj_from, const double *qj_to, uint8_t *out);
int mrmp_collide_pairs(mrmp_ctx *ctx, int64_t n, const int32_t *i, const int32_t *j,
                       const double *qi_from, const double *qi_to, const double *qj_from,
                       const double *qj_to, uint8_t *out);
/* collide(q_i, q_j, i, j): point.jl:14-16, arm22.jl:148-162 */
int mrmp_collide_static_pairs(mrmp_ctx *ctx, int64_t n, const int32_t *i, const int32_t *j,
                              const double *qi, const double *qj, uint8_t *out);
/* collide(Q_from, Q_to) for n_cfg configuration pairs (each N states): point.jl:18-25,
 * arm22.jl:164-171; callers utils.jl:360 (validate), sssp.jl:207.  first_pair_out (may be
 * NULL) receives i*N+j of the first colliding pair in the reference's (i asc, j asc)
 * loop order, or -1. */
int mrmp_collide_configs(mrmp_ctx *ctx, int64_t n_cfg, const double *Q_from, const double *Q_to,
                         uint8_t *out, int64_t *first_pair_out);
/* collide(Q, q_to, i) for n_cand candidate targets of agent i against the static other
 * agents of Q (N states): point.jl:27-33, arm22.jl:173-209; caller sssp.jl:206. */
int mrmp_collide_config_moves(mrmp_ctx *ctx, int agent_i, const double *Q, int64_t n_cand,
                              const double *q_to, uint8_t *out);

/* device-pointer forms (inputs already resident in HBM; asynchronous on `stream`) */
int mrmp_connect_points_dev(mrmp_ctx *ctx, int64_t n, const int32_t *agent, const double *q,
                            uint8_t *out, void *stream);
int mrmp_connect_edges_dev(mrmp_ctx *ctx, int64_t n, const int32_t *agent, const double *q_from,
                           const double *q_to, uint8_t *out, void *stream);
int mrmp_collide_pairs_dev(mrmp_ctx *ctx, int64_t n, const int32_t *i, const int32_t *j,
                           const double *qi_from, const double *qi_to, const double *qj_from,
                           const double *qj_to, uint8_t *out, void *stream);
int mrmp_collide_static_pairs_dev(mrmp_ctx *ctx, int64_t n, const int32_t *i, const int32_t *j,
                                  const double *qi, const double *qj, uint8_t *out,
                                  void *stream);
int mrmp_collide_configs_dev(mrmp_ctx *ctx, int64_t n_cfg, const double *Q_from,
                             const double *Q_to, uint8_t *out, int64_t *first_pair_out,
                             void *stream);
int mrmp_collide_config_moves_dev(mrmp_ctx *ctx, int agent_i, const double *Q, int64_t n_cand,
                                  const double *q_to, uint8_t *out, void *stream);

/* ---- roadmaps (new batched API; the reference builds them one check at a time) --- */
/* A roadmap set holds, on the device, the node tables and CSR adjacency of `n_agents`
 * consecutive agents starting at `agent0` (the shard of one GPU), nv_cap vertices each. */
int mrmp_roadmaps_create(mrmp_ctx *ctx, int agent0, int n_agents, int nv_cap,
                         mrmp_roadmaps **out);
void mrmp_roadmaps_destroy(mrmp_roadmaps *rm);
/* upload node tables: nodes[n_agents][nv][d] (Vector{Node}.q gathered flat by the glue) */
int mrmp_roadmaps_set_nodes(mrmp_roadmaps *rm, int nv, const double *nodes);
/* PRM! sampling loop (prm.jl:193-199) on the device: vertex 0 = q_init, 1 = q_goal
 * (prm.jl:232), then the first nv-2 free samples of the counter-based sampler
 * (Philox4x32-10, stream (seed, agent)) in try order.  tries_out[n_agents] may be NULL. */
int mrmp_roadmaps_sample(mrmp_roadmaps *rm, int nv, uint64_t seed, const double *q_init,
                         const double *q_goal, int64_t *tries_out);
/* PRM! neighbour pass (prm.jl:202-210) for every agent of the set: directed edges j->k,
 * j != k, (rad is NaN || dist(q_j,q_k) <= rad) && connect(q_j,q_k,i); rows ascending in k.
 * rad[n_agents].  Also covers the all-pairs pass of sssp.jl:388-395 (rad = epsilon). */
int mrmp_roadmaps_build(mrmp_roadmaps *rm, const double *rad);
/* Asynchronous pipeline (no host round trip between the kernels of one step): build_async
 * launches the same kernels and returns; the edge count stays on the device.  Until
 * mrmp_roadmaps_finish() the set may only be passed to the calls that take the count from the
 * device: mrmp_roadmaps_edge_conflicts_async (as the row set), mrmp_roadmaps_push_csr_shard,
 * mrmp_roadmaps_push_table_rows.  finish() waits for the ctx's stream, publishes nnz and checks
 * that the CSR buffer, a pushed shard slot and a table's row capacity were large enough
 * (MRMP_ERR_CAPACITY otherwise: those results are truncated, the buffers have been grown, build
 * again).  Large point roadmaps (uniform-grid path) build synchronously. */
int mrmp_roadmaps_build_async(mrmp_roadmaps *rm, const double *rad);
int mrmp_roadmaps_finish(mrmp_roadmaps *rm, int64_t *nnz_out);
/* diagnostics: 1 when the last build searched neighbours through the uniform-grid spatial hash
 * (point models, nv >= 2048, finite rad), 0 when it tested all ordered pairs */
int mrmp_roadmaps_used_grid(mrmp_roadmaps *rm);
int mrmp_roadmaps_nnz(mrmp_roadmaps *rm, int64_t *nnz_total_out);
int mrmp_roadmaps_get_nodes(mrmp_roadmaps *rm, double *nodes_out);
/* packed CSR over all rows of all agents: row_ptr[n_agents*nv+1] (offsets into col_idx),
 * col_idx[nnz] = target vertex id within the same agent.  MRMP_ERR_CAPACITY if
 * col_cap < nnz (nnz_out is still written so the caller can retry). */
int mrmp_roadmaps_get_csr(mrmp_roadmaps *rm, int32_t *row_ptr, int32_t *col_idx,
                          int64_t col_cap, int64_t *nnz_out);
/* adopt an adjacency assembled elsewhere (all-gathered shards): device arrays, packed CSR */
int mrmp_roadmaps_set_csr_dev(mrmp_roadmaps *rm, const int32_t *row_ptr_dev,
                              const int32_t *col_idx_dev, int64_t nnz, void *stream);
/* Multi-GPU builds: one collective instead of a size exchange.  Every rank packs the CSR of its
 * agent block into a fixed-size slot  int32 cnt[cap_agents*nv] | int32 col[cap_nnz]  (rows beyond
 * its agents count 0), all-gathers the slots (NCCL), and adopts the gathered slots as the CSR
 * of the full set: row counts -> scan -> columns copied row by row, all on `stream`.
 * pack returns MRMP_ERR_CAPACITY when the shard has more than cap_nnz edges. */
int mrmp_roadmaps_pack_csr_shard_dev(mrmp_roadmaps *rm, int cap_agents, int64_t cap_nnz,
                                     int32_t *slot_dev, void *stream);
int mrmp_roadmaps_set_csr_shards_dev(mrmp_roadmaps *rm, int world, int cap_agents, int64_t cap_nnz,
                                     const int32_t *slots_dev, void *stream, int64_t *nnz_out);
/* Asynchronous read-back of a built roadmap set (what PRMs returns, prm.jl:251-276) into PINNED
 * host buffers on the ctx's copy stream: returns at once, so the caller can launch the next
 * batch (e.g. the conflict table) while the roadmaps travel.  mrmp_roadmaps_read_wait() blocks
 * until they have arrived.  The set must not be rebuilt in between.
 * Behind mrmp_roadmaps_build_async (call it before the table is launched): the copies wait for the
 * build's kernels only and overlap the table kernel; the edge count is not known yet, so all
 * col_cap entries are copied, *nnz_out is -1, and mrmp_roadmaps_finish reports the count (and
 * MRMP_ERR_CAPACITY if it exceeded col_cap). */
int mrmp_roadmaps_read_async(mrmp_roadmaps *rm, double *nodes_out, int32_t *row_ptr,
                             int32_t *col_idx, int64_t col_cap, int64_t *nnz_out);
int mrmp_roadmaps_read_wait(mrmp_roadmaps *rm);
/* Pin (page-lock and map) / unpin a host buffer of the caller: host code without a CUDA binding of
 * its own (the reference's Julia: `Vector`s passed to ccall) gets the asynchronous read-back and
 * the in-place output of OR-reduced conflict tables for its ordinary arrays. */
int mrmp_host_register(void *ptr, size_t bytes);
int mrmp_host_unregister(void *ptr);

/* ---- peer-memory exchange between the GPUs of one box (one process per GPU) -------------------
 * What shards is the agent axis (prm.jl:265-275: roadmaps are built per agent; pp.jl:179-189: one
 * agent's candidate edges against the others' paths).  What must move is (1) the CSR of each
 * rank's agent block, so that every GPU holds all roadmaps, and (2) the conflict-table rows a rank
 * computed.  Both are written by the producing kernels STRAIGHT INTO THE PEERS' MEMORY over NVLink
 * (peer mappings from CUDA IPC handles), followed by one flag per peer; consumers wait on their
 * local flags.  No collective call and nothing on the host between producer and consumer.
 *
 * An exchange has n_chan channels; per channel every rank owns one slot of slot_bytes[ch] (double
 * buffered by epoch) in every rank's buffer.  Per round and channel, in stream order:
 *   producer: write own slot at mrmp_exchange_ptr(x, ch, dst) for each destination,
 *             mrmp_exchange_signal(x, ch, dst_mask, stream)
 *   consumer: mrmp_exchange_wait(x, ch, src_mask, stream), read mrmp_exchange_slots(x, ch)
 *   both:     mrmp_exchange_advance(x, ch) once the round's launches are enqueued
 * handle_out (mrmp_exchange_handle_bytes() bytes, may be NULL) is this rank's IPC handle; the host
 * framework all-gathers the handles (sssp_b200/peer.py uses torch.distributed) and passes the
 * [world][handle_bytes] array to mrmp_exchange_connect.  MRMP_ERR_UNSUPPORTED when the peer
 * mapping cannot be opened (the caller falls back to a collective).  A wait gives up after a few
 * seconds instead of hanging the GPU; mrmp_exchange_error() then names the missing rank. */
typedef struct mrmp_exchange mrmp_exchange;
int mrmp_exchange_handle_bytes(void);
int mrmp_exchange_create(mrmp_ctx *ctx, int world, int rank, int n_chan, const int64_t *slot_bytes,
                         mrmp_exchange **out, unsigned char *handle_out);
void mrmp_exchange_destroy(mrmp_exchange *x);
int mrmp_exchange_connect(mrmp_exchange *x, const unsigned char *handles);
/* one process driving several ranks (tests: emulated ranks on one device) */
int mrmp_exchange_connect_ptrs(mrmp_exchange *x, void *const *bufs);
void *mrmp_exchange_local_buffer(mrmp_exchange *x);
void *mrmp_exchange_ptr(mrmp_exchange *x, int ch, int dst);
void *mrmp_exchange_slots(mrmp_exchange *x, int ch);
int64_t mrmp_exchange_slot_bytes(mrmp_exchange *x, int ch);
/* copy kernel: `bytes` (multiple of 16) of this rank's device memory -> its slot in rank dst */
int mrmp_exchange_push(mrmp_exchange *x, int ch, int dst, const void *src_dev, int64_t bytes,
                       void *stream);
/* the same with the length on the device: min(*count_dev * unit_bytes, max_bytes), rounded up to 16 */
int mrmp_exchange_push_counted(mrmp_exchange *x, int ch, int dst, const void *src_dev,
                               const long long *count_dev, int64_t unit_bytes, int64_t max_bytes,
                               void *stream);
int mrmp_exchange_signal(mrmp_exchange *x, int ch, unsigned int dst_mask, void *stream);
int mrmp_exchange_wait(mrmp_exchange *x, int ch, unsigned int src_mask, void *stream);
int mrmp_exchange_advance(mrmp_exchange *x, int ch);
int mrmp_exchange_error(mrmp_exchange *x);
/* Multi-GPU roadmap build without a collective: pack the CSR of this rank's agent block (slot
 * layout of mrmp_roadmaps_pack_csr_shard_dev; 4*(cap_agents*nv + cap_nnz) must equal the channel's
 * slot size and be a multiple of 16) into the slot of this rank in EVERY rank's buffer and
 * signal; adopt = wait for all ranks, then stitch the slots into the CSR of the full set. */
int mrmp_roadmaps_push_csr_shard(mrmp_roadmaps *rm, mrmp_exchange *x, int ch, int world,
                                 int cap_agents, int64_t cap_nnz, void *stream);
int mrmp_roadmaps_adopt_csr_shards(mrmp_roadmaps *rm, mrmp_exchange *x, int ch, int world,
                                   int cap_agents, int64_t cap_nnz, void *stream, int64_t *nnz_out);
/* no wait: the stitched set stays pending (see mrmp_roadmaps_build_async) until
 * mrmp_roadmaps_finish (which waits on `stream` as well as on the ctx's stream) */
int mrmp_roadmaps_adopt_csr_shards_async(mrmp_roadmaps *rm, mrmp_exchange *x, int ch, int world,
                                         int cap_agents, int64_t cap_nnz, void *stream);
/* conflict-table rows of this set (words_per_row words each, device memory of this rank) -> this
 * rank's slot of channel ch in rank dst's buffer (copy kernel over NVLink); the row count is the
 * set's edge count, read on the device while a build is pending; max_rows = room of the source */
int mrmp_roadmaps_push_table_rows(mrmp_roadmaps *rm, mrmp_exchange *x, int ch, int dst,
                                  const uint32_t *bits_dev, int words_per_row, int64_t max_rows,
                                  void *stream);

/* raw device pointers (for NCCL all-gather of shards / zero-copy consumers) */
int mrmp_roadmaps_device_ptrs(mrmp_roadmaps *rm, int *nv_out, double **nodes_dev,
                              int32_t **row_ptr_dev, int32_t **col_idx_dev, int64_t *nnz_out);
/* adopt externally filled device arrays (after an all-gather): copies device->device */
int mrmp_roadmaps_set_nodes_dev(mrmp_roadmaps *rm, int nv, const double *nodes_dev,
                                void *stream);

/* use caller-owned device storage nodes_dev[n_agents][nv][d] as the node table (e.g. a slice
 * of an NCCL all-gather buffer, so sampling writes straight into the send buffer and the
 * gathered buffer is consumed in place); NULL switches back to the set's own storage */
int mrmp_roadmaps_bind_nodes(mrmp_roadmaps *rm, int nv, double *nodes_dev);

/* one-shot host-buffer forms */
/* PRM! neighbour pass on caller-provided nodes[n_agents][nv][d] (prm.jl:202-210) */
int mrmp_prm_build(mrmp_ctx *ctx, int agent0, int n_agents, int nv, const double *rad,
                   const double *nodes, int32_t *row_ptr, int32_t *col_idx, int64_t col_cap,
                   int64_t *nnz_out);
/* PRMs(config_init, config_goal, connect, nv; rads) (prm.jl:251-276): sampling + neighbour
 * pass for agents agent0..agent0+n_agents-1; q_init/q_goal [n_agents][d] */
int mrmp_prms(mrmp_ctx *ctx, int agent0, int n_agents, int nv, const double *rad, uint64_t seed,
              const double *q_init, const double *q_goal, double *nodes_out, int32_t *row_ptr,
              int32_t *col_idx, int64_t col_cap, int64_t *nnz_out);

/* collide(6-arity) over roadmap edges by id: check k tests agent i[k]'s edge
 * (vi_from[k] -> vi_to[k]) against agent j[k]'s edge (vj_from[k] -> vj_to[k]); agent ids are
 * global, vertices are looked up in `rm` (which must hold both agents).  Callers:
 * pp.jl:179-189, cbs.jl:336-347,384-409, smoother.jl:86-103. */
int mrmp_collide_edges_indexed(mrmp_roadmaps *rm, int64_t n, const int32_t *i, const int32_t *j,
                               const int32_t *vi_from, const int32_t *vi_to,
                               const int32_t *vj_from, const int32_t *vj_to, uint8_t *out);
int mrmp_collide_edges_indexed_dev(mrmp_roadmaps *rm, int64_t n, const int32_t *i,
                                   const int32_t *j, const int32_t *vi_from,
                                   const int32_t *vi_to, const int32_t *vj_from,
                                   const int32_t *vj_to, uint8_t *out, void *stream);

/* ---- cross-product collide (new batched API): every row edge x every column edge --------- */
/* verdict(r, c) = row_agent[r] != col_agent[c] &&
 *                 collide(row_from[r], row_to[r], col_from[c], col_to[c], row_agent[r], col_agent[c])
 * (point.jl:9-12; row edge is the FIRST segment).  Batched form of PP's invalid()
 * (pp.jl:179-189), CBS's soft conflict count (cbs.jl:336-347), the TPG dependency scan
 * (smoother.jl:86-103) and SSSP's successor filter (sssp.jl:196-223).
 * group_size == 0: out_bits is the full bit matrix, words_per_row = ceil(n_cols/32), bit c%32
 *   of word c/32 of row r.
 * group_size  > 0: columns form consecutive groups of group_size (e.g. all agents' path edges
 *   of one time step); bit g of row r = OR over group g; words_per_row =
 *   ceil(ceil(n_cols/group_size)/32). */
int mrmp_collide_cross(mrmp_ctx *ctx, int64_t n_rows, const int32_t *row_agent,
                       const double *row_from, const double *row_to, int64_t n_cols,
                       const int32_t *col_agent, const double *col_from, const double *col_to,
                       int group_size, uint32_t *out_bits);
int mrmp_collide_cross_dev(mrmp_ctx *ctx, int64_t n_rows, const int32_t *row_agent,
                           const double *row_from, const double *row_to, int64_t n_cols,
                           const int32_t *col_agent, const double *col_from,
                           const double *col_to, int group_size, uint32_t *out_bits,
                           void *stream);
/* Reductions of the same cross product (out: int32 [n_rows][n_groups]), for the callers that
 * need more than a bit per pair:
 *  MRMP_CROSS_COUNT  out[r][g] = number of colliding columns of group g.  With MRMP_CROSS_FLAG_CBS
 *    the pair test is the one of CBS's soft conflict count g_func_i (cbs.jl:336-347) and of
 *    get_num_collisions (cbs.jl:449-457 via :391-401):
 *      collide(col_to, row_to, j, i) || collide(col_from, col_to, row_from, row_to, j, i)
 *    (vertex conflict OR edge conflict, the column's agent j first, as written at :345-346).
 *    Rows = candidate edges of agent i, columns = the other agents' path edges, group = time step.
 *  MRMP_CROSS_FIRST  out[r][g] = smallest column index (caller's order) of group g that collides
 *    with row r, -1 if none: the type-2 dependency scan of get_temporal_plan_graph
 *    (smoother.jl:78-106: per action_self and other agent j the FIRST later action_other that
 *    collides).  row_key / col_key (both NULL, or both given): only pairs with
 *    col_key[c] > row_key[r] are tested (`filter(a -> a.t > action_self.t, TPG[j])`, :88).
 * Groups: col_group[c] in [0, n_groups) if non-NULL, else c / group_size (group_size > 0,
 * n_groups >= ceil(n_cols / group_size)).  Same-agent pairs never count. */
#define MRMP_CROSS_COUNT 2
#define MRMP_CROSS_FIRST 3
#define MRMP_CROSS_FLAG_CBS 1
int mrmp_collide_cross_reduce(mrmp_ctx *ctx, int64_t n_rows, const int32_t *row_agent,
                              const double *row_from, const double *row_to,
                              const int32_t *row_key, int64_t n_cols, const int32_t *col_agent,
                              const double *col_from, const double *col_to,
                              const int32_t *col_group, const int32_t *col_key, int group_size,
                              int n_groups, int mode, int flags, int32_t *out);
int mrmp_collide_cross_reduce_dev(mrmp_ctx *ctx, int64_t n_rows, const int32_t *row_agent,
                                  const double *row_from, const double *row_to,
                                  const int32_t *row_key, int64_t n_cols,
                                  const int32_t *col_agent, const double *col_from,
                                  const double *col_to, const int32_t *col_group,
                                  const int32_t *col_key, int group_size, int n_groups, int mode,
                                  int flags, int32_t *out, void *stream);
/* rows = every edge of the (built) roadmap set in packed CSR order (row e <-> col_idx[e]);
 * columns = roadmap edges given by ids (col_agent, col_vfrom -> col_vto) into rm_cols (NULL:
 * the same set; a shard passes the all-gathered full set here).  out_bits
 * [nnz][words_per_row]; out_cap_words = capacity of out_bits in 32-bit words. */
int mrmp_roadmaps_edge_conflicts(mrmp_roadmaps *rm, mrmp_roadmaps *rm_cols, int64_t n_cols,
                                 const int32_t *col_agent,
                                 const int32_t *col_vfrom, const int32_t *col_vto,
                                 int group_size, uint32_t *out_bits, int64_t out_cap_words);
int mrmp_roadmaps_edge_conflicts_dev(mrmp_roadmaps *rm, mrmp_roadmaps *rm_cols, int64_t n_cols,
                                     const int32_t *col_agent, const int32_t *col_vfrom,
                                     const int32_t *col_vto, int group_size, uint32_t *out_bits,
                                     void *stream);
/* the same behind mrmp_roadmaps_build_async: out_bits has room for out_cap_rows rows */
int mrmp_roadmaps_edge_conflicts_async(mrmp_roadmaps *rm, mrmp_roadmaps *rm_cols, int64_t n_cols,
                                       const int32_t *col_agent, const int32_t *col_vfrom,
                                       const int32_t *col_vto, int group_size, uint32_t *out_bits,
                                       int64_t out_cap_rows, void *stream);

/* One step of the strong-scaled path (prm.jl:265-275 sharded by agent + pp.jl:179-189 sharded by
 * row) on this rank in ONE call: [advance both channels] -> mrmp_roadmaps_sample(rm_all) ->
 * mrmp_roadmaps_build_async(rm_shard) -> mrmp_roadmaps_push_csr_shard(channel 0) ->
 * mrmp_roadmaps_edge_conflicts_async(rm_shard rows x rm_all columns -> tab_local) ->
 * mrmp_roadmaps_push_table_rows(channel 1, rank table_dst) + flag -> [wait_table: wait for every
 * rank's rows]; mrmp_roadmaps_adopt_csr_shards_async(rm_all) runs on the ctx's second stream.
 * Everything is enqueued on the ctx's stream; finish rm_shard, then rm_all. */
int mrmp_roadmaps_strong_step(mrmp_roadmaps *rm_all, mrmp_roadmaps *rm_shard, mrmp_exchange *x,
                              int world, int advance, int nv, uint64_t seed, const double *q_init,
                              const double *q_goal, const double *rad_shard, int cap_agents,
                              int64_t cap_nnz, int64_t n_cols, const int32_t *col_agent,
                              const int32_t *col_vfrom, const int32_t *col_vto, int group_size,
                              uint32_t *tab_local, int words_per_row, int64_t tab_rows_cap,
                              int table_dst, int wait_table);

/* sampler: gen_uniform_sampling (point2d.jl:71-75, point3d.jl:82-86, arm22.jl:214-218)
 * restated on a counter-based generator; writes try t0..t0+n-1 of (seed, agent) */
int mrmp_sample_states(mrmp_ctx *ctx, int agent, uint64_t seed, uint64_t t0, int64_t n,
                       double *q_out);
/* first n_wanted free samples in try order; *n_tried_out = tries consumed */
int mrmp_sample_free(mrmp_ctx *ctx, int agent, uint64_t seed, int64_t n_wanted, double *q_out,
                     int64_t *n_tried_out);

/* ---- model helpers for the host-side search loops ----------------------------------------- */
/* dist(q1, q2) of the model, n pairs: point2d.jl:12-14, point3d.jl:13-15, arm22.jl:15-20 */
int mrmp_state_dist(mrmp_ctx *ctx, int64_t n, const double *q1, const double *q2, double *out);
/* get_mid_status(p, q), n pairs: point2d.jl:8-10, point3d.jl:9-11, arm22.jl:8-13 */
int mrmp_mid_status(mrmp_ctx *ctx, int64_t n, const double *p, const double *q, double *out);
/* nearest vertex: argmin_v dist(V[v], q), or dist(q, V[v]) when reversed != 0; first minimum
 * like Julia's argmin (sssp.jl:416-417, extend!).  V[n_v][d]. */
int mrmp_nearest(mrmp_ctx *ctx, int n_v, const double *V, const double *q, int reversed,
                 int32_t *idx_out, double *dist_out);

/* extend!(connect, q_rand, V, from_start, steering_depth) of RRT-Connect (sssp.jl:404-440), which
 * builds SSSP's initial roadmaps (:325-401), in ONE call instead of one per closure call: nearest
 * vertex of the tree V[n_v][d] (argmin dist(q, q_rand), or dist(q_rand, q) for the goal tree),
 * the connection test towards q_rand -- connect(q_near, q) from the start tree, connect(q) &&
 * connect(q, q_near) from the goal tree, with the epsilon gate of conn (:87-89; NaN = nothing) --
 * and the bisection (get_mid_status) of steering_depth rounds.  flag_out: 0 = :trapped (q_new_out
 * is then not to be pushed), 1 = :advanced, 2 = :reached; near_out (may be NULL) = index of q_near. */
int mrmp_rrt_extend(mrmp_ctx *ctx, int agent, int n_v, const double *V, const double *q_rand,
                    int from_start, int steering_depth, double epsilon, int32_t *flag_out,
                    int32_t *near_out, double *q_new_out);

/* ---- SSSP: device-resident growing roadmaps and the fused node expansion ------------------ */
/* The reference's SSSP (src/solvers/sssp.jl:49-232) calls, for every search node it pops,
 * expand! (:235-267, K x [steering :269-291 + nearest-distance filter :250 + conn against the
 * whole roadmap :252-262]) and then the successor filter collide(S.Q, p.q, i) (:196-223) --
 * thousands of closure calls.  mrmp_sssp_expand does all of it in one call on roadmaps that
 * stay on the device; the best-first search (queue, visited set, distance tables) stays on
 * the host, exactly as in the reference. */
typedef struct mrmp_sssp mrmp_sssp;
int mrmp_sssp_create(mrmp_ctx *ctx, int nv_cap, mrmp_sssp **out); /* tables grow on demand */
void mrmp_sssp_destroy(mrmp_sssp *s);
/* upload one agent's initial roadmap vertices (gen_RRT_connect_roadmap, sssp.jl:294-401;
 * vertex 0 = q_init, 1 = q_goal); nodes[nv][d] */
int mrmp_sssp_set_roadmap(mrmp_sssp *s, int agent, int nv, const double *nodes);
int mrmp_sssp_nv(mrmp_sssp *s, int agent, int *nv_out);
int mrmp_sssp_get_nodes(mrmp_sssp *s, int agent, double *nodes_out, int cap_vertices);
/* One node expansion for `agent` at its vertex v_from (sssp.jl:181-223):
 *  - samples: q_rand[n_samples][d] if non-NULL, else tries t0 .. t0+n_samples-1 of the
 *    counter-based sampler stream (seed, agent) (same stream as mrmp_sample_states);
 *  - steering with conn(q_from,q_to) = (epsilon is NaN || dist <= epsilon) && connect(...)
 *    (sssp.jl:87-89, 269-291), acceptance  min_v dist(v.q, q_new) > min_dist_thread  in sample
 *    order against the roadmap INCLUDING vertices accepted earlier in the same call (:250);
 *  - accepted vertices get ids nv_before, nv_before+1, ...; q_new_out[n_new][d];
 *    out_bits[r][words]: bit t set <=> edge (new r -> t), t < id_r   (sssp.jl:252-257)
 *    in_bits [r][words]: bit t set <=> edge (t -> new r), t <= id_r  (self loop included,
 *    sssp.jl:259-262); words_out = ceil((nv_before + n_samples) / 32); rows r >= n_new are 0.
 *    Appending ids in ascending bit order reproduces the reference's neighbour-list order.
 *  - successor filter: Q_ids[N] = vertex id of every agent in the popped search node
 *    (Q_ids[agent] == v_from), old_nbrs = v_from's neighbour list before this call.
 *    succ_ids[n_succ] = old_nbrs, then the accepted vertices that v_from now reaches
 *    (ascending), then v_from (`vcat(v.neighbors, v.id)`, :197);
 *    succ_collide[k] = collide(S.Q, p.q, i) (:206), or collide(S.Q, Q) (:207) when
 *    slow_collide != 0 (no_fast_collision_check).  Both arrays hold n_old_nbrs+n_samples+1.
 * MRMP_ERR_CAPACITY if bits_cap_words < n_samples * words_out (words_out is still written). */
int mrmp_sssp_expand(mrmp_sssp *s, int agent, int v_from, int n_samples, const double *q_rand,
                     uint64_t seed, uint64_t t0, int steering_depth, double epsilon,
                     double min_dist_thread, const int32_t *Q_ids, int n_old_nbrs,
                     const int32_t *old_nbrs, int slow_collide, int32_t *n_new_out,
                     double *q_new_out, int32_t *words_out, uint32_t *out_bits, uint32_t *in_bits,
                     int64_t bits_cap_words, int32_t *n_succ_out, int32_t *succ_ids,
                     uint8_t *succ_collide);

/* diagnostics: host time of the last mrmp_sssp_expand spent inside the library (total / waiting
 * for the kernel after the launch call returned), microseconds */
int mrmp_sssp_last_call_us(mrmp_sssp *s, double *total_us, double *wait_us);

/* ---- distance tables: get_distance_table(s) (src/solvers/utils.jl:146-236) ------------------ */
/* n_graphs roadmaps at once (SSSP after every expansion that added a vertex, sssp.jl:193;
 * PP / CBS for all agents up front, pp.jl:161, cbs.jl:197).  Graph g owns vertices
 * v_off[g] .. v_off[g+1]-1 of nodes[V][d]; row_ptr[V+1] / col_idx is the CSR of the neighbour
 * lists (col ids local to the graph); goal[g] is the local id of the goal vertex
 * (roadmap[2] in the reference).  tables_out[V]: table[u] = cost from the goal along the
 * neighbour lists with g = dist(v.q, u.q) + table[v], Inf where unreachable -- bit-identical to
 * the reference's Dijkstra (the minimum over paths does not depend on relaxation order). */
int mrmp_distance_tables(mrmp_ctx *ctx, int n_graphs, const int32_t *v_off, const double *nodes,
                         const int32_t *row_ptr, const int32_t *col_idx, const int32_t *goal,
                         double *tables_out);
/* the same with the edge cost of gen_g_func(stay_penalty = p) (solvers/utils.jl:29-39), which
 * PP / CBS pass as g_func (pp.jl:54,161): dist(v.q, u.q) + (v.q == u.q ? p : 0). */
int mrmp_distance_tables_g(mrmp_ctx *ctx, int n_graphs, const int32_t *v_off, const double *nodes,
                           const int32_t *row_ptr, const int32_t *col_idx, const int32_t *goal,
                           double stay_penalty, double *tables_out);

/* ---- validate(config_init, connect, collide, check_goal, solution): utils.jl:320-367 -------- */
/* The transition checks of validate (:345-363) for a whole solution in two launches:
 * solution[T][N][d] (configuration per time step).  For t = 1..T-1 (0-based), in the
 * reference's order: connect(solution[t-1][i], solution[t][i], i) for i ascending, then
 * collide(solution[t-1], solution[t]).  ok_out = 1 when every check passes; otherwise the first
 * failure: t_out = t, kind_out = 1 (not connected, agent i_out) or 2 (colliding, first pair
 * i_out < j_out).  The initial-configuration and goal tests (:333-343) are O(N) equality
 * tests and stay with the caller. */
int mrmp_validate_solution(mrmp_ctx *ctx, int64_t T, const double *solution, int32_t *ok_out,
                           int64_t *t_out, int32_t *kind_out, int32_t *i_out, int32_t *j_out);

/* ---- first conflict of a joint plan: get_constraints (src/solvers/cbs.jl:374-412) ------------- */
/* configs[T][N][d] = convert_paths_to_configurations(paths) (solvers/utils.jl:128-134: agents
 * that arrived wait at their last vertex).  Scans t = 1..T-1 and pairs i < j in the reference's
 * order; per pair the vertex conflict collide(q_i_to, q_j_to, i, j) (:391) is tested before
 * the edge conflict collide(q_i_from, q_i_to, q_j_from, q_j_to, i, j) (:401).  t_out = -1 when
 * the plan is conflict free, else the step, kind_out = 1 (vertex) / 2 (edge) and the pair.
 * (With check_all_collisions the reference still returns at the first vertex conflict because
 * of `&` at :395; later conflicts are obtained by calling again on the remaining steps.) */
int mrmp_first_conflict(mrmp_ctx *ctx, int64_t T, const double *configs, int64_t *t_out,
                        int32_t *kind_out, int32_t *i_out, int32_t *j_out);

#ifdef __cplusplus
}
#endif
#endif /* MRMP_B200_H */
