#!/usr/bin/env python
"""bench.py -- connect+collide edge checks/sec and roadmap build ms (BASELINE.json metric).

Workload (configs[1]: StatePoint2D point2d_many, N=50, M=10): one STEP =
  (A) PRMs(num_vertices=500, rad=0.1): free-space sampling + neighbour pass for the N agents
      (sharded by agent across ranks, node tables + CSR all-gathered over NCCL);
  (B) a batch of explicit connect(q_from,q_to,i) edge checks (PRM-like edges);
  (C) a batch of collide(6-arity) checks between random roadmap edges of random agent
      pairs, referenced by (agent, vertex) ids into the device-resident roadmaps.
`value`  = (B)+(C) checks / step time, inputs resident in HBM (weak scaling: the per-GPU
           batch is fixed, the N=50 roadmap build is split across ranks);
`e2e`    = the same through the host-buffer C ABI: pinned host inputs, H2D/D2H inside the
           timed region, roadmaps (nodes + CSR) and verdicts read back every step;
`--impl reference` = the CPU oracle (C restatement of the reference, OpenMP, all host
           cores) on the identical workload.  The reference itself is pure Julia and cannot
           run in this image (no Julia runtime), see DESIGN.md.

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--log2-checks 24]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "connect+collide edge checks/sec"
UNIT = "checks/s"
N_AGENTS, NV, RAD = 50, 500, 0.1

# algorithmic figures (DESIGN.md "Kernels"): bytes per check of the explicit-list forms
BYTES_CONNECT = 2 * 16 + 4 + 1      # q_from, q_to (2 x double2), agent id, verdict byte
BYTES_COLLIDE_IDX = 6 * 4 + 4 * 16 + 1  # ids (6 x int32) + 4 gathered states + verdict


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2-checks", type=int, default=24, help="checks per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


# ------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the CPU oracle on the same workload
# ------------------------------------------------------------------------------------------
def cpu_reference_step(orc, inst, con, col, threads):
    """One full step on the host: roadmap build for all N agents + both check batches."""
    N, nv = inst.rads.size, inst.nv
    nodes = np.empty((N, nv, 2))
    nodes[:, 0], nodes[:, 1] = inst.q_init, inst.q_goal
    t0 = time.perf_counter()
    for a in range(N):  # prm.jl:193-199 per agent (sequential sampler loop, cheap)
        nodes[a, 2:], _ = orc.sample_free(a, inst.seed, nv - 2)
    row_ptr, colidx, nnz = orc.prm_build(0, nodes, inst.rad, nthreads=threads)
    t1 = time.perf_counter()
    v1 = orc.connect_edges(*con, nthreads=threads)
    v2 = orc.collide_edges_indexed(*col, nodes, nv, nthreads=threads)
    t2 = time.perf_counter()
    return t2 - t0, t1 - t0, int(v1.sum()) + int(v2.sum())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the host arm
    from oracle import oracle as O
    from sssp_b200 import workload as W
    O.build()
    threads = O.max_threads()
    inst = W.point2d_many(N_AGENTS, seed=0, nv=NV, rad=RAD)
    orc = O.Oracle(O.POINT2D, inst.rads, inst.obstacles)
    n = 1 << args.log2_checks
    n_con, n_col = n // 2, n - n // 2
    nodes = np.empty((N_AGENTS, NV, 2))
    nodes[:, 0], nodes[:, 1] = inst.q_init, inst.q_goal
    for a in range(N_AGENTS):
        nodes[a, 2:], _ = orc.sample_free(a, inst.seed, NV - 2)
    rp, colidx, nnz = orc.prm_build(0, nodes, inst.rad, nthreads=threads)
    row_ptr = np.zeros(N_AGENTS * NV + 1, np.int64)
    cols = []
    for a in range(N_AGENTS):
        row_ptr[a * NV + 1:(a + 1) * NV + 1] = row_ptr[a * NV] + rp[a, 1:]
        cols.append(colidx[a, :nnz[a]])
    cols = np.concatenate(cols)
    con = W.connect_list(inst, n_con)
    col = W.collide_list(inst, row_ptr, cols, n_col)
    times, builds = [], []
    for it in range(args.warmup + args.steps):
        t, tb, _ = cpu_reference_step(orc, inst, con, col, threads)
        if it >= args.warmup:
            times.append(t)
            builds.append(tb)
    T = float(np.sum(times))
    value = n * args.steps / T
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "point2d_many N=50 M=10: PRMs(500, rad=0.1) + 2^%d checks "
                               "(half connect edges, half indexed collide)" % args.log2_checks,
                   "n_agents": N_AGENTS, "num_vertices": NV, "rad": RAD,
                   "checks_per_step": n},
        "roadmap_build_ms": 1e3 * float(np.mean(builds)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "full step (N=50 roadmap build + 2^%d checks), C oracle, "
                                   "OpenMP %d threads; reference is pure Julia and cannot run "
                                   "here" % (args.log2_checks, threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.p, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                    "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


class CudaArray:
    """Zero-copy view of a raw device pointer for torch.as_tensor()."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 3}


# ------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import sssp_b200 as S
    from sssp_b200 import lib as L, sharding as SH, workload as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = L.load()

    inst = W.point2d_many(N_AGENTS, seed=0, nv=NV, rad=RAD)
    N, nv, d = N_AGENTS, NV, 2
    ctx = S.Context(inst.model, inst.rads, inst.obstacles, device=local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    a0, n_local, cap = SH.agent_block(N, world, rank)

    # device-resident roadmaps: the gathered node table is the NCCL buffer; this rank's shard
    # is a slice of it (kernels write straight into the send buffer), rm_all consumes it in place
    gathered = torch.zeros((world * cap, nv, d), dtype=torch.float64, device=dev)
    rm_all = S.Roadmaps(ctx, 0, N, nv)
    rm_all.bind_nodes(nv, gathered.data_ptr())
    rm_sh = S.Roadmaps(ctx, a0, max(n_local, 1), nv) if n_local > 0 else None
    if rm_sh is not None:
        rm_sh.bind_nodes(nv, gathered[a0].data_ptr())
    radv = np.full(max(n_local, 1), inst.rad)
    init_sh = np.ascontiguousarray(inst.q_init[a0:a0 + n_local])
    goal_sh = np.ascontiguousarray(inst.q_goal[a0:a0 + n_local])

    def build_roadmaps(gather_csr=True):
        """(A): sample + all-gather nodes + neighbour pass + all-gather CSR."""
        if rm_sh is not None:
            rm_sh.sample(nv, inst.seed, init_sh, goal_sh)
        if world > 1:
            SH.allgather_nodes_inplace(gathered, cap, nv, d)
        nnz = rm_sh.build(radv) if rm_sh is not None else 0
        if gather_csr and world > 1:
            if rm_sh is not None:
                p = rm_sh.device_ptrs()
                rp_t = torch.as_tensor(CudaArray(p["row_ptr"], (n_local * nv + 1,), "<i4"), device=dev)
                ci_t = torch.as_tensor(CudaArray(p["col_idx"], (max(nnz, 1),), "<i4"), device=dev)[:nnz]
            else:
                rp_t = torch.zeros(1, dtype=torch.int32, device=dev)
                ci_t = torch.zeros(0, dtype=torch.int32, device=dev)
            return SH.allgather_csr(rp_t, ci_t, N, nv, world, rank)
        return nnz

    # ---- setup: one build to obtain the CSR the collide list draws its edges from
    build_roadmaps(gather_csr=False)
    torch.cuda.synchronize()
    full = S.Roadmaps(ctx, 0, N, nv)
    full.sample(nv, inst.seed, inst.q_init, inst.q_goal)
    full.build(np.full(N, inst.rad))
    row_ptr_h, col_h = full.csr()
    nnz_total = int(col_h.size)
    full.close()

    n = 1 << args.log2_checks
    n_con, n_col = n // 2, n - n // 2
    con = W.connect_list(inst, n_con, seed=1234 + rank)
    col = W.collide_list(inst, row_ptr_h, col_h, n_col, seed=4321 + rank)

    # pinned host copies (e2e inputs / outputs) and device-resident copies (value)
    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t

    con_h = [pin(x) for x in con]
    col_hh = [pin(x) for x in col]
    con_d = [t.to(dev) for t in con_h]
    col_d = [t.to(dev) for t in col_hh]
    out_con_d = torch.empty(n_con, dtype=torch.uint8, device=dev)
    out_col_d = torch.empty(n_col, dtype=torch.uint8, device=dev)
    out_con_h = torch.empty(n_con, dtype=torch.uint8).pin_memory()
    out_col_h = torch.empty(n_col, dtype=torch.uint8).pin_memory()
    rows_sh = max(n_local, 1) * nv
    nodes_out_h = torch.empty((max(n_local, 1), nv, d), dtype=torch.float64).pin_memory()
    rowptr_out_h = torch.empty(rows_sh + 1, dtype=torch.int32).pin_memory()
    colidx_out_h = torch.empty(max(nnz_total, 1), dtype=torch.int32).pin_memory()
    st = stream.cuda_stream
    import ctypes as C

    def k_connect():
        L.check(lib.mrmp_connect_edges_dev(ctx._h, n_con, con_d[0].data_ptr(), con_d[1].data_ptr(),
                                           con_d[2].data_ptr(), out_con_d.data_ptr(), st))

    def k_collide():
        L.check(lib.mrmp_collide_edges_indexed_dev(
            rm_all._h, n_col, *[t.data_ptr() for t in col_d], out_col_d.data_ptr(), st))

    def step_value():
        build_roadmaps(gather_csr=True)
        k_connect()
        k_collide()

    def step_e2e():
        build_roadmaps(gather_csr=True)
        if rm_sh is not None:  # roadmaps back to the host (what PRMs() returns)
            L.check(lib.mrmp_roadmaps_get_nodes(rm_sh._h, C.c_void_p(nodes_out_h.data_ptr())))
            nnz = C.c_int64(0)
            L.check(lib.mrmp_roadmaps_get_csr(rm_sh._h, C.c_void_p(rowptr_out_h.data_ptr()),
                                              C.c_void_p(colidx_out_h.data_ptr()),
                                              colidx_out_h.numel(), C.byref(nnz)))
        L.check(lib.mrmp_connect_edges(ctx._h, n_con, C.c_void_p(con_h[0].data_ptr()),
                                       C.c_void_p(con_h[1].data_ptr()),
                                       C.c_void_p(con_h[2].data_ptr()),
                                       C.c_void_p(out_con_h.data_ptr())))
        L.check(lib.mrmp_collide_edges_indexed(rm_all._h, n_col,
                                               *[C.c_void_p(t.data_ptr()) for t in col_hh],
                                               C.c_void_p(out_col_h.data_ptr())))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: device-resident, CUDA events on the launching stream
    for _ in range(max(args.warmup, 3)):
        step_value()
    barrier()
    clocks = Clocks(local)
    clocks.start()
    launches0 = L.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_value()
    e1.record(stream)
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = L.kernel_launch_count() - launches0
    clk = clocks.stop()
    ms_step = ms_total / args.steps
    value = n * world / (ms_step * 1e-3)

    # ---- per-kernel durations (same stream, same inputs) for the roofline
    def timed(fn, reps=10):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    ms_con, ms_col = timed(k_connect), timed(k_collide)
    ms_build = timed(lambda: build_roadmaps(gather_csr=False), reps=5)
    pk, pk_kind = peaks()
    if ms_con >= ms_col:
        dom, dom_ms, dom_bytes = "k_connect_edges<2>", ms_con, n_con * BYTES_CONNECT
    else:
        dom, dom_ms, dom_bytes = "k_collide_edges_indexed<2>", ms_col, n_col * BYTES_COLLIDE_IDX
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": pk["hbm_gbs"],
                "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": None,
                "peak_source": pk_kind + " (MEASURED_PEAKS.json hbm_gbs)",
                "kernel_ms": {"connect_edges": ms_con, "collide_edges_indexed": ms_col,
                              "roadmap_build": ms_build},
                "fp64_unfused_peak_tflops": 18.58}

    # ---- e2e: host buffers through the C ABI
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    h2d = n_con * (4 + 32) + n_col * 24 + 2 * n_local * d * 8
    d2h = n + n_local * nv * d * 8 + (rows_sh + 1) * 4 + nnz_total * 4 // world
    e2e = {"value": n * world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3}

    # ---- cpu_baseline (rank 0, N=1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        threads = O.max_threads()
        orc = O.Oracle(O.POINT2D, inst.rads, inst.obstacles)
        t, tb, _ = cpu_reference_step(orc, inst, con, col, threads)   # warm-up
        ts = [cpu_reference_step(orc, inst, con, col, threads) for _ in range(3)]
        tbest = min(x[0] for x in ts)
        cpu = {"value": n / tbest, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "full step (N=50 roadmap build + 2^%d checks) x3, best; C oracle with "
                         "OpenMP on %d threads" % (args.log2_checks, threads),
               "roadmap_build_ms": 1e3 * min(x[1] for x in ts)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "point2d_many N=50 M=10: PRMs(500, rad=0.1) + 2^%d checks per "
                                   "GPU (half connect edges, half indexed collide)"
                                   % args.log2_checks,
                       "n_agents": N, "num_vertices": nv, "rad": RAD, "checks_per_step": n * world,
                       "l2": "inputs larger than L2 (%.0f MB per step per GPU)"
                             % ((n_con * 36 + n_col * 24) / 1e6)},
            "roadmap_build_ms": ms_build, "roadmap_nnz": nnz_total,
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
