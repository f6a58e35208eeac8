#!/usr/bin/env python
"""bench.py -- connect+collide edge checks/sec and roadmap build ms (BASELINE.json metric).

Workload = configs[1] of BASELINE.json: StatePoint2D point2d_many at N=50 agents, M=10
obstacles (scripts/config/eval/point2d_many.yaml).  One STEP is one pass of the hot path:

  (A) roadmap build  PRMs(num_vertices=500, rad=0.1)  (prm.jl:251-276): free-space sampling
      + all-pairs radius gate + connect(q_j,q_k,i) for the N agents -> CSR adjacency.
      connect checks counted: the pairs that pass the radius gate (= connect() calls the
      reference issues, prm.jl:207).
  (B) inter-agent collide batch in the shape PP's invalid() issues it (pp.jl:179-189): EVERY
      roadmap edge of every agent against the path edges of all OTHER agents over T=64 time
      steps (max_makespan of point2d_many.yaml), OR-reduced per time step -> the conflict
      table the search consults.  collide checks counted: nnz * T * (N-1).

Multi-GPU (weak scaling): every rank evaluates the full table against its OWN path scenario
(e.g. PP's random-priority restarts, pp.jl:104-128): per-GPU work is fixed.  The roadmap build
has two modes (--build): "sharded" = the neighbour pass is sharded by agent and the CSR shards
are exchanged by one NCCL all-gather of fixed-size slots (vertices are re-drawn on every rank
from the counter-based sampler instead of being communicated); "replicated" = every rank
builds all roadmaps itself, no exchange.  "auto" (default) times both at start-up and takes
the faster one; both timings are reported (`multi_gpu_build`).  The connect checks of the
build are counted once per step in either mode.

`value`  = checks / step time, inputs resident in HBM, CUDA events on the launching stream.
`e2e`    = same through the host-buffer C ABI: host init/goal and path ids in, roadmaps
           (nodes + CSR) and the conflict table bits out, every step.
`--impl reference` = the CPU oracle (C restatement of the reference, OpenMP on all host
           cores).  The reference is pure Julia and cannot run in this image;
           each step builds all N roadmaps and evaluates a bounded 1/64 row sample of (B).

  python bench.py --gpus N --steps K --warmup W [--impl reference]
"""
from __future__ import annotations

import argparse
import ctypes as C
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
N_AGENTS, NV, RAD, T_STEPS = 50, 500, 0.1, 64
CPU_ROW_STRIDE = 64            # reference arm: every 64th roadmap edge as a row of (B)

# algorithmic work per unit (DESIGN.md "Kernels"), FP64 flops with add/mul/cmp = 1 each
FLOPS_BROAD32 = 20             # FP32 capsule bound per tested pair (2-D, FMA = 2): w 2, |w|^2 3, w.u 3,
                               # excess 2, d^2 4, relative slack 2, reach 1, reach^2 + slack 2, compare 1
FLOPS_FILTER = 91              # FP64 verdict filter per surviving pair (2-D, FMA = 2): differences 10,
                               # three cross products 9, conditioning 5, crossing test 6, four clamped
                               # point-segment distances 52, minimum 3, band compare 6
FLOPS_EXACT = 170              # reference segment-segment test, SURVEY 8(d)
FP64_UNFUSED_PEAK_TFLOPS = 18.58   # tools/fp64_peak.cu on this pool (profiles/r1_fp64_peak.json)
NCU_DRAM_BYTES_PER_LAUNCH = 16.12e6   # profiles/r1_i_ncu_full_summary.md, last table (read 16.13 MB, written 16 KB)
BYTES_CONNECT = 2 * 16 + 4 + 1
BYTES_COLLIDE = 4 * 16 + 8 + 1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--build", default="auto", choices=["auto", "sharded", "replicated"],
                    help="multi-GPU roadmap build: neighbour pass sharded by agent + CSR "
                         "all-gather, or every rank builds all roadmaps (no exchange); auto "
                         "times both at start-up and takes the faster")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def config_dict(world):
    return {"workload": "point2d_many N=50 M=10: PRMs(500, rad=0.1) + conflict table (all "
                        "roadmap edges x other agents' path edges, T=64) per GPU",
            "n_agents": N_AGENTS, "num_vertices": NV, "rad": RAD, "path_steps": T_STEPS,
            "n_obstacles": 10, "path_scenarios": world,
            "parallelism": "single GPU" if world == 1 else
                           "weak scaling: one path scenario per GPU, roadmap build per --build",
            "l2": "L2 flushed between timed steps (256 MB memset); per-step inputs are KB-sized"}


# ------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the CPU oracle on the same workload
# ------------------------------------------------------------------------------------------
class CpuArm:
    def __init__(self, inst):
        from oracle import oracle as O
        O.build()
        self.O = O
        self.threads = O.host_cores()      # not OMP_NUM_THREADS: torchrun sets it to 1
        self.inst = inst
        self.orc = O.Oracle(O.POINT2D, inst.rads, inst.obstacles)

    def build(self):
        """(A) on the host: sampling loop per agent (prm.jl:193-199), neighbour pass with
        OpenMP over agents (the reference's `map` over agents, prm.jl:265-275)."""
        inst, N, nv = self.inst, self.inst.rads.size, self.inst.nv
        nodes = np.empty((N, nv, 2))
        nodes[:, 0], nodes[:, 1] = inst.q_init, inst.q_goal
        for a in range(N):
            nodes[a, 2:], _ = self.orc.sample_free(a, inst.seed, nv - 2)
        rp, colidx, nnz = self.orc.prm_build(0, nodes, inst.rad, nthreads=self.threads)
        row_ptr = np.zeros(N * nv + 1, np.int64)
        cols = []
        for a in range(N):
            row_ptr[a * nv + 1:(a + 1) * nv + 1] = row_ptr[a * nv] + rp[a, 1:]
            cols.append(colidx[a, :nnz[a]])
        return nodes, row_ptr, np.concatenate(cols)

    def table_sample(self, nodes, row_ptr, col, ca, vf, vt, stride):
        """(B) on a bounded row sample: every `stride`-th roadmap edge against all columns."""
        nv = self.inst.nv
        rows = np.repeat(np.arange(row_ptr.size - 1), np.diff(row_ptr))[::stride]
        ra = (rows // nv).astype(np.int32)
        rf = nodes[ra, rows % nv]
        rt = nodes[ra, col[::stride]]
        cf, ct = nodes[ca, vf], nodes[ca, vt]
        bits = self.orc.collide_cross(ra, rf, rt, ca, cf, ct, self.inst.rads.size,
                                      nthreads=self.threads)
        return rows.size, bits

    def step(self, ca, vf, vt, stride):
        t0 = time.perf_counter()
        nodes, row_ptr, col = self.build()
        t1 = time.perf_counter()
        n_rows, _ = self.table_sample(nodes, row_ptr, col, ca, vf, vt, stride)
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1, n_rows, nodes, row_ptr, col


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the host arm
    from sssp_b200 import workload as W
    inst = W.point2d_many(N_AGENTS, seed=0, nv=NV, rad=RAD)
    cpu = CpuArm(inst)
    nodes, row_ptr, col = cpu.build()
    _, ca, vf, vt = W.random_walk_paths(inst, row_ptr, col, T_STEPS, seed=4321)
    gate = W.gate_pass_count(nodes, RAD)
    tb, tt, checks = [], [], 0
    for it in range(args.warmup + args.steps):
        b, t, n_rows, *_ = cpu.step(ca, vf, vt, CPU_ROW_STRIDE)
        if it >= args.warmup:
            tb.append(b)
            tt.append(t)
            checks += gate + n_rows * T_STEPS * (N_AGENTS - 1)
    T = float(np.sum(tb) + np.sum(tt))
    value = checks / T
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_dict(1),
        "roadmap_build_ms": 1e3 * float(np.mean(tb)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.threads, "kind": "port",
                         "sample": "per step: full N=50 roadmap build + conflict table on every "
                                   "%dth roadmap edge (%d rows x %d columns); C oracle, OpenMP %d "
                                   "threads; the reference is pure Julia and cannot run here"
                                   % (CPU_ROW_STRIDE, n_rows, ca.size, cpu.threads)},
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
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            time.sleep(0.3)   # let the sampler come up before the timed region
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t_begin=None, t_end=None):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm, mx, reasons = [], [], set()
        for ts, r in self.rows:
            if t_begin is not None and not (t_begin - 0.06 <= ts <= t_end + 0.06):
                continue
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
    warmup = max(args.warmup, 3)

    inst = W.point2d_many(N_AGENTS, seed=0, nv=NV, rad=RAD)
    N, nv, d = N_AGENTS, NV, 2
    ctx = S.Context(inst.model, inst.rads, inst.obstacles, device=local)
    stream = torch.cuda.current_stream()
    st = stream.cuda_stream
    ctx.set_stream(st)
    a0, n_local, cap = SH.agent_block(N, world, rank)

    # Node tables: the sampler is counter-based (any try can be generated anywhere), so every
    # rank draws ALL agents' vertices itself (one launch, ~8 us) instead of all-gathering them;
    # the O(nv^2) neighbour pass -- the expensive part -- is sharded by agent.
    nodes_all = torch.zeros((N, nv, d), dtype=torch.float64, device=dev)
    rm_all = S.Roadmaps(ctx, 0, N, nv)
    rm_all.bind_nodes(nv, nodes_all.data_ptr())
    if world > 1:
        rm_sh = S.Roadmaps(ctx, a0, max(n_local, 1), nv)
        rm_sh.bind_nodes(nv, nodes_all[a0].data_ptr())
    else:
        rm_sh = rm_all
    radv = np.full(max(n_local, 1), inst.rad)
    # CSR exchange: ONE all-gather of fixed-size slots (cnt[cap*nv] | col[cap_nnz]); cap_nnz is
    # sized from a first build, with head room (a later overflow raises MRMP_ERR_CAPACITY)
    state = {"cap_nnz": 0, "slots": None}

    radv_all = np.full(N, inst.rad)
    build_mode = {"mode": "sharded" if world > 1 else "single"}

    def build_sharded():
        """(A): sample + neighbour pass of this rank's agents + [all-gather CSR shards]."""
        rm_all.sample(nv, inst.seed, inst.q_init, inst.q_goal, tries=False)
        nnz = rm_sh.build(radv)
        if world > 1:
            if state["slots"] is None:
                t = torch.tensor([nnz], dtype=torch.int64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                state["cap_nnz"] = int(t.item()) * 5 // 4 + 1024
                state["slots"] = torch.zeros((world, cap * nv + state["cap_nnz"]),
                                             dtype=torch.int32, device=dev)
            slots = state["slots"]
            L.check(lib.mrmp_roadmaps_pack_csr_shard_dev(rm_sh._h, cap, state["cap_nnz"],
                                                         slots[rank].data_ptr(), st))
            dist.all_gather_into_tensor(slots.view(-1), slots[rank])
            out = C.c_int64(0)
            L.check(lib.mrmp_roadmaps_set_csr_shards_dev(rm_all._h, world, cap, state["cap_nnz"],
                                                         slots.data_ptr(), st, C.byref(out)))
            return int(out.value)
        return nnz

    def build_replicated():
        """(A'): every rank builds all N roadmaps itself -- no exchange.  Cheaper than the
        sharded build whenever the whole build takes less than the all-gather round trip."""
        rm_all.sample(nv, inst.seed, inst.q_init, inst.q_goal, tries=False)
        return rm_all.build(radv_all)

    def build_roadmaps():
        return build_replicated() if build_mode["mode"] == "replicated" else build_sharded()

    # ---- setup: one build gives the CSR the synthetic paths walk on
    nnz_total = build_roadmaps()
    torch.cuda.synchronize()
    build_modes_ms = None
    if world > 1:
        def wall_ms(fn, reps=10):
            for _ in range(3):
                fn()
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            t = torch.tensor([(time.perf_counter() - t0) / reps * 1e3], dtype=torch.float64,
                             device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        build_modes_ms = {"sharded": wall_ms(build_sharded), "replicated": wall_ms(build_replicated)}
        if args.build == "auto":
            build_mode["mode"] = min(build_modes_ms, key=build_modes_ms.get)
        else:
            build_mode["mode"] = args.build
        assert build_roadmaps() == nnz_total   # both modes give the same roadmaps
    nodes_h = rm_all.nodes()
    row_ptr_h, col_h = rm_all.csr()
    gate = W.gate_pass_count(nodes_h, RAD)
    _, ca, vf, vt = W.random_walk_paths(inst, row_ptr_h, col_h, T_STEPS, seed=4321 + rank)
    n_cols = int(ca.size)
    pair_checks = nnz_total * T_STEPS * (N - 1)
    wpr = (T_STEPS + 31) // 32

    def pin(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory()

    cols_h = [pin(x) for x in (ca, vf, vt)]
    cols_d = [t.to(dev) for t in cols_h]
    bits_d = torch.zeros((nnz_total, wpr), dtype=torch.int32, device=dev)
    bits_h = torch.zeros((nnz_total, wpr), dtype=torch.int32).pin_memory()
    # e2e returns the roadmaps this rank built to the host: its shard, or all of them
    rm_ret = rm_all if build_mode["mode"] == "replicated" else rm_sh
    n_ret = N if build_mode["mode"] == "replicated" else max(n_local, 1)
    rows_sh = n_ret * nv
    nodes_out_h = torch.empty((n_ret, nv, d), dtype=torch.float64).pin_memory()
    rowptr_out_h = torch.empty(rows_sh + 1, dtype=torch.int32).pin_memory()
    colidx_out_h = torch.empty(max(nnz_total, 1), dtype=torch.int32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def k_table():
        L.check(lib.mrmp_roadmaps_edge_conflicts_dev(rm_all._h, None, n_cols,
                                                     cols_d[0].data_ptr(), cols_d[1].data_ptr(),
                                                     cols_d[2].data_ptr(), N, bits_d.data_ptr(), st))

    def step_value():
        build_roadmaps()
        k_table()

    def step_e2e():
        build_roadmaps()
        # roadmaps back to the host (what PRMs() returns to the solver): on the copy stream,
        # beside the conflict-table kernel
        nnz = C.c_int64(0)
        L.check(lib.mrmp_roadmaps_read_async(rm_ret._h, nodes_out_h.data_ptr(),
                                             rowptr_out_h.data_ptr(), colidx_out_h.data_ptr(),
                                             colidx_out_h.numel(), C.byref(nnz)))
        L.check(lib.mrmp_roadmaps_edge_conflicts(rm_all._h, None, n_cols, cols_h[0].data_ptr(),
                                                 cols_h[1].data_ptr(), cols_h[2].data_ptr(), N,
                                                 bits_h.data_ptr(), bits_h.numel()))
        L.check(lib.mrmp_roadmaps_read_wait(rm_ret._h))

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

    # ---- value: device-resident, CUDA events per step on the launching stream, L2 flushed
    for _ in range(warmup):
        step_value()
    barrier()
    clocks = Clocks(local)
    clocks.start()
    launches0 = L.kernel_launch_count()
    t_begin = time.perf_counter()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    for k in range(args.steps):
        flush.zero_()                      # evict L2 (outside the per-step event pair)
        ev[k][0].record(stream)
        step_value()
        ev[k][1].record(stream)
    barrier()
    t_end = time.perf_counter()
    launches = L.kernel_launch_count() - launches0
    ms_total = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
    clk = clocks.stop(t_begin, t_end)
    ms_step = ms_total / args.steps
    checks_step = gate + world * pair_checks
    value = checks_step / (ms_step * 1e-3)

    # ---- per-kernel durations for the roofline (same stream, same inputs)
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

    ms_build = timed(build_roadmaps, reps=5)
    ctx.stats(reset=True)
    ms_table = timed(k_table, reps=10)
    st_ = ctx.stats()
    exact = int(st_[0]) / 11.0            # pairs that ran the reference's own arithmetic
    tested = int(st_[2]) / 11.0           # pairs that ran the FP32 bound (rest: tile boxes)
    survivors = int(st_[3]) / 11.0        # pairs that ran the FP64 verdict filter
    n_pairs = nnz_total * n_cols
    flops = survivors * FLOPS_FILTER + exact * FLOPS_EXACT      # FP64 work only
    flops32 = tested * FLOPS_BROAD32
    achieved = flops / (ms_table * 1e-3) / 1e12
    pk, pk_kind = peaks()
    roofline = {
        "bound": "fp64", "kernel": "k_cross_collide<2,true> (+ column sort/prep)",
        "achieved": achieved, "peak": FP64_UNFUSED_PEAK_TFLOPS, "unit": "TFLOP/s",
        "frac": achieved / FP64_UNFUSED_PEAK_TFLOPS, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
        "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` "
                          "capture of this kernel on this workload (profiles/r1_i_ncu_full_summary.md); "
                          "algorithmic bytes per launch: 17.0e6",
        "peak_source": "un-fused DADD/DMUL issue peak measured with tools/fp64_peak.cu on this "
                       "pool (profiles/r1_fp64_peak.json); MEASURED_PEAKS.json has no FP64 figure "
                       "and bit parity forbids FMA",
        "algorithmic_flops": {"pairs": n_pairs, "pairs_bound_tested_fp32": tested,
                              "fp32_per_tested_pair": FLOPS_BROAD32, "fp32_tflops": flops32 / (ms_table * 1e-3) / 1e12,
                              "pairs_filtered_fp64": survivors, "fp64_per_filtered_pair": FLOPS_FILTER,
                              "pairs_reference_arithmetic": exact, "fp64_per_exact_pair": FLOPS_EXACT,
                              "note": "work the kernel performs: `achieved` counts the FP64 part "
                                      "(verdict filter x 91 + reference arithmetic x 170, FMA = 2); "
                                      "the FP32 broad phase (20 per tested pair) is listed beside it; "
                                      "the reference itself spends 170 FP64 per pair"},
        "reference_equivalent_tflops": n_pairs * FLOPS_EXACT / (ms_table * 1e-3) / 1e12,
        "pair_checks_per_s": n_pairs / (ms_table * 1e-3),
        "kernel_ms": {"conflict_table": ms_table, "roadmap_build": ms_build},
    }

    # ---- e2e: host buffers through the C ABI
    for _ in range(2):
        step_e2e()
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    h2d = 2 * N * d * 8 + 3 * n_cols * 4
    nnz_ret = nnz_total if build_mode["mode"] == "replicated" else nnz_total // world
    d2h = n_ret * nv * d * 8 + (rows_sh + 1) * 4 + nnz_ret * 4 + nnz_total * wpr * 4
    e2e = {"value": checks_step / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3}

    # ---- extras: explicit-list kernels (HBM-side of the path), inputs resident in HBM
    extras = None
    if rank == 0 and not args.no_extras:
        n = 1 << 24
        con = W.connect_list(inst, n, seed=1234)
        con_d = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in con]
        cl = W.collide_list(inst, row_ptr_h, col_h, n, seed=99)
        cl_d = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in cl]
        out = torch.empty(n, dtype=torch.uint8, device=dev)
        ms_con = timed(lambda: L.check(lib.mrmp_connect_edges_dev(
            ctx._h, n, con_d[0].data_ptr(), con_d[1].data_ptr(), con_d[2].data_ptr(),
            out.data_ptr(), st)))
        ms_idx = timed(lambda: L.check(lib.mrmp_collide_edges_indexed_dev(
            rm_all._h, n, *[t.data_ptr() for t in cl_d], out.data_ptr(), st)))
        extras = {"explicit_list_n": n,
                  "connect_edges": {"ms": ms_con, "checks_per_s": n / ms_con * 1e3,
                                    "hbm_gbs": n * BYTES_CONNECT / ms_con / 1e6,
                                    "hbm_frac": n * BYTES_CONNECT / ms_con / 1e6 / pk["hbm_gbs"]},
                  "collide_edges_indexed": {"ms": ms_idx, "checks_per_s": n / ms_idx * 1e3}}
        del con_d, cl_d, out
        extras["sssp_expansion"] = extra_sssp_expansion(inst, nodes_h, local)
        extras["arm22"] = extra_arm22(local, dev, st, timed)

    # ---- cpu_baseline (rank 0, N=1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        arm = CpuArm(inst)
        arm.step(ca, vf, vt, CPU_ROW_STRIDE)      # warm-up
        best = None
        for _ in range(3):
            b, t, n_rows, *_ = arm.step(ca, vf, vt, CPU_ROW_STRIDE)
            if best is None or b + t < best[0] + best[1]:
                best = (b, t, n_rows)
        cchecks = gate + best[2] * T_STEPS * (N - 1)
        cpu = {"value": cchecks / (best[0] + best[1]), "unit": UNIT, "cores": arm.threads,
               "kind": "port",
               "sample": "full N=50 roadmap build + conflict table on every %dth roadmap edge "
                         "(%d rows x %d columns), best of 3; C oracle, OpenMP %d threads"
                         % (CPU_ROW_STRIDE, best[2], n_cols, arm.threads),
               "roadmap_build_ms": 1e3 * best[0],
               "collide_pair_checks_per_s": best[2] * T_STEPS * (N - 1) / best[1]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(world),
            "checks_per_step": {"connect": gate, "collide": world * pair_checks},
            "roadmap_build_ms": ms_build, "roadmap_nnz": nnz_total,
            "multi_gpu_build": None if world == 1 else dict(mode=build_mode["mode"],
                                                            wall_ms=build_modes_ms),
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "extras": extras,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def extra_sssp_expansion(inst, nodes_h, device):
    """SURVEY 8(d) config 1/2 deliverable 'ms per expansion': the fused node expansion
    (mrmp_sssp_expand: K=10 samples, steering depth 2, epsilon 0.2 as in point2d_many.yaml,
    roadmaps of 300 vertices per agent, N=50) through the host API, against the same
    expansion in C on one host core (a reference solver call is single-threaded)."""
    import sssp_b200 as S
    from sssp_b200 import lib as L
    from oracle import oracle as O
    N, nv0, K = inst.rads.size, 300, 10
    ctx = S.Context(inst.model, inst.rads, inst.obstacles, device=device)
    orc = O.Oracle(O.POINT2D, inst.rads, inst.obstacles)
    dev_rm = S.SsspRoadmaps(ctx, nv_cap=1024)
    tabs = [np.ascontiguousarray(nodes_h[a, :nv0]) for a in range(N)]
    for a in range(N):
        dev_rm.set_roadmap(a, tabs[a])
    rng = np.random.default_rng(7)
    calls = []
    for k in range(40):
        a = k % N
        ids = rng.integers(0, nv0, N).astype(np.int32)
        old = rng.choice(nv0, 12, replace=False).astype(np.int32)
        calls.append((a, int(ids[a]), ids, old, rng.random((K, 2))))
    # timed through the C ABI itself (the drop-in boundary) with caller-owned buffers, exactly
    # what the Julia glue passes; decoding the bitmaps is host bookkeeping outside the call
    lib = L.load()
    words_cap = (nv0 + 64 * K + 31) // 32
    q_new = np.empty((K, 2))
    ob_g, ib_g = np.zeros((K, words_cap), np.uint32), np.zeros((K, words_cap), np.uint32)
    succ, hit = np.empty(12 + K + 1, np.int32), np.empty(12 + K + 1, np.uint8)
    n_new, n_succ, w = C.c_int32(0), C.c_int32(0), C.c_int32(0)

    def call(a, v, ids, old, qr):
        L.check(lib.mrmp_sssp_expand(dev_rm._h, a, v, K, qr.ctypes.data, 0, 0, 2, 0.2, 0.02,
                                     ids.ctypes.data, old.size, old.ctypes.data, 0, C.byref(n_new),
                                     q_new.ctypes.data, C.byref(w), ob_g.ctypes.data, ib_g.ctypes.data,
                                     ob_g.size, C.byref(n_succ), succ.ctypes.data, hit.ctypes.data))

    call(*calls[0])
    tabs[calls[0][0]] = dev_rm.nodes(calls[0][0])
    t_gpu, t_cpu, mism, n_acc = 0.0, 0.0, 0, 0
    for a, v, ids, old, qr in calls[1:]:
        Q = np.stack([tabs[j][ids[j]] for j in range(N)])
        t0 = time.perf_counter()
        call(a, v, ids, old, qr)
        t1 = time.perf_counter()
        after, ob, ib, osucc, ohit = orc.sssp_expand(a, tabs[a], v, qr, 2, 0.2, 0.02, Q, old)
        t2 = time.perf_counter()
        t_gpu += t1 - t0
        t_cpu += t2 - t1
        k, ns, wd_ = n_new.value, n_succ.value, w.value
        ok = (after.shape[0] - tabs[a].shape[0] == k and np.array_equal(after[tabs[a].shape[0]:], q_new[:k])
              and np.array_equal(succ[:ns], osucc) and np.array_equal(hit[:ns], ohit)
              and np.array_equal(ob_g.reshape(-1)[:K * wd_].reshape(K, wd_)[:k, :ob.shape[1]], ob)
              and np.array_equal(ib_g.reshape(-1)[:K * wd_].reshape(K, wd_)[:k, :ib.shape[1]], ib))
        mism += 0 if ok else 1
        n_acc += k
        tabs[a] = after
    n_new = n_acc
    n = len(calls) - 1
    return {"calls": n, "K": K, "n_agents": int(N), "roadmap_vertices": nv0,
            "gpu_ms_per_expansion": 1e3 * t_gpu / n, "cpu_c_1core_ms_per_expansion": 1e3 * t_cpu / n,
            "timed": "mrmp_sssp_expand through the C ABI (host buffers in, results out)",
            "accepted_vertices": n_new, "mismatching_calls": mism}


def extra_arm22(device, dev, st, timed):
    """BASELINE configs[3]: StateArm22 N=10 (link 0.15-0.25, 3 obstacles): swept-link collide
    (6-arity, arm22.jl:100-146) and edge connect (arm22.jl:56-86) on uniformly random joint
    pairs, inputs resident in HBM; CPU = C oracle with OpenMP on a 1/16 sample."""
    import torch
    import sssp_b200 as S
    from sssp_b200 import lib as L
    from oracle import oracle as O
    lib = L.load()
    rng = np.random.default_rng(22)
    N, M = 10, 3
    rads = rng.uniform(0.15, 0.25, N)
    base = rng.uniform(0.3, 0.7, (N, 2))
    obs = np.concatenate([rng.random((M, 2)), rng.uniform(0.05, 0.1, (M, 1))], axis=1)
    ctx = S.Context(L.MODEL_ARM22, rads, obs, base_pos=base, device=device)
    orc = O.Oracle(O.ARM22, rads, obs, base=base)
    n_col, n_con = 1 << 16, 1 << 18
    ai = rng.integers(0, N, n_col).astype(np.int32)
    aj = ((ai + 1 + rng.integers(0, N - 1, n_col)) % N).astype(np.int32)
    q = [rng.uniform(0, 2 * np.pi, (n_col, 2)) for _ in range(4)]
    ag = rng.integers(0, N, n_con).astype(np.int32)
    cf, ct = rng.uniform(0, 2 * np.pi, (n_con, 2)), rng.uniform(0, 2 * np.pi, (n_con, 2))
    d_ = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (ai, aj, *q, ag, cf, ct)]
    out = torch.empty(max(n_col, n_con), dtype=torch.uint8, device=dev)
    ctx.set_stream(st)
    ms_col = timed(lambda: L.check(lib.mrmp_collide_pairs_dev(
        ctx._h, n_col, *[t.data_ptr() for t in d_[:6]], out.data_ptr(), st)), reps=3)
    g_col = out[:n_col].cpu().numpy().copy()
    ms_con = timed(lambda: L.check(lib.mrmp_connect_edges_dev(
        ctx._h, n_con, d_[6].data_ptr(), d_[7].data_ptr(), d_[8].data_ptr(), out.data_ptr(), st)), reps=3)
    g_con = out[:n_con].cpu().numpy().copy()
    sl = slice(0, None, 16)
    t0 = time.perf_counter()
    nt = O.host_cores()
    o_col = orc.collide_pairs(ai[sl], aj[sl], *[x[sl] for x in q], nthreads=nt)
    t1 = time.perf_counter()
    o_con = orc.connect_edges(ag[sl], cf[sl], ct[sl], nthreads=nt)
    t2 = time.perf_counter()
    return {"n_agents": N, "n_obstacles": M,
            "collide_pairs": {"n": n_col, "gpu_ms": ms_col, "gpu_checks_per_s": n_col / ms_col * 1e3,
                              "cpu_checks_per_s": o_col.size / (t1 - t0), "cpu_threads": nt,
                              "mismatches_on_sample": int((g_col[sl] != o_col).sum()),
                              "hit_rate": float(g_col.mean())},
            "connect_edges": {"n": n_con, "gpu_ms": ms_con, "gpu_checks_per_s": n_con / ms_con * 1e3,
                              "cpu_checks_per_s": o_con.size / (t2 - t1),
                              "mismatches_on_sample": int((g_con[sl] != o_con).sum()),
                              "free_rate": float(g_con.mean())}}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
