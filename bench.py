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

Multi-GPU (--gpus N > 1, default --scaling strong): ONE scenario is strong-scaled.  Agents are
block-partitioned over the ranks (prm.jl:265-275: the roadmaps are independent).  Every rank draws
all vertices (counter-based sampler: nothing to communicate), builds the roadmaps of ITS agents,
writes its CSR shard straight into every peer's memory over NVLink (peer mappings, one flag per
peer -- sssp_b200/peer.py; NCCL all-gather only as a fallback or with --exchange nccl), computes
the table rows of ITS agents' edges against all agents' path edges, copies them into rank 0's
memory and adopts the peers' shards -- one call into the library per rank and step
(mrmp_roadmaps_strong_step; --separate-calls issues one call per stage, which is how the stage
timeline in `multi_gpu.timeline` is taken).  The step ends -- inside the CUDA-event region --
when every rank holds all roadmaps and rank 0 the whole table; both are checked byte for byte
against the 1-GPU result before timing.  checks/step are those of the one scenario; `value` =
checks / max-over-ranks step time.  --scaling weak keeps the round-1 shape (one scenario per
rank on a sharded build) as an extra.

`value`  = checks / step time, inputs resident in HBM, CUDA events on the launching stream (the
           ctx is put on torch's current stream: Context.set_stream maps the default stream's
           handle 0 to cudaStreamLegacy, so flush, events and every launch share ONE stream).
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
NCU_DRAM_BYTES_PER_LAUNCH = 19.37e6   # profiles/r2_r_cross_ncu_summary.md (read 19.36 MB, written 9 KB: the table stays in L2)
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
    ap.add_argument("--separate-calls", action="store_true",
                    help="multi-GPU strong step as one library call per stage (the form the stage "
                         "timeline is taken with) instead of mrmp_roadmaps_strong_step")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="multi-GPU: strong = ONE scenario, roadmaps and table rows sharded by agent "
                         "(default); weak = one scenario per GPU on a sharded build")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU exchange: peer stores over NVLink from the producing kernels "
                         "(falls back to nccl when the peer mapping cannot be opened), or NCCL all-gather")
    ap.add_argument("--num-vertices", type=int, default=NV,
                    help="vertices per roadmap (500 = point2d_many.yaml; larger sizes for scaling studies)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def config_dict(world, strong=True, nv=NV):
    return {"workload": "point2d_many N=50 M=10: PRMs(%d, rad=0.1) + conflict table (all "
                        "roadmap edges x other agents' path edges, T=64)" % nv,
            "n_agents": N_AGENTS, "num_vertices": nv, "rad": RAD, "path_steps": T_STEPS,
            "n_obstacles": 10, "path_scenarios": 1 if (strong or world == 1) else world,
            "parallelism": "single GPU" if world == 1 else (
                "strong scaling of ONE scenario: agents block-partitioned over the GPUs, each rank "
                "builds its agents' roadmaps and computes their table rows; CSR shards and table "
                "rows are exchanged inside the timed step" if strong else
                "weak scaling: one path scenario per GPU, sharded roadmap build + exchange"),
            "l2": "L2 flushed between timed steps (256 MB memset); per-step inputs are KB-sized"}


# ------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the CPU oracle on the same workload
# ------------------------------------------------------------------------------------------
class CpuArm:
    def __init__(self, inst, threads=None):
        from oracle import oracle as O
        O.build()
        self.O = O
        self.threads = threads or O.host_cores()   # not OMP_NUM_THREADS: torchrun sets it to 1
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
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_dict(1),
        "cores_available": os.cpu_count(),
        "roadmap_build_ms": 1e3 * float(np.mean(tb)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.threads, "kind": "port",
                         "sample": "per step: full N=50 roadmap build + conflict table on every "
                                   "%dth roadmap edge (%d rows x %d columns); C oracle, OpenMP %d "
                                   "threads; the reference is pure Julia and cannot run here"
                                   % (CPU_ROW_STRIDE, n_rows, ca.size, cpu.threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    jl = julia_reference()
    if jl is not None:
        line["julia"] = jl
    print(json.dumps(line), flush=True)


def julia_reference():
    """when a `julia` binary and a checkout of the reference (MRMP_REFERENCE_DIR) are
    present, run the REAL reference on the same workload (julia/ref_bench.jl over a fixture of
    tools/export_ref_fixture.py): its timings and its agreement with the oracle ("oracle_pinned")
    travel in the line under "julia".  Neither exists in this image or on the GPU box: None."""
    import shutil
    import tempfile
    jl, ref = shutil.which("julia"), os.environ.get("MRMP_REFERENCE_DIR")
    if not jl or not ref or not os.path.isdir(ref):
        return None
    try:
        root = os.path.dirname(os.path.abspath(__file__))
        sys.path.insert(0, os.path.join(root, "tools"))
        import export_ref_fixture as E
        fx = E.make(N_AGENTS, NV, RAD, T_STEPS, CPU_ROW_STRIDE, 2000, 0)
        with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
            json.dump(fx, f)
        p = subprocess.run([jl, "--project=" + ref, "--threads=auto",
                            os.path.join(root, "julia", "ref_bench.jl"), f.name, "3"],
                           capture_output=True, text=True, timeout=3600)
        os.unlink(f.name)
        lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
        if p.returncode == 0 and lines:
            return json.loads(lines[-1])
        return {"error": (p.stderr or p.stdout)[-400:]}
    except Exception as e:   # never lets the optional leg break the arm
        return {"error": str(e)[:400]}


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
    from sssp_b200 import lib as L, peer as P, sharding as SH, workload as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = L.load()
    warmup = max(args.warmup, 3)
    strong = args.scaling == "strong"

    N, nv, d = N_AGENTS, args.num_vertices, 2
    inst = W.point2d_many(N, seed=0, nv=nv, rad=RAD)
    ctx = S.Context(inst.model, inst.rads, inst.obstacles, device=local)
    stream = torch.cuda.current_stream()
    st = stream.cuda_stream
    ctx.set_stream(st)
    a0, n_local, cap = SH.agent_block(N, world, rank)

    # Node tables: the sampler is counter-based (any try can be generated anywhere), so every
    # rank draws ALL agents' vertices itself (one launch, ~8 us) instead of receiving them;
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
    radv_all = np.full(N, inst.rad)
    wpr = (T_STEPS + 31) // 32

    def max_over_ranks(x, op=None):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op or dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- setup: a full build on every rank gives the reference CSR, the synthetic paths and (for
    # the multi-GPU modes) the table that the sharded computation must reproduce byte for byte
    rm_all.sample(nv, inst.seed, inst.q_init, inst.q_goal, tries=False)
    nnz_total = rm_all.build(radv_all)
    nodes_h = rm_all.nodes()
    row_ptr_h, col_h = rm_all.csr()
    gate = W.gate_pass_count(nodes_h, RAD)
    scen = 4321 if strong else 4321 + rank        # strong: ONE scenario shared by all ranks
    _, ca, vf, vt = W.random_walk_paths(inst, row_ptr_h, col_h, T_STEPS, seed=scen)
    n_cols = int(ca.size)
    pair_checks = nnz_total * T_STEPS * (N - 1)

    def pin(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory()

    cols_h = [pin(x) for x in (ca, vf, vt)]
    cols_d = [t.to(dev) for t in cols_h]
    bits_d = torch.zeros((nnz_total + 4096, wpr), dtype=torch.int32, device=dev)[:nnz_total]
    bits_h = torch.zeros((nnz_total, wpr), dtype=torch.int32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def k_table_full():
        L.check(lib.mrmp_roadmaps_edge_conflicts_dev(rm_all._h, None, n_cols,
                                                     cols_d[0].data_ptr(), cols_d[1].data_ptr(),
                                                     cols_d[2].data_ptr(), N, bits_d.data_ptr(), st))

    # ---- multi-GPU plumbing
    mg = None
    if world > 1:
        nnz_sh = rm_sh.build(radv)
        nnz_max = int(max_over_ranks(float(nnz_sh)))
        cap_nnz = P.csr_slot_words(cap, nv, nnz_max)
        tab_words = (nnz_max + nnz_max // 4 + 1024) * wpr
        tab_words += (-tab_words) % 4
        ex = P.Exchange(ctx, world, rank, [4 * (cap * nv + cap_nnz), 4 * tab_words])
        p2p = ex.connect() and args.exchange != "nccl"
        slots = None if p2p else torch.zeros((world, cap * nv + cap_nnz), dtype=torch.int32, device=dev)
        tab_slots = None if p2p else torch.zeros((world, tab_words), dtype=torch.int32, device=dev)
        nnz_all = torch.zeros(world, dtype=torch.int64, device=dev)
        nnz_all[rank] = nnz_sh
        dist.all_reduce(nnz_all)
        mg = dict(ex=ex, p2p=p2p, round=0, cap_nnz=cap_nnz, steps={},
                  tab_local=torch.zeros(tab_words, dtype=torch.int32, device=dev), tab_words=tab_words, slots=slots,
                  tab_slots=tab_slots, nnz_sh=[int(x) for x in nnz_all.tolist()],
                  exchange="nvlink peer stores from the producing kernels (CUDA IPC mappings) + flags"
                  if p2p else "nccl all-gather (%s)" % (getattr(ex, "connect_error", None) or args.exchange))

    tab_rows_cap = 0 if mg is None else mg["tab_words"] // wpr

    def exchange_csr():
        """this rank's CSR shard -> every rank; afterwards rm_all holds all roadmaps"""
        ex = mg["ex"]
        if mg["p2p"]:
            if mg["round"] > 0:      # next epoch of both channels (slots are double buffered)
                ex.advance(0)
                ex.advance(1)
            mg["round"] += 1
            ex.push_csr_shard(rm_sh, 0, cap, mg["cap_nnz"], st)
            return None
        rm_sh.finish()
        L.check(lib.mrmp_roadmaps_pack_csr_shard_dev(rm_sh._h, cap, mg["cap_nnz"],
                                                     mg["slots"][rank].data_ptr(), st))
        dist.all_gather_into_tensor(mg["slots"].view(-1), mg["slots"][rank])
        return None

    def adopt_csr():
        ex = mg["ex"]
        if mg["p2p"]:
            ex.adopt_csr_shards_async(rm_all, 0, cap, mg["cap_nnz"], st)
            return None
        out = C.c_int64(0)
        L.check(lib.mrmp_roadmaps_set_csr_shards_dev(rm_all._h, world, cap, mg["cap_nnz"],
                                                     mg["slots"].data_ptr(), st, C.byref(out)))
        return int(out.value)

    def step_strong(cols=cols_d, marks=None):
        """ONE scenario over all ranks: draw all vertices, build the roadmaps of this rank's agent
        block, publish the CSR shard to every rank, evaluate the table rows of this rank's agents
        against all agents' path edges and send them to rank 0, adopt the peers' shards.  With the
        peer exchange nothing waits on the host until the end of the step (asynchronous build: the
        edge count stays on the device).  The timed region ends when every rank holds all roadmaps
        and rank 0 the table.  marks: optional list that receives a CUDA event after every stage."""
        ex = mg["ex"]
        if mg["p2p"] and marks is None and not args.separate_calls:
            # the whole step of this rank is ONE call into the library (mrmp_roadmaps_strong_step:
            # the same launches as below, adoption of the peers' shards on the ctx's second stream)
            key = cols[0].data_ptr()
            ss = mg["steps"].get(key)
            if ss is None:
                ss = mg["steps"][key] = P.StrongStep(
                    ex, rm_all, rm_sh, nv, inst.seed, inst.q_init, inst.q_goal, radv, cap, mg["cap_nnz"],
                    n_cols, cols[0].data_ptr(), cols[1].data_ptr(), cols[2].data_ptr(), N,
                    mg["tab_local"].data_ptr(), wpr, tab_rows_cap, table_dst=0)
            t_a = time.perf_counter()
            ss.enqueue(advance=mg["round"] > 0)
            mg["round"] += 1
            t_b = time.perf_counter()
            n = ss.finish()
            mg["host_s"] = (t_b - t_a, time.perf_counter() - t_b)   # (enqueue, waits) of the last step
            return n

        def mark():
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(stream)
                marks.append(e)

        mark()
        rm_all.sample(nv, inst.seed, inst.q_init, inst.q_goal, tries=False)
        rm_sh.build_async(radv)
        mark()                                   # 0: sample + neighbour pass of this rank's agents
        exchange_csr()
        mark()                                   # 1: shard packed and pushed / all-gathered
        if mg["p2p"]:
            # rows of this rank's agents: into local memory (a rank's share of the table is small,
            # the kernel cuts it into finer work items that OR into local memory), then one
            # coalesced copy kernel into rank 0's buffer over NVLink + a flag
            rm_sh.edge_conflicts_async(n_cols, cols[0].data_ptr(), cols[1].data_ptr(), cols[2].data_ptr(),
                                       N, mg["tab_local"].data_ptr(), tab_rows_cap, st, cols=rm_all)
            ex.push_table_rows(rm_sh, 1, 0, mg["tab_local"].data_ptr(), wpr, tab_rows_cap, st)
            ex.signal(1, st, dst_mask=1)
            mark()                               # 2: table rows of this rank (sent to rank 0)
            adopt_csr()
            mark()                               # 3: peers' shards arrived and stitched
            if rank == 0:
                ex.wait(1, st)
            mark()                               # 4: rank 0: all table rows arrived
            rm_sh.finish()                       # the only host waits of the step
            n = rm_all.finish()
            if ex.error():
                raise RuntimeError("exchange: rank %d never published" % (ex.error() - 1))
        else:
            rm_sh.edge_conflicts_dev(n_cols, cols[0].data_ptr(), cols[1].data_ptr(), cols[2].data_ptr(),
                                     N, mg["tab_slots"][rank].data_ptr(), st, cols=rm_all)
            mark()
            n = adopt_csr()
            mark()
            dist.all_gather_into_tensor(mg["tab_slots"].view(-1), mg["tab_slots"][rank])
            mark()
        return n

    def stage_timeline(reps=10):
        """per-stage CUDA-event durations of the strong step on this rank, mean over reps, gathered
        from every rank (there is no nsys in this image: the events are the timeline)"""
        names = ["build_own_agents", "publish_csr_shard", "table_rows_own_agents",
                 "adopt_peer_shards", "wait_table_rows"]
        acc = np.zeros(len(names))
        for _ in range(reps):
            flush.zero_()
            dist.barrier()
            marks = []
            step_strong(marks=marks)
            torch.cuda.synchronize()
            acc += np.array([marks[k].elapsed_time(marks[k + 1]) for k in range(len(names))])
        t = torch.tensor(acc / reps, dtype=torch.float64, device=dev)
        allt = torch.empty((world, len(names)), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allt.view(-1), t)
        return {"stages": names, "ms_by_rank": [[round(float(x), 4) for x in row] for row in allt.tolist()],
                "rows_by_rank": mg["nnz_sh"]}

    def gathered_table():
        """rank 0: the table rows of all ranks in packed CSR order (after a strong step)"""
        if mg["p2p"]:
            sl = torch.as_tensor(CudaArray(mg["ex"].slots(1), (world, mg["tab_words"]), "<i4"), device=dev)
        else:
            sl = mg["tab_slots"]
        return torch.cat([sl[r, : mg["nnz_sh"][r] * wpr].view(-1, wpr) for r in range(world)])

    def step_weak():
        """every rank its own scenario on all roadmaps; the build is sharded and exchanged"""
        rm_all.sample(nv, inst.seed, inst.q_init, inst.q_goal, tries=False)
        rm_sh.build_async(radv)
        exchange_csr()
        adopt_csr()
        rm_sh.finish()
        rm_all.finish()
        k_table_full()

    def step_single():
        """sample -> neighbour pass -> conflict table without a host round trip in between"""
        rm_all.sample(nv, inst.seed, inst.q_init, inst.q_goal, tries=False)
        rm_all.build_async(radv_all)
        rm_all.edge_conflicts_async(n_cols, cols_d[0].data_ptr(), cols_d[1].data_ptr(),
                                    cols_d[2].data_ptr(), N, bits_d.data_ptr(), nnz_total + 4096, st)
        rm_all.finish()

    step_value = step_single if world == 1 else (step_strong if strong else step_weak)

    # ---- correctness of the sharded path before anything is timed
    strong_check = None
    if world > 1:
        k_table_full()                     # 1-GPU table of this rank's scenario (setup build)
        torch.cuda.synchronize()
        ref_bits = bits_d.clone()
        ref_rp, ref_col = row_ptr_h, col_h
        if strong:
            # keep a previous round's data out of the buffers the check reads
            n = step_strong()
            barrier()
            assert n == nnz_total, "stitched roadmaps differ in size"
            rp2, col2 = rm_all.csr()
            ok = bool(np.array_equal(rp2, ref_rp) and np.array_equal(col2, ref_col))
            if rank == 0:
                ok = ok and bool(torch.equal(gathered_table(), ref_bits))
            strong_check = {"roadmaps_identical": bool(max_over_ranks(0.0 if ok else 1.0) == 0.0),
                            "table_identical_on_rank0": ok if rank == 0 else None}
            assert strong_check["roadmaps_identical"], "sharded path differs from the 1-GPU result"
        else:
            step_weak()
            torch.cuda.synchronize()
            assert torch.equal(bits_d, ref_bits)
    if world == 1:     # the asynchronous pipeline gives the table of the synchronous calls
        k_table_full()
        torch.cuda.synchronize()
        ref_bits = bits_d.clone()
        bits_d.zero_()
        step_single()
        assert torch.equal(bits_d, ref_bits), "asynchronous step differs from the synchronous one"

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
        if world > 1 and strong:
            dist.barrier()                 # ranks start a step together (outside the event pair)
        ev[k][0].record(stream)
        step_value()
        ev[k][1].record(stream)
    barrier()
    t_end = time.perf_counter()
    launches = L.kernel_launch_count() - launches0
    per_step = [a.elapsed_time(b) for a, b in ev]
    ms_total = max_over_ranks(sum(per_step))
    clk = clocks.stop(t_begin, t_end)
    ms_step = ms_total / args.steps
    checks_step = gate + (pair_checks if (strong or world == 1) else world * pair_checks)
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

    def build_local():
        rm_all.sample(nv, inst.seed, inst.q_init, inst.q_goal, tries=False)
        rm_sh.build(radv)

    def build_full():
        rm_all.sample(nv, inst.seed, inst.q_init, inst.q_goal, tries=False)
        rm_all.build(radv_all)

    timeline = stage_timeline() if (world > 1 and strong) else None
    ms_build_local = timed(build_local, reps=5)
    ms_build = timed(build_full, reps=5)       # all N roadmaps on one GPU (leaves rm_all complete)
    ctx.stats(reset=True)
    ms_table = timed(k_table_full, reps=10)
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
        "bound": "fp64", "kernel": "k_cross_collide<2,or,fast-drop> (+ column sort/prep), all %d "
                                   "roadmap edges x %d columns on one GPU" % (nnz_total, n_cols),
        "achieved": achieved, "peak": FP64_UNFUSED_PEAK_TFLOPS, "unit": "TFLOP/s",
        "frac": achieved / FP64_UNFUSED_PEAK_TFLOPS, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
        "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` "
                          "capture of this kernel on this workload (profiles/r2_r_cross_ncu_summary.md); "
                          "algorithmic bytes per launch: 22.4e6 (396166 row records x 48 B + 3200 columns x 80 B + 8 B of table per row)",
        "peak_source": "un-fused DADD/DMUL issue peak measured with tools/fp64_peak.cu on this "
                       "pool (profiles/r1_fp64_peak.json); MEASURED_PEAKS.json has no FP64 figure "
                       "and bit parity forbids FMA in the reference arithmetic",
        "algorithmic_flops": {"pairs": n_pairs, "pairs_bound_tested_fp32": tested,
                              "fp32_per_tested_pair": FLOPS_BROAD32, "fp32_tflops": flops32 / (ms_table * 1e-3) / 1e12,
                              "pairs_filtered_fp64": survivors, "fp64_per_filtered_pair": FLOPS_FILTER,
                              "pairs_reference_arithmetic": exact, "fp64_per_exact_pair": FLOPS_EXACT,
                              "note": "work the kernel performs: `achieved` counts the FP64 part "
                                      "(verdict filter x 91 + reference arithmetic x 170, FMA = 2); "
                                      "the FP32 broad phase (20 per tested pair) is listed beside it; "
                                      "the reference itself spends 170 FP64 per pair"},
        "issue_slot_note": "the kernel is issue-bound, not pipe-bound: see profiles/r2_* (issue "
                           "active, pipe utilisation per pipe)",
        "reference_equivalent_tflops": n_pairs * FLOPS_EXACT / (ms_table * 1e-3) / 1e12,
        "pair_checks_per_s": n_pairs / (ms_table * 1e-3),
        "kernel_ms": {"conflict_table": ms_table, "roadmap_build": ms_build,
                      "roadmap_build_this_rank_shard": ms_build_local},
    }

    # ---- e2e: host buffers through the C ABI (host -> device inputs, device -> host results)
    e2e_roadmaps = rm_sh if world > 1 else rm_all
    n_ret = max(n_local, 1) if world > 1 else N
    rows_sh = n_ret * nv
    nodes_out_h = torch.empty((n_ret, nv, d), dtype=torch.float64).pin_memory()
    rowptr_out_h = torch.empty(rows_sh + 1, dtype=torch.int32).pin_memory()
    colidx_out_h = torch.empty(max(nnz_total, 1), dtype=torch.int32).pin_memory()
    cols_h_all = torch.stack(cols_h).pin_memory()            # (agent, v_from, v_to) ids in one pinned block
    cols_up_all = torch.empty_like(cols_h_all, device=dev)
    cols_up = [cols_up_all[k] for k in range(3)]

    e2e_mode = {"direct": True, "early_read": not os.environ.get("BENCH_E2E_LATE_READ")}
    e2e_trace = [] if os.environ.get("BENCH_E2E_TRACE") else None   # host seconds: (step, read-back)

    def step_e2e():
        if world == 1:
            # host buffers in (init / goal states, path ids), host buffers out (roadmaps: nodes + CSR,
            # conflict table), through the asynchronous pipeline of the C ABI: one host wait
            cols_up_all.copy_(cols_h_all, non_blocking=True)         # H2D: path ids (one copy)
            rm_all.sample(nv, inst.seed, inst.q_init, inst.q_goal, tries=False)   # host init / goal
            rm_all.build_async(radv_all)
            nnz = C.c_int64(0)
            if e2e_mode["early_read"]:   # D2H of the roadmaps beside the table kernel
                L.check(lib.mrmp_roadmaps_read_async(rm_all._h, nodes_out_h.data_ptr(),
                                                     rowptr_out_h.data_ptr(), colidx_out_h.data_ptr(),
                                                     colidx_out_h.numel(), C.byref(nnz)))
            out_ptr = bits_h.data_ptr() if e2e_mode["direct"] else bits_d.data_ptr()
            rm_all.edge_conflicts_async(n_cols, cols_up[0].data_ptr(), cols_up[1].data_ptr(),
                                        cols_up[2].data_ptr(), N, out_ptr, nnz_total, st)
            if not e2e_mode["direct"]:
                # D2H of the table queued behind the kernel on the same stream: the step's one host
                # wait (finish) covers it, no second round trip
                bits_h.copy_(bits_d, non_blocking=True)
            rm_all.finish()
            if not e2e_mode["early_read"]:
                L.check(lib.mrmp_roadmaps_read_async(rm_all._h, nodes_out_h.data_ptr(),
                                                     rowptr_out_h.data_ptr(), colidx_out_h.data_ptr(),
                                                     colidx_out_h.numel(), C.byref(nnz)))   # D2H: roadmaps
            L.check(lib.mrmp_roadmaps_read_wait(rm_all._h))
            return
        # multi-GPU: path ids from pinned host memory on every rank, this rank's roadmaps back to
        # its host, the gathered table back to rank 0's host
        tr = e2e_trace
        t_a = time.perf_counter()
        cols_up_all.copy_(cols_h_all, non_blocking=True)
        if strong:
            step_strong(cols_up)
        else:
            step_weak()
        t_b = time.perf_counter()
        nnz = C.c_int64(0)
        L.check(lib.mrmp_roadmaps_read_async(rm_sh._h, nodes_out_h.data_ptr(), rowptr_out_h.data_ptr(),
                                             colidx_out_h.data_ptr(), colidx_out_h.numel(), C.byref(nnz)))
        if strong:
            if rank == 0:
                bits_h.copy_(gathered_table(), non_blocking=True)
        else:
            bits_h.copy_(bits_d, non_blocking=True)
        L.check(lib.mrmp_roadmaps_read_wait(rm_sh._h))
        torch.cuda.current_stream().synchronize()
        if tr is not None:
            tr.append((t_b - t_a, time.perf_counter() - t_b) + tuple(mg.get("host_s", (0.0, 0.0))))

    for _ in range(2):
        step_e2e()
    barrier()
    if world == 1:
        # the table either goes straight into the caller's pinned buffer (kernel stores over PCIe)
        # or into device memory followed by one copy: take the faster, check both give the table
        def wall(n=5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(n):
                step_e2e()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / n
        t_direct = wall()
        ref_h = bits_h.clone()
        e2e_mode["direct"] = False
        step_e2e()
        t_staged = wall()
        assert torch.equal(bits_h, ref_h), "e2e: direct and staged tables differ"
        e2e_mode["direct"] = t_direct <= t_staged
        e2e_variants = {"table_direct_to_pinned_host_ms": t_direct * 1e3, "table_staged_copy_ms": t_staged * 1e3}
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    if world == 1:   # what travelled to the host is the roadmap set the device holds
        rp_chk, col_chk = rm_all.csr()
        assert np.array_equal(rowptr_out_h.numpy(), rp_chk), "e2e: row_ptr read back differs"
        assert np.array_equal(colidx_out_h.numpy()[: col_chk.size], col_chk), "e2e: col_idx read back differs"
        assert np.array_equal(nodes_out_h.numpy(), rm_all.nodes()), "e2e: nodes read back differ"
    if e2e_trace:
        tail = e2e_trace[-e2e_steps:]
        mean = [1e6 * sum(t[k] for t in tail) / len(tail) for k in range(4)]
        sys.stderr.write("e2e trace rank %d: step %.1f us (one-call form: enqueue %.1f, waits %.1f), "
                         "read-back %.1f us (means of %d)\n" % (rank, mean[0], mean[2], mean[3], mean[1], len(tail)))
    h2d = 2 * N * d * 8 + 3 * n_cols * 4
    if world == 1:
        d2h = N * nv * d * 8 + (N * nv + 1) * 4 + nnz_total * 4 + nnz_total * wpr * 4
    else:     # rank 0, the busiest: its roadmaps + the whole table (strong) / its table (weak)
        d2h = n_ret * nv * d * 8 + (rows_sh + 1) * 4 + (nnz_total // world) * 4 + nnz_total * wpr * 4
    e2e = {"value": checks_step / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3,
           "note": "per rank; rank 0 also reads the gathered table" if world > 1 else
                   "asynchronous pipeline, one host wait; roadmaps read back beside the table kernel; "
                   + json.dumps(e2e_variants)}

    # ---- explicit-list kernels (HBM side of the path), SSSP expansion, arm22: rank 0 only
    lists = sssp_x = arm = sssp_solve = p3d = None
    if rank == 0 and world == 1 and not args.no_extras:
        build_full()
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
        # collide(Q, q_to, i) (point.jl:27-33, the SSSP successor form): n_cand candidates against
        # the N-1 static others
        n_cand = 1 << 20
        Qs = torch.from_numpy(np.ascontiguousarray(nodes_h[:, 0])).to(dev)
        cand = torch.rand((n_cand, 2), dtype=torch.float64, device=dev)
        outc = torch.empty(n_cand, dtype=torch.uint8, device=dev)
        ms_mv = timed(lambda: L.check(lib.mrmp_collide_config_moves_dev(
            ctx._h, 0, Qs.data_ptr(), n_cand, cand.data_ptr(), outc.data_ptr(), st)))
        flops_con = 6 + 10 * (15 * 2 + 4)
        flops_mv = (N - 1) * (15 * 2 + 4)
        lists = {"n": n,
                 "connect_edges": {"ms": ms_con, "checks_per_s": n / ms_con * 1e3,
                                   "hbm_gbs": n * BYTES_CONNECT / ms_con / 1e6,
                                   "hbm_frac": n * BYTES_CONNECT / ms_con / 1e6 / pk["hbm_gbs"],
                                   "fp64_tflops": n * flops_con / ms_con / 1e9,
                                   "fp64_frac": n * flops_con / ms_con / 1e9 / FP64_UNFUSED_PEAK_TFLOPS},
                 "collide_edges_indexed": {"ms": ms_idx, "checks_per_s": n / ms_idx * 1e3,
                                           "fp64_tflops": n * FLOPS_EXACT / ms_idx / 1e9,
                                           "fp64_frac": n * FLOPS_EXACT / ms_idx / 1e9 / FP64_UNFUSED_PEAK_TFLOPS},
                 "collide_config_moves": {"n_candidates": n_cand, "ms": ms_mv,
                                          "candidates_per_s": n_cand / ms_mv * 1e3,
                                          "segment_point_tests_per_s": n_cand * (N - 1) / ms_mv * 1e3,
                                          "fp64_tflops": n_cand * flops_mv / ms_mv / 1e9,
                                          "fp64_frac": n_cand * flops_mv / ms_mv / 1e9 / FP64_UNFUSED_PEAK_TFLOPS},
                 "peak_hbm_gbs": pk["hbm_gbs"], "peak_source": pk_kind}
        del con_d, cl_d, out, cand, outc
        sssp_x = extra_sssp_expansion(inst, nodes_h[:, :500] if nv >= 500 else nodes_h, local)
        arm = extra_arm22(local, dev, st, timed)
        sssp_solve = extra_sssp_solve(local)
        p3d = extra_point3d(local, dev, st, timed)

    # ---- cpu_baseline (rank 0, N=1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        arm_cpu = CpuArm(inst)
        arm_cpu.step(ca, vf, vt, CPU_ROW_STRIDE)      # warm-up
        best = None
        for _ in range(3):
            b, t, n_rows, *_ = arm_cpu.step(ca, vf, vt, CPU_ROW_STRIDE)
            if best is None or b + t < best[0] + best[1]:
                best = (b, t, n_rows)
        cchecks = gate + best[2] * T_STEPS * (N - 1)
        # one host thread (a single reference solver call is single-threaded)
        arm1 = CpuArm(inst, threads=1)
        b1, t1, n1, *_ = arm1.step(ca, vf, vt, CPU_ROW_STRIDE * 8)
        cpu = {"value": cchecks / (best[0] + best[1]), "unit": UNIT, "cores": arm_cpu.threads,
               "kind": "port",
               "sample": "full N=50 roadmap build + conflict table on every %dth roadmap edge "
                         "(%d rows x %d columns), best of 3; C oracle, OpenMP %d threads"
                         % (CPU_ROW_STRIDE, best[2], n_cols, arm_cpu.threads),
               "roadmap_build_ms": 1e3 * best[0],
               "collide_pair_checks_per_s": best[2] * T_STEPS * (N - 1) / best[1],
               "one_thread": {"value": (gate + n1 * T_STEPS * (N - 1)) / (b1 + t1),
                              "roadmap_build_ms": 1e3 * b1,
                              "collide_pair_checks_per_s": n1 * T_STEPS * (N - 1) / t1,
                              "sample": "every %dth roadmap edge" % (CPU_ROW_STRIDE * 8)}}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(world, strong, nv),
            "checks_per_step": {"connect": gate, "collide": checks_step - gate},
            "roadmap_build_ms": ms_build, "roadmap_nnz": nnz_total,
            "multi_gpu": None if world == 1 else {
                "mode": "strong: one scenario, agents' roadmaps and table rows sharded over the ranks"
                        if strong else "weak: one scenario per rank, sharded build",
                "exchange": mg["exchange"], "csr_slot_bytes": 4 * (cap * nv + mg["cap_nnz"]),
                "table_bytes_to_rank0": int(nnz_total * wpr * 4) if strong else 0,
                "check": strong_check, "ms_per_step_each": per_step[:8], "timeline": timeline,
                "host_calls_per_step": ("one call per stage (--separate-calls)" if args.separate_calls or
                                        not (strong and mg["p2p"]) else
                                        "1 (mrmp_roadmaps_strong_step) + 2 finish; the stage timeline is "
                                        "taken with one call per stage")},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu,
            "explicit_lists": lists, "sssp_expansion": sssp_x, "sssp_solve": sssp_solve, "arm22": arm, "point3d": p3d,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        barrier()
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
    t_lib = t_waitk = 0.0
    tot_us, wait_us = C.c_double(0), C.c_double(0)
    for a, v, ids, old, qr in calls[1:]:
        Q = np.stack([tabs[j][ids[j]] for j in range(N)])
        t0 = time.perf_counter()
        call(a, v, ids, old, qr)
        t1 = time.perf_counter()
        lib.mrmp_sssp_last_call_us(dev_rm._h, C.byref(tot_us), C.byref(wait_us))
        t_lib += tot_us.value
        t_waitk += wait_us.value
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
            "inside_library_ms_per_expansion": 1e-3 * t_lib / n,
            "of_which_waiting_for_the_kernel_ms": 1e-3 * t_waitk / n,
            "timed": "mrmp_sssp_expand through the C ABI (host buffers in, results out)",
            "accepted_vertices": n_new, "mismatching_calls": mism}


def extra_sssp_solve(device):
    """End-to-end SSSP (sssp.jl:49-232) on point2d_many instances with the solver parameters of
    point2d_many.yaml:22-24: the GPU-backed host mirror (sssp_b200/solvers.py: one fused expansion
    call per popped node) against the per-call restatement over the C oracle (oracle/sssp_ref.py).
    Instances and the oracle's solutions are the committed golden fixtures; the solutions must be
    identical.  The N=50 oracle run (34 s in the build container) is not repeated here."""
    import sssp_b200 as S
    from sssp_b200 import solvers
    from oracle import oracle as O
    from oracle import sssp_ref as R
    out = {}
    for name, live in (("sssp_point2d_many_n30_seed0", True), ("sssp_point2d_many_n50_seed0", False)):
        path = os.path.join(ROOT, "tests", "golden", name + ".json")
        if not os.path.exists(path):
            continue
        with open(path) as f:
            fx = json.load(f)
        inst = fx["instance"]
        P = S.StatePoint2D
        init, goal = [P(*q) for q in inst["config_init"]], [P(*q) for q in inst["config_goal"]]
        obs = np.array(inst["obstacles"])
        connect = S.gen_connect(init[0], obs, inst["rads"], device=device)
        collide = S.gen_collide(init[0], inst["rads"], device=device)
        st = {}
        t0 = time.perf_counter()
        sol, _ = solvers.SSSP(init, goal, connect, collide, S.gen_check_goal(goal), TIME_LIMIT=None,
                              seed=fx["seed"], stats=st, **fx["params"])
        t_gpu = time.perf_counter() - t0
        row = {"n_agents": len(init), "expansions": st.get("expansions"),
               "solution_steps": None if sol is None else len(sol),
               "gpu_backed_s": t_gpu, "gpu_ms_per_expansion_incl_host_search": 1e3 * t_gpu / max(1, st.get("expansions", 1)),
               "identical_to_oracle_solution": st.get("solution_ids") == fx["solution"]}
        if live:
            orc = O.Oracle(O.POINT2D, inst["rads"], obs)
            t0 = time.perf_counter()
            osol, _, ost = R.SSSP(orc, inst["config_init"], inst["config_goal"], seed=fx["seed"], **fx["params"])
            row["oracle_per_call_s"] = time.perf_counter() - t0
            row["identical_to_oracle_solution"] = row["identical_to_oracle_solution"] and osol == fx["solution"]
        out[name] = row
    return out


def extra_point3d(device, dev, st, timed):
    """BASELINE configs[2]: StatePoint3D N=20 with sphere obstacles: PRMs(200, rad 0.3) + the
    conflict table (all roadmap edges x the other agents' path edges, T=64) through the same
    kernels (3-D instantiations), with the kernel counters and a sampled parity check."""
    import torch
    import sssp_b200 as S
    from sssp_b200 import lib as L, workload as W
    from oracle import oracle as O
    lib = L.load()
    inst = W.point3d_n20()
    N, nv = inst.rads.size, inst.nv
    ctx = S.Context(inst.model, inst.rads, inst.obstacles, device=device)
    ctx.set_stream(st)
    rm = S.Roadmaps(ctx, 0, N, nv)

    def build():
        rm.sample(nv, inst.seed, inst.q_init, inst.q_goal, tries=False)
        return rm.build(inst.rad)

    nnz = build()
    ms_build = timed(build, reps=5)
    nodes = rm.nodes()
    rp, col = rm.csr()
    _, ca, vf, vt = W.random_walk_paths(inst, rp, col, T_STEPS, seed=77)
    cols = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (ca, vf, vt)]
    wpr = (T_STEPS + 31) // 32
    bits = torch.zeros((nnz, wpr), dtype=torch.int32, device=dev)
    table = lambda: L.check(lib.mrmp_roadmaps_edge_conflicts_dev(
        rm._h, None, ca.size, cols[0].data_ptr(), cols[1].data_ptr(), cols[2].data_ptr(), N,
        bits.data_ptr(), st))
    ctx.stats(reset=True)
    ms = timed(table, reps=10)
    s4 = ctx.stats()
    exact, tested, filt = int(s4[0]) / 11.0, int(s4[2]) / 11.0, int(s4[3]) / 11.0
    flops = filt * 150 + exact * 243            # 3-D: verdict filter ~150 (FMA = 2), reference test 243
    # parity on every 97th row against the oracle
    rows = np.repeat(np.arange(N * nv), np.diff(rp))
    sel = slice(0, None, 97)
    ra = (rows // nv).astype(np.int32)
    orc = O.Oracle(O.POINT3D, inst.rads, inst.obstacles)
    obits = orc.collide_cross(ra[sel], nodes[ra, rows % nv][sel], nodes[ra, col][sel], ca,
                              nodes[ca, vf], nodes[ca, vt], N, nthreads=O.host_cores())
    g = bits.cpu().numpy().view(np.uint32)[sel]
    return {"n_agents": int(N), "num_vertices": nv, "rad": inst.rad, "roadmap_nnz": int(nnz),
            "roadmap_build_ms": ms_build, "table_ms": ms, "pair_checks_per_s": nnz * ca.size / ms * 1e3,
            "pairs": int(nnz) * int(ca.size), "pairs_bound_tested_fp32": tested,
            "pairs_filtered_fp64": filt, "pairs_reference_arithmetic": exact,
            "roofline": {"bound": "fp64", "achieved": flops / ms / 1e9, "peak": FP64_UNFUSED_PEAK_TFLOPS,
                         "unit": "TFLOP/s", "frac": flops / ms / 1e9 / FP64_UNFUSED_PEAK_TFLOPS},
            "mismatching_rows_on_sample": int((g != obits).any(axis=1).sum()), "sample_rows": int(g.shape[0])}


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
    run_col = lambda: L.check(lib.mrmp_collide_pairs_dev(
        ctx._h, n_col, *[t.data_ptr() for t in d_[:6]], out.data_ptr(), st))
    ms_col = timed(run_col, reps=3)
    ctx.stats8(reset=True)
    run_col()
    s8 = ctx.stats8()
    bounds, exact_links = int(s8[4]), int(s8[1])
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
                              "hit_rate": float(g_col.mean()),
                              # as-executed FP64 work of k_arm_collide<ListSrc> (kernel counters): 7 per
                              # link-pair midpoint bound (2 sub, 2 mul, add, compare = 6, + 1 select),
                              # 170 per link pair that ran the reference segment-segment test, against the
                              # un-fused DADD/DMUL peak; trig of the posture tables not counted
                              "roofline": {"bound": "fp64", "link_pair_bounds": bounds, "exact_link_pairs": exact_links,
                                           "achieved": (bounds * 7 + exact_links * FLOPS_EXACT) / ms_col / 1e9,
                                           "peak": FP64_UNFUSED_PEAK_TFLOPS, "unit": "TFLOP/s",
                                           "frac": (bounds * 7 + exact_links * FLOPS_EXACT) / ms_col / 1e9 / FP64_UNFUSED_PEAK_TFLOPS,
                                           "reference_equivalent_link_pair_tests": None}},
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
