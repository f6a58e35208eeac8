"""BASELINE.json configs[4]: synthetic throughput sweep, 10^4 .. 10^8 edge checks per launch.

point2d: N=50 agents, M=10 obstacles, radii as point2d_many.yaml; half of each batch are
connect(q_from,q_to,i) edges (length <= 0.1), half collide 6-arity pairs of nearby edges.
arm22: N=10, M=3 (capped at 10^6 checks: one arm collide is ~10^4 primitive tests).
Inputs are generated on the device (torch, seed 1234) and stay resident; timing = CUDA events
around the two launches, best of 5 after 2 warm-ups.  For n <= 10^7 the same batch also goes
through the host-buffer C ABI (e2e: pinned host arrays in, verdict bytes out).  Rank r of a
multi-GPU run processes its own n checks (weak scaling, no collective on the data path).

  python tools/sweep.py [--max-log 8] [--arm]      (torchrun for several GPUs)
"""
import argparse, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import sssp_b200 as S
from sssp_b200 import lib as L, workload as W

ap = argparse.ArgumentParser()
ap.add_argument("--max-log", type=int, default=8)
ap.add_argument("--arm", action="store_true")
ap.add_argument("--cpu", action="store_true", help="also time the C oracle on a bounded sample")
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
lib = L.load()
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device=dev).manual_seed(1234 + rank)

def rand(*shape): return torch.rand(shape, device=dev, dtype=torch.float64, generator=g)
def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

if args.arm:
    rng = np.random.default_rng(22); N, M = 10, 3
    rads = rng.uniform(0.15, 0.25, N); base = rng.uniform(0.3, 0.7, (N, 2))
    obs = np.concatenate([rng.random((M, 2)), rng.uniform(0.05, 0.1, (M, 1))], axis=1)
    ctx = S.Context(L.MODEL_ARM22, rads, obs, base_pos=base, device=local); scale = 2 * np.pi
    logs = range(3, min(args.max_log, 6) + 1)
else:
    inst = W.point2d_many(50, seed=0); N = 50
    ctx = S.Context(inst.model, inst.rads, inst.obstacles, device=local); scale = 1.0
    logs = range(4, args.max_log + 1)
ctx.set_stream(st)
if args.cpu and rank == 0:
    from oracle import oracle as O
    orc = O.Oracle(O.ARM22, rads, obs, base=base) if args.arm else O.Oracle(O.POINT2D, inst.rads, inst.obstacles)
for lg in logs:
    n = 10 ** lg; h = n // 2
    ag = torch.randint(0, N, (h,), device=dev, dtype=torch.int32, generator=g)
    aj = ((ag + 1 + torch.randint(0, N - 1, (h,), device=dev, dtype=torch.int32, generator=g)) % N).to(torch.int32)
    if args.arm:
        qf, qt, cf, ct = (rand(h, 2) * scale for _ in range(4))
    else:
        qf = rand(h, 2); qt = qf + (rand(h, 2) - 0.5) * 0.14
        cf = qf + (rand(h, 2) - 0.5) * 0.2; ct = cf + (rand(h, 2) - 0.5) * 0.14
    out1 = torch.empty(h, dtype=torch.uint8, device=dev); out2 = torch.empty(h, dtype=torch.uint8, device=dev)
    def run():
        L.check(lib.mrmp_connect_edges_dev(ctx._h, h, ag.data_ptr(), qf.data_ptr(), qt.data_ptr(), out1.data_ptr(), st))
        L.check(lib.mrmp_collide_pairs_dev(ctx._h, h, ag.data_ptr(), aj.data_ptr(), qf.data_ptr(), qt.data_ptr(),
                                           cf.data_ptr(), ct.data_ptr(), out2.data_ptr(), st))
    if world > 1: dist.barrier()
    ms = timed(run)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    rec = {"model": "arm22" if args.arm else "point2d", "n_gpus": world, "checks_per_launch_per_gpu": n,
           "device_ms": ms, "checks_per_s": world * n / ms * 1e3,
           "connect_free_rate": float(out1.float().mean().item()), "collide_hit_rate": float(out2.float().mean().item())}
    if lg <= 7 and world == 1:
        hb = [x.cpu().pin_memory().numpy() for x in (ag, aj, qf, qt, cf, ct)]   # mrmp_host_alloc-style pinned inputs
        ctx.set_stream(None)
        t0 = time.perf_counter()
        o1 = ctx.connect_edges(hb[0], hb[2], hb[3]); o2 = ctx.collide_pairs(hb[0], hb[1], hb[2], hb[3], hb[4], hb[5])
        t1 = time.perf_counter()
        o1 = ctx.connect_edges(hb[0], hb[2], hb[3]); o2 = ctx.collide_pairs(hb[0], hb[1], hb[2], hb[3], hb[4], hb[5])
        t2 = time.perf_counter()
        ctx.set_stream(st)
        rec["e2e_ms"] = (t2 - t1) * 1e3; rec["e2e_checks_per_s"] = n / (t2 - t1)
        assert np.array_equal(o1, out1.cpu().numpy()) and np.array_equal(o2, out2.cpu().numpy())
        if args.cpu:
            m = min(h, 20000 if args.arm else 2_000_000)
            t0 = time.perf_counter()
            c1 = orc.connect_edges(hb[0][:m], hb[2][:m], hb[3][:m], nthreads=0)
            c2 = orc.collide_pairs(hb[0][:m], hb[1][:m], hb[2][:m], hb[3][:m], hb[4][:m], hb[5][:m], nthreads=0)
            t1 = time.perf_counter()
            rec["cpu_checks_per_s"] = 2 * m / (t1 - t0); rec["cpu_threads"] = O.max_threads()
            rec["mismatches_on_cpu_sample"] = int((c1 != o1[:m]).sum() + (c2 != o2[:m]).sum())
    if rank == 0: print(json.dumps(rec), flush=True)
    del ag, aj, qf, qt, cf, ct, out1, out2
if world > 1: dist.destroy_process_group()
