"""Scratch timing of the arm22 6-arity collide kernel alone."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sssp_b200 as S
from sssp_b200 import lib as L
lib = L.load(); dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(22); N, M = 10, 3
rads = rng.uniform(0.15, 0.25, N); base = rng.uniform(0.3, 0.7, (N, 2))
obs = np.concatenate([rng.random((M, 2)), rng.uniform(0.05, 0.1, (M, 1))], axis=1)
ctx = S.Context(L.MODEL_ARM22, rads, obs, base_pos=base); ctx.set_stream(st)
for n in (512, 4096, 65536, 524288):
    ai = rng.integers(0, N, n).astype(np.int32); aj = ((ai + 1 + rng.integers(0, N - 1, n)) % N).astype(np.int32)
    q = [rng.uniform(0, 2 * np.pi, (n, 2)) for _ in range(4)]
    d = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (ai, aj, *q)]
    out = torch.empty(n, dtype=torch.uint8, device=dev)
    f = lambda: L.check(lib.mrmp_collide_pairs_dev(ctx._h, n, *[t.data_ptr() for t in d], out.data_ptr(), st))
    for _ in range(2): f()
    torch.cuda.synchronize(); best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    print(json.dumps({"n": n, "ms": round(best, 4), "checks_per_s": n / best * 1e3, "hit": float(out.float().mean())}))
