"""Scratch timing of the explicit-list kernels (device-resident inputs, CUDA events)."""
import ctypes as C
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sssp_b200 as S
from sssp_b200 import lib as L
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import make_point_instance, random_edges

dev = torch.device("cuda:0")
lib = L.load()
N, M = 50, 10
rads, obs = make_point_instance(0, N, M, seed=0)
ctx = S.Context(0, rads, obs)
res = {}
for logn in (20, 24, 26):
    n = 1 << logn
    g = torch.Generator(device=dev).manual_seed(1234)
    ag = torch.randint(0, N, (n,), device=dev, dtype=torch.int32, generator=g)
    aj = (ag + 1 + torch.randint(0, N - 1, (n,), device=dev, dtype=torch.int32, generator=g)) % N
    qf = torch.rand((n, 2), device=dev, dtype=torch.float64, generator=g)
    qt = qf + (torch.rand((n, 2), device=dev, dtype=torch.float64, generator=g) - 0.5) * 0.14
    c = torch.rand((n, 2), device=dev, dtype=torch.float64, generator=g)
    e = c + (torch.rand((n, 2), device=dev, dtype=torch.float64, generator=g) - 0.5) * 0.14
    out = torch.empty(n, device=dev, dtype=torch.uint8)
    st = torch.cuda.current_stream().cuda_stream

    def t(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms = t(lambda: L.check(lib.mrmp_connect_edges_dev(ctx._h, n, ag.data_ptr(), qf.data_ptr(),
                                                      qt.data_ptr(), out.data_ptr(), st)))
    res[f"connect_edges_2^{logn}"] = dict(ms=ms, gchecks_s=n / ms / 1e6, free=float(out.float().mean()))
    ms = t(lambda: L.check(lib.mrmp_collide_pairs_dev(ctx._h, n, ag.data_ptr(), aj.data_ptr(),
                                                      qf.data_ptr(), qt.data_ptr(), c.data_ptr(),
                                                      e.data_ptr(), out.data_ptr(), st)))
    res[f"collide_pairs_2^{logn}"] = dict(ms=ms, gchecks_s=n / ms / 1e6, hit=float(out.float().mean()))
# roadmap build N=50, nv=500
import time
rm = S.Roadmaps(ctx, 0, N, 500)
rng = np.random.default_rng(0)
init, goal = rng.random((N, 2)) * 0.8 + 0.1, rng.random((N, 2)) * 0.8 + 0.1
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rm.sample(500, 1, init, goal)
    t1 = time.perf_counter()
    nnz = rm.build(0.1)
    t2 = time.perf_counter()
    rp, col = rm.csr()
    t3 = time.perf_counter()
res["prm_N50_nv500"] = dict(sample_ms=(t1 - t0) * 1e3, build_ms=(t2 - t1) * 1e3, csr_d2h_ms=(t3 - t2) * 1e3, nnz=int(nnz))
print(json.dumps(res, indent=1))
