"""Scratch timing of the explicit-list connect_edges kernel (2^24 device-resident checks)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import sssp_b200 as S
from sssp_b200 import lib as L
from helpers import make_point_instance
dev = torch.device("cuda:0"); lib = L.load()
for M in (10, 16, 3):
    rads, obs = make_point_instance(0, 50, M, seed=0)
    ctx = S.Context(0, rads, obs)
    n = 1 << 24
    g = torch.Generator(device=dev).manual_seed(1234)
    ag = torch.randint(0, 50, (n,), device=dev, dtype=torch.int32, generator=g)
    qf = torch.rand((n, 2), device=dev, dtype=torch.float64, generator=g)
    qt = qf + (torch.rand((n, 2), device=dev, dtype=torch.float64, generator=g) - 0.5) * 0.14
    out = torch.empty(n, device=dev, dtype=torch.uint8)
    st = torch.cuda.current_stream().cuda_stream
    f = lambda: L.check(lib.mrmp_connect_edges_dev(ctx._h, n, ag.data_ptr(), qf.data_ptr(), qt.data_ptr(), out.data_ptr(), st))
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    print("M", M, "ms", e0.elapsed_time(e1) / 10, "free", float(out.float().mean()))
