"""Scratch timing of the roadmap edge-conflict table (cross kernel)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sssp_b200 as S
from sssp_b200 import lib as L, workload as W
lib = L.load()
dev = torch.device("cuda:0")
inst = W.point2d_many(50, seed=0)
N, nv, T = 50, 500, 64
ctx = S.Context(inst.model, inst.rads, inst.obstacles)
st = torch.cuda.current_stream()
ctx.set_stream(st.cuda_stream)
rm = S.Roadmaps(ctx, 0, N, nv)
rm.sample(nv, 0, inst.q_init, inst.q_goal)
nnz = rm.build(np.full(N, 0.1))
row_ptr, col = rm.csr()
rng = np.random.default_rng(1)
paths = np.zeros((N, T + 1), np.int32)
for j in range(N):
    v = 0
    for t in range(T):
        r = j * nv + v
        deg = row_ptr[r + 1] - row_ptr[r]
        v = int(col[row_ptr[r] + rng.integers(0, deg)]) if deg else v
        paths[j, t + 1] = v
ca = np.tile(np.arange(N, dtype=np.int32), T)
vf = paths[:, :-1].T.reshape(-1).astype(np.int32)
vt = paths[:, 1:].T.reshape(-1).astype(np.int32)
ca_d, vf_d, vt_d = (torch.from_numpy(x).to(dev) for x in (ca, vf, vt))
res = {}
for gs in (N, 0):
    wpr = ((len(ca) // gs if gs else len(ca)) + 31) // 32
    out = torch.zeros((nnz, wpr), dtype=torch.int32, device=dev)
    def run():
        L.check(lib.mrmp_roadmaps_edge_conflicts_dev(rm._h, None, len(ca), ca_d.data_ptr(), vf_d.data_ptr(),
                                                     vt_d.data_ptr(), gs, out.data_ptr(), st.cuda_stream))
    for _ in range(3): run()
    torch.cuda.synchronize()
    ctx.stats(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); 
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    stt = ctx.stats()
    surv = int(stt[0]) / 10
    tested = int(stt[2]) / 10
    pairs = nnz * len(ca)
    res["gs=%d" % gs] = dict(ms=ms, pairs=pairs, gpairs_s=pairs / ms / 1e6, survivors=surv, surv_frac=surv / pairs, tested=tested, tested_frac=tested / pairs,
                             hits=int(torch.count_nonzero(out).item()))
    t0 = time.perf_counter(); g = rm.edge_conflicts(ca, vf, vt, gs); t1 = time.perf_counter()
    res["gs=%d" % gs]["e2e_ms"] = (t1 - t0) * 1e3
print(json.dumps(res, indent=1))
