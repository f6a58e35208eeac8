"""Scratch: conflict-table time of a rank's share (agents [0, n)) of the bench instance on one GPU."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sssp_b200 as S
from sssp_b200 import lib as L, workload as W
lib = L.load()
N, nv, T = 50, 500, 64
inst = W.point2d_many(N, seed=0, nv=nv, rad=0.1)
ctx = S.Context(inst.model, inst.rads, inst.obstacles)
stream = torch.cuda.current_stream(); st = stream.cuda_stream
ctx.set_stream(st)
full = S.Roadmaps(ctx, 0, N, nv)
full.sample(nv, inst.seed, inst.q_init, inst.q_goal)
nnz = full.build(inst.rad)
rp, col = full.csr()
_, ca, vf, vt = W.random_walk_paths(inst, rp, col, T, seed=4321)
cols = [torch.from_numpy(x).cuda() for x in (ca, vf, vt)]
out = torch.zeros((nnz, 2), dtype=torch.int32, device="cuda")
res = {}
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps): fn()
    b.record(stream); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
nodes_ptr = full.device_ptrs()["nodes"]
for n_ag in (50, 25, 13, 7):
    sh = S.Roadmaps(ctx, 0, n_ag, nv)
    sh.bind_nodes(nv, nodes_ptr)
    nn = sh.build(inst.rad)
    def table():
        sh.edge_conflicts_dev(ca.size, cols[0].data_ptr(), cols[1].data_ptr(), cols[2].data_ptr(), N, out.data_ptr(), st, cols=full)
    ctx.stats(reset=True)
    ms = timed(table)
    stt = ctx.stats()
    def build():
        sh.build(inst.rad)
    msb = timed(build)
    def both():
        sh.build(inst.rad); table()
    msbt = timed(both)
    res[n_ag] = dict(rows=nn, table_ms=ms, build_ms=msb, build_table_ms=msbt, tested=int(stt[2]) / 11, filtered=int(stt[3]) / 11)
print(json.dumps(dict(split=os.environ.get("MRMP_CROSS_SPLIT", "auto"), res=res)))
