"""Step-by-step timing of the arm roadmap / cross path (debug aid)."""
import sys, time, math
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import sssp_b200 as S
from sssp_b200 import lib as L
from oracle import oracle as O
from helpers import make_arm_instance

def T(msg, t0):
    print("%-40s %.3f s" % (msg, time.time() - t0), flush=True)

N, M, nv, rad = 6, 2, 100, 0.3
rads, obs, base = make_arm_instance(N, M, 17)
ctx = S.Context(L.MODEL_ARM22, rads, obs, base_pos=base)
orc = O.Oracle(O.ARM22, rads, obs, base=base)
rng = np.random.default_rng(26)
t = time.time()
init = np.stack([orc.sample_free(a, 900, 1)[0][0] for a in range(N)])
goal = np.stack([orc.sample_free(a, 901, 1)[0][0] for a in range(N)])
T("oracle init/goal", t)
rm = S.Roadmaps(ctx, 0, N, nv)
t = time.time(); tries = rm.sample(nv, 77, init, goal); nodes = rm.nodes(); T("gpu sample", t)
t = time.time(); exp = [orc.sample_free(a, 77, nv - 2) for a in range(N)]; T("oracle sample", t)
print("sample equal", all(np.array_equal(nodes[a, 2:], exp[a][0]) for a in range(N)), tries, flush=True)
t = time.time(); nnz = rm.build(rad); T("gpu build nnz=%d" % nnz, t)
row_ptr, col = rm.csr()
t = time.time(); orp, ocol, onnz = orc.prm_build(0, nodes, rad, nthreads=0); T("oracle build nnz=%d" % onnz.sum(), t)
rows = np.repeat(np.arange(N * nv), np.diff(row_ptr))
n = 1500
ei = rng.integers(0, nnz, n); ej = rng.integers(0, nnz, n)
ai, aj = (rows[ei] // nv).astype(np.int32), (rows[ej] // nv).astype(np.int32)
keep = ai != aj
ai, aj, ei, ej = ai[keep], aj[keep], ei[keep], ej[keep]
args = (ai, aj, (rows[ei] % nv).astype(np.int32), col[ei], (rows[ej] % nv).astype(np.int32), col[ej])
t = time.time(); g = rm.collide_edges_indexed(*args); T("gpu indexed", t)
t = time.time(); o = orc.collide_edges_indexed(*args, nodes, nv, nthreads=0); T("oracle indexed", t)
print("indexed equal", np.array_equal(g, o), o.sum(), flush=True)
ra, ca = ai[:40], aj[:96]
rf, rt = nodes[ra, args[2][:40]], nodes[ra, args[3][:40]]
cf, ct = nodes[ca, args[4][:96]], nodes[ca, args[5][:96]]
for gs in (0, 8):
    t = time.time(); g = ctx.collide_cross(ra, rf, rt, ca, cf, ct, gs); T("gpu cross gs=%d" % gs, t)
    t = time.time(); o = orc.collide_cross(ra, rf, rt, ca, cf, ct, gs, nthreads=0); T("oracle cross", t)
    print("cross equal", np.array_equal(g, o), flush=True)
t = time.time(); g = rm.edge_conflicts(ca, args[4][:96], args[5][:96], group_size=8); T("gpu edge_conflicts", t)
rows_f = nodes[rows // nv, rows % nv]; rows_t = nodes[rows // nv, col]
t = time.time()
o = orc.collide_cross((rows // nv).astype(np.int32)[::50], rows_f[::50], rows_t[::50], ca, cf, ct, 8, nthreads=0)
T("oracle edge_conflicts sample", t)
print("conflicts equal", np.array_equal(g[::50], o), flush=True)
