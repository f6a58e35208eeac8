"""Scratch timing of mrmp_sssp_expand: raw C-ABI call vs Python wrapper."""
import ctypes as C, os, sys, time, math
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sssp_b200 as S
from sssp_b200 import lib as L, workload as W
lib = L.load()
inst = W.point2d_many(50, seed=0)
N, nv0, K = 50, 300, 10
ctx = S.Context(inst.model, inst.rads, inst.obstacles)
rm = S.Roadmaps(ctx, 0, N, 500); rm.sample(500, 0, inst.q_init, inst.q_goal); nodes = rm.nodes()
dev = S.SsspRoadmaps(ctx, nv_cap=4096)
for a in range(N): dev.set_roadmap(a, nodes[a, :nv0])
rng = np.random.default_rng(1)
ids = rng.integers(0, nv0, N).astype(np.int32)
old = rng.choice(nv0, 12, replace=False).astype(np.int32)
qr = rng.random((K, 2))
words = (nv0 + 2000 + 31) // 32
q_new = np.empty((K, 2)); ob = np.zeros((K, words), np.uint32); ib = np.zeros((K, words), np.uint32)
succ = np.empty(old.size + K + 1, np.int32); hit = np.empty(old.size + K + 1, np.uint8)
n_new, n_succ, w = C.c_int32(0), C.c_int32(0), C.c_int32(0)
def raw(a, thr):
    return lib.mrmp_sssp_expand(dev._h, a, int(ids[a]), K, qr.ctypes.data, 0, 0, 2, 0.2, thr, ids.ctypes.data,
                                old.size, old.ctypes.data, 0, C.byref(n_new), q_new.ctypes.data, C.byref(w),
                                ob.ctypes.data, ib.ctypes.data, ob.size, C.byref(n_succ), succ.ctypes.data, hit.ctypes.data)
for thr in (10.0, 0.02):     # 10.0: nothing accepted (steer + filter only); 0.02: vertices accepted
    for _ in range(5): L.check(raw(0, thr))
    t0 = time.perf_counter()
    for k in range(200): L.check(raw(k % N, thr))
    t1 = time.perf_counter()
    print("raw C-ABI call, thr=%g: %.1f us/call (n_new last=%d)" % (thr, (t1 - t0) / 200 * 1e6, n_new.value))
t0 = time.perf_counter()
for k in range(200): dev.expand(k % N, int(ids[k % N]), ids, old, K, q_rand=qr, steering_depth=2, epsilon=0.2, min_dist_thread=10.0)
t1 = time.perf_counter()
print("python wrapper: %.1f us/call" % ((t1 - t0) / 200 * 1e6))
# tiny single check for the launch+sync floor
q = np.array([[0.5, 0.5]]); ag = np.array([0], np.int32); o = np.empty(1, np.uint8)
t0 = time.perf_counter()
for k in range(200): lib.mrmp_connect_points(ctx._h, 1, ag.ctypes.data, q.ctypes.data, o.ctypes.data)
t1 = time.perf_counter()
print("single connect(q,i) call: %.1f us" % ((t1 - t0) / 200 * 1e6))
