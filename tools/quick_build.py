"""Scratch timing of the roadmap build (bench workload: point2d_many N=50, nv=500, rad=0.1)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sssp_b200 as S
from sssp_b200 import workload as W

inst = W.point2d_many(50, seed=0, nv=500, rad=0.1)
ctx = S.Context(inst.model, inst.rads, inst.obstacles)
N, nv = 50, 500
rm = S.Roadmaps(ctx, 0, N, nv)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
res = {}
for label, tries in (("sync_sample", True), ("deferred_sample", False)):
    for _ in range(3):
        rm.sample(nv, inst.seed, inst.q_init, inst.q_goal, tries=tries)
        rm.build(inst.rad)
    torch.cuda.synchronize()
    ts, tb = 0.0, 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        rm.sample(nv, inst.seed, inst.q_init, inst.q_goal, tries=tries)
        t1 = time.perf_counter()
        nnz = rm.build(inst.rad)
        t2 = time.perf_counter()
        ts += t1 - t0
        tb += t2 - t1
    res[label] = dict(sample_us=ts / reps * 1e6, build_us=tb / reps * 1e6, nnz=nnz)
print(json.dumps(res))
