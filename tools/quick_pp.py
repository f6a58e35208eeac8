"""PP on point2d_many (N agents, PRMs(500, rad=0.1), scripts/config/eval/point2d_many.yaml:25-29):
time of the batched conflict tables vs the host search, and validate() of the result."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sssp_b200 as S
from sssp_b200 import harness as H

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50
res = []
for seed in range(3):
    # point2d_many.yaml:12-19 through gen_random_instance (valid starts / goals)
    init, goal, obstacles, (rads,) = H.gen_random_instance(
        S.StatePoint2D, np.random.default_rng(seed), N=N, num_obs=10, rad_min=0.015625,
        rad_max=0.0234375, rad_obs_min=0.025, rad_obs_max=0.04)
    connect = S.gen_connect(S.StatePoint2D, obstacles, rads)
    collide = S.gen_collide(S.StatePoint2D, rads)
    check_goal = S.gen_check_goal(goal)
    stats = {}
    t0 = time.perf_counter()
    sol, rms = S.PP(init, goal, connect, collide, check_goal, num_vertices=500, rad=0.1,
                    order_randomize=True, seed=seed, TIME_LIMIT=30, stats=stats,
                    max_makespan=64, roadmaps_growing_rate=1.0)
    t1 = time.perf_counter()
    ok = None if sol is None else bool(S.validate(init, connect, collide, check_goal, sol))
    res.append(dict(seed=seed, N=N, solved=sol is not None, makespan=None if sol is None else len(sol),
                    total_s=round(t1 - t0, 3), tables_s=round(stats.get("table_s", 0), 3),
                    search_s=round(stats.get("search_s", 0), 3), attempts=stats.get("attempts", 1),
                    valid=ok))
    print(json.dumps(res[-1]), flush=True)
