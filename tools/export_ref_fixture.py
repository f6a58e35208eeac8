#!/usr/bin/env python
"""Fixture for julia/ref_bench.jl: the bench workload (or a small one) with the oracle's answers.

The reference cannot run in this image (pure Julia, no julia binary), so the oracle's parity is
"unpinned" (DESIGN.md section 2).  This script writes everything a machine WITH Julia needs to pin
it and to time the real reference on the same inputs:

  instance   rads, obstacles, rad                      (point2d_many.yaml:12-29)
  vertices   [N][nv][2]; vertex 0 / 1 = init / goal    (PRM! accepts a pre-filled roadmap, prm.jl:178-213)
  neighbors  the oracle's adjacency lists, 0-based     (what PRM! must reproduce, prm.jl:202-210)
  table      a row sample (agent, v_from, v_to), the path-edge columns (agent, v_from, v_to) and the
             oracle's verdict of collide(row, column) for every pair of the sample (point.jl:5-12)
  points     random states + agent and the oracle's connect(q, i) (point2d.jl:46-57)

Floats are written with repr() (shortest round trip), so Julia's JSON parser restores the same
binary64 values.

  python tools/export_ref_fixture.py tests/golden/ref_pin_point2d_n6.json --agents 6 --vertices 60
  python tools/export_ref_fixture.py /tmp/ref_bench_n50.json            # the BASELINE configs[1] workload
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make(n_agents: int, nv: int, rad: float, t_steps: int, row_stride: int, n_points: int, seed: int):
    from oracle import oracle as O
    from sssp_b200 import workload as W
    O.build()
    inst = W.point2d_many(n_agents, seed=seed, nv=nv, rad=rad)
    orc = O.Oracle(O.POINT2D, inst.rads, inst.obstacles)
    N = n_agents
    nodes = np.empty((N, nv, 2))
    nodes[:, 0], nodes[:, 1] = inst.q_init, inst.q_goal
    for a in range(N):
        nodes[a, 2:], _ = orc.sample_free(a, inst.seed, nv - 2)
    rp, colidx, nnz = orc.prm_build(0, nodes, rad, nthreads=O.host_cores())
    neighbors = [[colidx[a, rp[a, v]:rp[a, v + 1]].tolist() for v in range(nv)] for a in range(N)]
    row_ptr = np.zeros(N * nv + 1, np.int64)
    cols = []
    for a in range(N):
        row_ptr[a * nv + 1:(a + 1) * nv + 1] = row_ptr[a * nv] + rp[a, 1:]
        cols.append(colidx[a, :nnz[a]])
    col = np.concatenate(cols)
    _, ca, vf, vt = W.random_walk_paths(inst, row_ptr, col, t_steps, seed=4321)
    rows = np.repeat(np.arange(N * nv), np.diff(row_ptr))[::row_stride]
    ra = (rows // nv).astype(np.int32)
    rvf = (rows % nv).astype(np.int32)
    rvt = col[::row_stride].astype(np.int32)
    bits = orc.collide_cross(ra, nodes[ra, rvf], nodes[ra, rvt], ca, nodes[ca, vf], nodes[ca, vt], 0,
                             nthreads=O.host_cores())
    verdict = np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little")[:, :ca.size]
    rng = np.random.default_rng(seed + 99)
    pq = rng.random((n_points, 2)) * 1.1 - 0.05          # some outside the unit square
    pa = rng.integers(0, N, n_points).astype(np.int32)
    pv = orc.connect_points(pa, pq, nthreads=O.host_cores())
    return {
        "model": "StatePoint2D", "n_agents": N, "num_vertices": nv, "rad": rad,
        "rads": inst.rads.tolist(), "obstacles": inst.obstacles.tolist(),
        "vertices": nodes.tolist(), "neighbors": neighbors,
        "table": {"row_agent": ra.tolist(), "row_vfrom": rvf.tolist(), "row_vto": rvt.tolist(),
                  "col_agent": ca.tolist(), "col_vfrom": vf.tolist(), "col_vto": vt.tolist(),
                  # one string of '0'/'1' per row: collide(row edge, column edge, i, j); pairs of the
                  # same agent are never asked by the reference (pp.jl:183) and are '0' here
                  "verdict": ["".join("1" if b else "0" for b in r) for r in verdict]},
        "points": {"agent": pa.tolist(), "q": pq.tolist(), "connect": [int(b) for b in pv]},
        "indexing": "0-based agents and vertex ids (add 1 in Julia)",
        "generator": "tools/export_ref_fixture.py --agents %d --vertices %d --rad %r --steps %d "
                     "--row-stride %d --points %d --seed %d" % (N, nv, rad, t_steps, row_stride, n_points, seed),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--agents", type=int, default=50)
    ap.add_argument("--vertices", type=int, default=500)
    ap.add_argument("--rad", type=float, default=0.1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--row-stride", type=int, default=64)
    ap.add_argument("--points", type=int, default=2000)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    fx = make(a.agents, a.vertices, a.rad, a.steps, a.row_stride, a.points, a.seed)
    with open(a.out, "w") as f:
        json.dump(fx, f, separators=(",", ":"))
    print("%s: %d agents x %d vertices, %d edges, table sample %d rows x %d columns, %d points" % (
        a.out, fx["n_agents"], fx["num_vertices"], sum(len(x) for nb in fx["neighbors"] for x in nb),
        len(fx["table"]["row_agent"]), len(fx["table"]["col_agent"]), len(fx["points"]["agent"])))


if __name__ == "__main__":
    main()
