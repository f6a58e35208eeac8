"""Host logic of sssp_b200.pp (priority loop, table look-ups, space-time A*) on CPU: the
device calls are answered by a stand-in built on the oracle, and the result must equal the
per-call restatement of pp.jl.  (The real device path is tests/test_gpu_pp.py.)"""
import numpy as np
import pytest

from sssp_b200 import pp as PPmod
from sssp_b200.closures import Node, StatePoint2D
from oracle import oracle as O
from oracle import pp_ref as P
from oracle.sssp_ref import Node as ONode
from helpers import make_point_instance


class _OracleCtx:
    """Answers the three batched calls PP makes, from the CPU oracle."""

    def __init__(self, orc, d):
        self.orc, self.d = orc, d

    def state_dist(self, a, b):
        return np.array([self.orc.state_dist(x, y) for x, y in zip(a, b)])

    def collide_cross(self, ra, rf, rt, ca, cf, ct, group_size=0):
        return self.orc.collide_cross(ra, rf, rt, ca, cf, ct, group_size, nthreads=0)

    def distance_tables(self, graphs, goal=1, stay_penalty=0.0):
        g = P.gen_g_func(self.orc, False, stay_penalty)
        out = []
        for q, nbrs in graphs:
            rm = [ONode(q[v], v, list(nbrs[v])) for v in range(len(nbrs))]
            out.append(np.array(P.get_distance_table(rm, g, goal)))
        return out


def _instance(N, M, nv, rad, seed, rad_agents=(0.015625, 0.0234375)):
    rads, obs = make_point_instance(O.POINT2D, N, M, seed=seed, rad=rad_agents)
    orc = O.Oracle(O.POINT2D, rads, obs)
    nodes = np.zeros((N, nv, 2))
    for i in range(N):
        nodes[i, :2] = orc.sample_free(i, 1000 + seed, 2)[0]
        nodes[i, 2:] = orc.sample_free(i, seed, nv - 2)[0]
    rp, col, _ = orc.prm_build(0, nodes, rad, nthreads=0)
    nbrs = [[col[a, rp[a, v]:rp[a, v + 1]].tolist() for v in range(nv)] for a in range(N)]
    return orc, nodes, nbrs


@pytest.mark.parametrize("N,M,nv,rad,seed,rad_agents,kw", [
    (8, 6, 70, 0.25, 6, (0.015625, 0.0234375), {}),
    (8, 6, 70, 0.25, 11, (0.015625, 0.0234375), {"greedy": True}),
    (5, 3, 60, 0.3, 41, (0.03, 0.05), {"max_makespan": 25}),
    (5, 3, 60, 0.3, 42, (0.03, 0.05), {"max_makespan": 25}),     # planning fails
    (5, 3, 60, 0.3, 43, (0.03, 0.05), {"stay_penalty": 0.0, "max_makespan": 9}),
])
def test_pp_host_logic_equals_per_call_restatement(N, M, nv, rad, seed, rad_agents, kw):
    orc, nodes, nbrs = _instance(N, M, nv, rad, seed, rad_agents)
    o_rms = [[ONode(nodes[a, v], v, list(nbrs[a][v])) for v in range(nv)] for a in range(N)]
    g_rms = [[Node(StatePoint2D(*nodes[a, v]), v, list(nbrs[a][v])) for v in range(nv)]
             for a in range(N)]
    order = list(range(N))[::-1]
    want, _ = P.PP(orc, o_rms, [nodes[a, 1] for a in range(N)], order=order, **kw)

    def collide(*a):
        raise AssertionError("PP must not issue per-node collide calls")
    collide.ctx = _OracleCtx(orc, 2)

    def check_goal(v, i):
        return orc.state_dist(np.array(v.q), nodes[i, 1]) <= 0.0
    got, _ = PPmod.PP_on_roadmaps(g_rms, collide, check_goal, order=order, **kw)
    assert (None if got is None else [[v.id for v in Q] for Q in got]) == want


def test_grow_keeps_old_neighbour_lists_in_front():
    """PRMs! on existing roadmaps (prm.jl:178-213) re-runs the all-pairs pass over the whole
    roadmap and pushes onto the old lists."""
    old = [[Node((0.0, 0.0), 0, [1]), Node((1.0, 0.0), 1, [0])]]
    new = [[Node((0.0, 0.0), 0, [1, 2]), Node((1.0, 0.0), 1, [0]), Node((0.5, 0.5), 2, [0])]]
    rm = PPmod._grow(old, new)[0]
    assert [v.neighbors for v in rm] == [[1, 1, 2], [0, 0], [0]]
