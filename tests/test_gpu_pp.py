"""Prioritized planning (pp.jl) on the GPU conflict table against the CPU restatement that
issues one collide call per generated search node: identical priorities -> identical paths."""
import math

import numpy as np
import pytest

import sssp_b200 as S
from sssp_b200 import pp as PPmod
from oracle import oracle as O
from oracle import pp_ref as P
from oracle.sssp_ref import Node as ONode
from helpers import make_point_instance

pytestmark = pytest.mark.gpu


def _instance(model, d, N, M, nv, rad, seed, rad_agents=None):
    kw = {} if rad_agents is None else {"rad": rad_agents}
    rads, obs = make_point_instance(model, N, M, seed=seed, **kw)
    orc = O.Oracle(model, rads, obs)
    nodes = np.zeros((N, nv, d))
    for i in range(N):
        ends, _ = orc.sample_free(i, 1000 + seed, 2)
        rest, _ = orc.sample_free(i, seed, nv - 2)
        nodes[i, :2], nodes[i, 2:] = ends, rest
    rp, col, _ = orc.prm_build(0, nodes, rad, nthreads=0)
    nbrs = [[col[a, rp[a, v]:rp[a, v + 1]].tolist() for v in range(nv)] for a in range(N)]
    return rads, obs, orc, nodes, nbrs


def _both(model, d, rads, obs, orc, nodes, nbrs, order, **kw):
    N, nv = nodes.shape[:2]
    State = S.StatePoint2D if d == 2 else S.StatePoint3D
    o_rms = [[ONode(nodes[a, v], v, list(nbrs[a][v])) for v in range(nv)] for a in range(N)]
    g_rms = [[S.Node(State(*nodes[a, v]), v, list(nbrs[a][v])) for v in range(nv)] for a in range(N)]
    goal = [nodes[a, 1] for a in range(N)]
    cnt = {}
    want, _ = P.PP(orc, o_rms, goal, order=order, counters=cnt, **kw)
    collide = S.gen_collide(State, rads)
    check_goal = S.gen_check_goal([State(*q) for q in goal])
    stats = {}
    got, _ = PPmod.PP_on_roadmaps(g_rms, collide, check_goal, order=order, stats=stats, **kw)
    got_ids = None if got is None else [[v.id for v in Q] for Q in got]
    return want, got_ids, cnt


@pytest.mark.parametrize("seed,N,order", [(6, 8, None), (6, 12, None), (7, 12, "rev"),
                                          (11, 8, "rev"), (12, 12, None), (12, 12, "perm")])
def test_pp_point2d_identical_paths(seed, N, order):
    inst = _instance(O.POINT2D, 2, N, 6, 70, 0.25, seed)
    order = {None: list(range(N)), "rev": list(range(N))[::-1],
             "perm": np.random.default_rng(seed).permutation(N).tolist()}[order]
    want, got, cnt = _both(O.POINT2D, 2, *inst, order)
    assert got == want
    assert cnt["collide_calls"] > 100


@pytest.mark.parametrize("seed", [41, 43, 49, 50])
def test_pp_larger_agents_wait_detour_or_fail(seed):
    """Radii 0.03-0.05: agents wait (41, 43), detour (50) or the planning fails after 31 000
    collide calls (49) -- in every case exactly as the per-call restatement does."""
    inst = _instance(O.POINT2D, 2, 5, 3, 60, 0.3, seed, rad_agents=(0.03, 0.05))
    want, got, cnt = _both(O.POINT2D, 2, *inst, list(range(5)), max_makespan=25)
    assert got == want and cnt["collide_calls"] > 500
    assert (want is None) == (seed == 49)


def test_pp_point3d_and_greedy_cost():
    inst = _instance(O.POINT3D, 3, 6, 5, 60, 0.4, 3)
    want, got, _ = _both(O.POINT3D, 3, *inst, list(range(6)))
    assert got == want and want is not None
    want, got, _ = _both(O.POINT3D, 3, *inst, list(range(6)), greedy=True)
    assert got == want and want is not None
    want, got, _ = _both(O.POINT3D, 3, *inst, list(range(6)), stay_penalty=0.0, max_makespan=6)
    assert got == want


def test_pp_end_to_end_from_configurations():
    """PP(config_init, config_goal, connect, collide, check_goal; ...): roadmaps from the GPU PRM,
    then the same priorities on the CPU restatement over those roadmaps."""
    N, nv = 8, 80
    rads, obs = make_point_instance(O.POINT2D, N, 6, seed=12)
    orc = O.Oracle(O.POINT2D, rads, obs)
    ends = [orc.sample_free(i, 1012, 2)[0] for i in range(N)]
    init = [S.StatePoint2D(*e[0]) for e in ends]
    goal = [S.StatePoint2D(*e[1]) for e in ends]
    obstacles = [S.CircleObstacle2D(*o) for o in obs]
    connect = S.gen_connect(S.StatePoint2D, obstacles, rads)
    collide = S.gen_collide(S.StatePoint2D, rads)
    check_goal = S.gen_check_goal(goal)
    order = list(range(N))[::-1]
    got, rms = PPmod.PP(init, goal, connect, collide, check_goal, num_vertices=nv, rad=0.25,
                        order=order, seed=5, roadmaps_growing_rate=1.3, TIME_LIMIT=60)
    o_rms = [[ONode(np.array(v.q), v.id, list(v.neighbors)) for v in rm] for rm in rms]
    want, _ = P.PP(orc, o_rms, [np.array(g) for g in goal], order=order)
    assert (None if got is None else [[v.id for v in Q] for Q in got]) == want
    if got is not None:
        ok = S.validate(init, connect, collide, check_goal, got)
        assert ok
