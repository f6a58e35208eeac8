"""The oracle's restatement of the reference's smoother (oracle/smoother_ref.py, smoother.jl) on
the golden SSSP solutions: structural invariants of the temporal plan graph that hold by
construction, and the hand-derived example of the module's docstring conventions."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import smoother_ref as SR

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "sssp_*.json")))
SMALL = [p for p in GOLD if "n30" not in p and "n50" not in p]


def load_solution(path):
    with open(path) as f:
        fx = json.load(f)
    inst = fx["instance"]
    orc = O.Oracle(inst["model"], inst["rads"], np.array(inst["obstacles"]),
                   base=None if inst["base"] is None else np.array(inst["base"]),
                   aux=None if inst.get("aux") is None else np.array(inst["aux"]))
    tabs = [np.array([[float.fromhex(x) for x in q] for q in g["nodes"]]) for g in fx["roadmaps"]]
    N = len(inst["rads"])
    sol = [[SR.Vertex(tabs[i][ids[i]], int(ids[i])) for i in range(N)] for ids in fx["solution"]]
    return fx, inst, orc, sol


@pytest.mark.parametrize("path", SMALL, ids=[os.path.basename(p)[:-5] for p in SMALL])
def test_tpg_invariants(path):
    fx, inst, orc, sol = load_solution(path)
    collide, connect, dist = SR.oracle_closures(orc)
    N = len(sol[0])
    TPG, greedy, cost = SR.smoothing(sol, connect, collide, dist)
    for i in range(N):
        ids = [a.id for a in TPG[i]]
        assert len(set(ids)) == len(ids)
        if TPG[i]:                       # the actions of an agent chain from its start to its goal
            assert TPG[i][0].frm.id == sol[0][i].id and TPG[i][-1].to.id == sol[-1][i].id
            for a, b in zip(TPG[i][:-1], TPG[i][1:]):
                assert a.to.id == b.frm.id and a.t < b.t
        for a in TPG[i]:                 # predecessors and successors mirror each other
            for (j, id_) in a.predecessors:
                assert (i, a.id) in SR._find(TPG, j, id_).successors
            for (j, id_) in a.successors:
                assert (i, a.id) in SR._find(TPG, j, id_).predecessors
            assert len({j for (j, _) in a.predecessors}) == len(a.predecessors)   # one per agent
    # the greedy schedule starts at the start, ends at the goal and only moves along actions
    assert [v.id for v in greedy[0]] == [v.id for v in sol[0]]
    assert [v.id for v in greedy[-1]] == [v.id for v in sol[-1]]
    assert len(greedy) <= len(sol) + 1
    # executing it is collision free and connected under the oracle closures
    for a, b in zip(greedy[:-1], greedy[1:]):
        for i in range(N):
            assert connect(a[i].q, b[i].q, i)
        hit, _ = orc.collide_configs(np.stack([v.q for v in a]), np.stack([v.q for v in b]))
        assert not hit
    assert cost["makespan"] <= cost["sum_of_cost"] + 1e-12 and cost["makespan"] > 0


def test_two_agents_crossing_by_hand():
    """Two point agents whose second moves cross: the later action must wait for the earlier one.
    Agent 0: a -> b -> c along y = 0.5; agent 1 crosses that line during its second move."""
    rads = np.array([0.05, 0.05])
    orc = O.Oracle(O.POINT2D, rads, np.zeros((0, 3)))
    collide, connect, dist = SR.oracle_closures(orc)
    V = lambda x, y, k: SR.Vertex(np.array([x, y]), k)
    a0, a1, a2 = V(0.2, 0.5, 0), V(0.5, 0.5, 1), V(0.8, 0.5, 2)
    b0, b1, b2 = V(0.65, 0.9, 0), V(0.65, 0.8, 1), V(0.65, 0.2, 2)
    sol = [[a0, b0], [a1, b0], [a1, b1], [a2, b1], [a2, b2]]
    TPG = SR.get_temporal_plan_graph(sol, collide, connect, skip_connection=False)
    assert [a.id for a in TPG[0]] == ["0_1_1", "1_2_3"] and [a.id for a in TPG[1]] == ["0_1_2", "1_2_4"]
    # agent 1's second move (t = 4) crosses agent 0's second move (t = 3): it depends on it
    assert (0, "1_2_3") in TPG[1][1].predecessors and (1, "1_2_4") in TPG[0][1].successors
    # with skip connection agent 0 may merge a -> b -> c into one straight move
    TPG2 = SR.get_temporal_plan_graph(sol, collide, connect, skip_connection=True)
    assert sum(len(x) for x in TPG2) <= sum(len(x) for x in TPG)
