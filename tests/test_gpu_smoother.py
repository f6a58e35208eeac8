"""GPU parity of the batched smoother (sssp_b200/smoother.py) with the per-call restatement of
src/smoother.jl (oracle/smoother_ref.py) on the golden SSSP solutions: the same temporal plan
graph, greedy schedule and cost; and the planning server's reply format (scripts/server.jl)."""
import glob
import json
import os

import numpy as np
import pytest

import sssp_b200 as S
from sssp_b200 import harness as H, smoother as SM
from oracle import oracle as O
from oracle import smoother_ref as SR
from test_gpu_sssp import closures, load
from test_smoother_ref import load_solution

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "sssp_*.json")))
CASES = [p for p in GOLD if "n50" not in p]


def _same_tpg(TPG, ref):
    assert len(TPG) == len(ref)
    for acts, racts in zip(TPG, ref):
        assert [a.id for a in acts] == [a.id for a in racts]
        for a, r in zip(acts, racts):
            assert (a.t, a.agent, a.frm.id, a.to.id) == (r.t, r.agent, r.frm.id, r.to.id)
            assert a.predecessors == r.predecessors and sorted(a.successors) == sorted(r.successors)


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-5] for p in CASES])
def test_smoothing_identical_to_the_per_call_restatement(path):
    fx, inst, orc, rsol = load_solution(path)
    _, _, _, base = load(path)
    connect, collide, check_goal, State = closures(inst, base)
    N = len(inst["rads"])
    sol = [[S.Node(State(*v.q), v.id, []) for v in Q] for Q in rsol]
    ocollide, oconnect, odist = SR.oracle_closures(orc)
    for skip in (False, True):
        ref = SR.get_temporal_plan_graph(rsol, ocollide, oconnect, skip_connection=skip)
        got = SM.get_temporal_plan_graph(sol, collide, connect, skip_connection=skip)
        _same_tpg(got, ref)
    TPG, greedy, cost = SM.smoothing(sol, connect, collide)
    rTPG, rgreedy, rcost = SR.smoothing(rsol, oconnect, ocollide, odist)
    _same_tpg(TPG, rTPG)
    assert [[v.id for v in Q] for Q in greedy] == [[v.id for v in Q] for Q in rgreedy]
    assert cost == rcost            # bit-identical doubles
    assert sum(len(a) for a in TPG) > 0
    # the refined schedule is a valid solution of the instance
    init = [State(*q) for q in inst["config_init"]]
    assert S.validate(init, connect, collide, check_goal, greedy)


def test_planning_server_reply_format():
    """scripts/server.jl:84-122: request in field coordinates -> {"status", "instructions"}."""
    req = {"field": {"x": 100.0, "y": 50.0, "size": 400.0},
           "agents": [{"x_init": 140, "y_init": 250, "x_goal": 460, "y_goal": 250, "rad": 16},
                      {"x_init": 300, "y_init": 90, "x_goal": 300, "y_goal": 410, "rad": 16},
                      {"x_init": 460, "y_init": 410, "x_goal": 140, "y_goal": 90, "rad": 16}],
           "obstacles": [{"x": 220, "y": 170, "rad": 20}],
           "solver": {"_target_": "MRMP.Solvers.SSSP", "TIME_LIMIT": 30, "seed": 1}}
    res = H.planning(req)
    assert res["status"] == "success"
    json.dumps(res)                                      # serialisable as is
    ins = res["instructions"]
    assert len(ins) == 3 and all(len(a) > 0 for a in ins)
    ids = [{a["id"] for a in acts} for acts in ins]
    for i, acts in enumerate(ins):
        assert set(acts[0]) == {"id", "x_from", "y_from", "x_to", "y_to", "pre", "suc"}
        # first action leaves the start, last reaches the goal, in FIELD coordinates
        assert abs(acts[0]["x_from"] - req["agents"][i]["x_init"]) < 1e-9
        assert abs(acts[-1]["y_to"] - req["agents"][i]["y_goal"]) < 1e-9
        for a in acts:
            for j, ref in a["pre"] + a["suc"]:           # 1-based agents, ids that exist
                assert 1 <= j <= 3 and ref in ids[j - 1]
    # an invalid instance (overlapping starts) fails like the server's @fail_to_solve
    bad = json.loads(json.dumps(req))
    bad["agents"][1]["x_init"], bad["agents"][1]["y_init"] = 141, 251
    assert H.planning(bad) == {"status": "failure"}
