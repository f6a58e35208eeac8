"""GPU parity of the SSSP hot loops: the fused node expansion (mrmp_sssp_expand) against the
oracle's expand! / successor filter, and whole SSSP runs (sssp_b200.solvers.SSSP on the CUDA
path) against the committed golden solutions and the live oracle -- identical solutions,
identical roadmaps (vertices bit for bit, neighbour lists in the same order)."""
import glob
import math
import json
import os

import numpy as np
import pytest

import sssp_b200 as S
from sssp_b200 import lib as L, solvers
from oracle import oracle as O
from oracle import sssp_ref as R

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "sssp_*.json")))
STATE = {O.POINT2D: S.StatePoint2D, O.POINT3D: S.StatePoint3D, O.ARM22: S.StateArm22,
         O.LINE2D: S.StateLine2D, O.SNAKE2D: S.StateSnake2D, O.ARM33: S.StateArm33,
         O.CAPSULE3D: S.StateCapsule3D}


def load(path):
    with open(path) as f:
        fx = json.load(f)
    inst = fx["instance"]
    base = None if inst["base"] is None else np.array(inst["base"])
    aux = None if inst.get("aux") is None else np.array(inst["aux"])
    orc = O.Oracle(inst["model"], inst["rads"], np.array(inst["obstacles"]), base=base, aux=aux)
    return fx, inst, orc, base


def closures(inst, base):
    State = STATE[inst["model"]]
    tag = State(*([0.0] * len(State._fields)))
    params = (inst["rads"],) if base is None else (base, inst["rads"])
    if inst.get("aux") is not None:
        params = (inst["rads"], inst["aux"])
    connect = S.gen_connect(tag, np.array(inst["obstacles"]), *params)
    collide = S.gen_collide(tag, *params)
    goal = [State(*q) for q in inst["config_goal"]]
    return connect, collide, S.gen_check_goal(goal), State


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-5] for p in GOLD])
def test_sssp_identical_to_golden(path):
    fx, inst, orc, base = load(path)
    connect, collide, check_goal, State = closures(inst, base)
    st = {}
    sol, rms = solvers.SSSP([State(*q) for q in inst["config_init"]],
                            [State(*q) for q in inst["config_goal"]], connect, collide, check_goal,
                            TIME_LIMIT=None, seed=fx["seed"], stats=st, **fx["params"])
    assert sol is not None
    assert st["solution_ids"] == fx["solution"] and st["expansions"] == fx["expansions"]
    for rm, g in zip(rms, fx["roadmaps"]):
        assert [[float(x).hex() for x in v.q] for v in rm] == g["nodes"]
        assert [list(v.neighbors) for v in rm] == g["neighbors"]
    # reference return shape: Vector{Vector{Node}} per time step, roadmaps per agent
    assert len(sol[0]) == len(inst["rads"]) and isinstance(sol[0][0], S.Node)


@pytest.mark.parametrize("model,slow", [(O.POINT2D, False), (O.POINT2D, True), (O.POINT3D, False),
                                        (O.ARM22, False)])
def test_expand_matches_oracle(model, slow):
    """One fused call == expand! + successor filter of the oracle, over several consecutive
    expansions on the same growing roadmap (the second sees the vertices of the first)."""
    path = [p for p in GOLD if {O.POINT2D: "point2d_many", O.POINT3D: "point3d", O.ARM22: "arm22"}[model] in p][0]
    fx, inst, orc, base = load(path)
    N = len(inst["rads"])
    ctx = S.Context(inst["model"], inst["rads"], np.array(inst["obstacles"]), base_pos=base,
                    aux=inst.get("aux"))
    # start from the golden (oracle-built) initial roadmaps truncated to their first vertices
    rms = []
    for g in fx["roadmaps"]:
        keep = min(8, len(g["nodes"]))
        rm = [R.Node([float.fromhex(x) for x in q], v, [u for u in nb if u < keep])
              for v, (q, nb) in enumerate(zip(g["nodes"][:keep], g["neighbors"][:keep]))]
        rms.append(rm)
    dev = S.closures.SsspRoadmaps(ctx, nv_cap=8)            # forces the tables to grow
    for i in range(N):
        dev.set_roadmap(i, np.stack([v.q for v in rms[i]]))
    eps = fx["params"].get("epsilon")
    depth = fx["params"].get("steering_depth", 2)
    K = 12 if model != O.ARM22 else 5
    rng = np.random.default_rng(5)
    t = [0] * N
    thr = 0.08
    for step in range(6):
        i = step % N
        v = int(rng.integers(0, len(rms[i])))
        ids = [int(rng.integers(0, len(rms[j]))) for j in range(N)]
        ids[i] = v
        old = list(rms[i][v].neighbors)
        q_new, outs, ins, succ, hit = dev.expand(i, v, ids, old, K, seed=9, t0=t[i], steering_depth=depth,
                                                 epsilon=eps, min_dist_thread=thr, slow_collide=slow)
        # oracle: same samples, sequential expand!
        nv0 = len(rms[i])
        samples = iter([orc.sample_state(9, i, t[i] + k) for k in range(K)])
        t[i] += K
        conn = lambda a, b: (eps is None or orc.state_dist(a, b) <= eps) and orc.connect_edge(a, b, i)
        R.expand(orc, conn, lambda: next(samples), rms[i][v], rms[i], thr, K, depth)
        new = rms[i][nv0:]
        assert len(new) == len(q_new)
        for r, u in enumerate(new):
            assert np.array_equal(u.q.view(np.uint64), q_new[r].view(np.uint64))
            assert [x for x in u.neighbors if x < u.id] == outs[r]
            want_in = [w.id for w in rms[i][: u.id + 1] if u.id in w.neighbors]
            assert want_in == ins[r]
        assert succ.tolist() == list(rms[i][v].neighbors) + [v]
        Q = np.stack([rms[j][ids[j]].q for j in range(N)])
        for p_id, h in zip(succ.tolist(), hit.tolist()):
            if slow:
                Qt = Q.copy()
                Qt[i] = rms[i][p_id].q
                want = orc.collide_configs(Q, Qt)[0]
            else:
                want = orc.collide_config_move(Q, rms[i][p_id].q, i)
            assert bool(h) == want
        assert np.array_equal(dev.nodes(i).view(np.uint64), np.stack([w.q for w in rms[i]]).view(np.uint64))
        thr *= 0.9


def test_model_helpers_match_oracle():
    rng = np.random.default_rng(3)
    for model, d in ((O.POINT2D, 2), (O.POINT3D, 3), (O.ARM22, 2)):
        base = rng.uniform(0.3, 0.7, (2, 2)) if model == O.ARM22 else None
        ctx = S.Context(model, [0.05, 0.05], np.zeros((0, d + 1)), base_pos=base)
        orc = O.Oracle(model, [0.05, 0.05], np.zeros((0, d + 1)), base=base)
        a = rng.random((500, d)) * (2 * np.pi if model == O.ARM22 else 1)
        b = rng.random((500, d)) * (2 * np.pi if model == O.ARM22 else 1)
        dist = ctx.state_dist(a, b)
        mid = ctx.mid_status(a, b)
        for k in range(500):
            assert dist[k].view(np.uint64) == np.float64(orc.state_dist(a[k], b[k])).view(np.uint64)
            assert np.array_equal(mid[k].view(np.uint64), orc.mid_status(a[k], b[k]).view(np.uint64))
        V = a[:300].copy()
        V[17] = V[5]                      # duplicate: the first minimum wins
        for rev in (False, True):
            for q in (b[0], V[17]):
                ds = [orc.state_dist(q, v) if rev else orc.state_dist(v, q) for v in V]
                idx, dd = ctx.nearest(V, q, reversed=rev)
                assert idx == int(np.argmin(ds)) and dd == min(ds)


def test_distance_tables_match_the_oracle_dijkstra():
    """mrmp_distance_tables (parallel relaxation to the fixed point) == the reference's Dijkstra
    (solvers/utils.jl:146-179) bit for bit, for PRM roadmaps of all three models incl. vertices
    the goal cannot reach (Inf) and several roadmaps per launch."""
    from helpers import make_point_instance, make_arm_instance
    rng = np.random.default_rng(12)
    for model, d in ((O.POINT2D, 2), (O.POINT3D, 3), (O.ARM22, 2)):
        N, nv = 5, 260
        if model == O.ARM22:
            rads, obs, base = make_arm_instance(N, 2, 19)
            nv = 90
        else:
            (rads, obs), base = make_point_instance(model, N, 10, seed=3), None
        ctx = S.Context(model, rads, obs, base_pos=base)
        orc = O.Oracle(model, rads, obs, base=base)
        nodes = np.stack([orc.sample_free(a, 5, nv)[0] for a in range(N)])
        if model != O.ARM22:
            nodes[:, nv - 1] = 3.0          # outside the field: no edge leads there -> Inf
        rad = 0.12 if model == O.POINT2D else 0.3
        rp, col, nnz = orc.prm_build(0, nodes, rad, nthreads=0)
        graphs, want = [], []
        for a in range(N):
            nb = [col[a, rp[a, v]:rp[a, v + 1]].tolist() for v in range(nv)]
            graphs.append((nodes[a], nb))
            want.append(orc.distance_table(nodes[a], rp[a], col[a, :nnz[a]], goal=1))
        got = ctx.distance_tables(graphs, goal=1)
        for a in range(N):
            assert np.array_equal(got[a].view(np.uint64), want[a].view(np.uint64)), (model, a)
        assert model == O.ARM22 or all(np.isinf(w[nv - 1]) for w in want)
        assert all(np.isfinite(w).sum() > nv // 2 for w in want)
    # the host-side helper of the solver mirror
    t = solvers.get_distance_table(ctx, graphs[0][0], graphs[0][1])
    assert np.array_equal(np.array(t).view(np.uint64), want[0].view(np.uint64))


def _oracle_validate(orc, N, sol):
    """utils.jl:345-363 as the sequential loop over the oracle closures -> first failure."""
    for t in range(1, len(sol)):
        for i in range(N):
            if not orc.connect_edge(sol[t - 1][i], sol[t][i], i):
                return ("not connected", t, i)
        hit, fp = orc.collide_configs(sol[t - 1], sol[t])
        if hit:
            return ("colliding", t, fp // N, fp % N)
    return None


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-5] for p in GOLD])
def test_validate_matches_the_sequential_reference_loop(path):
    """validate (utils.jl:320-367) in two launches == the oracle's per-step loop: golden SSSP
    solutions are valid; corrupted ones fail at the same first check."""
    fx, inst, orc, base = load(path)
    connect, collide, check_goal, State = closures(inst, base)
    N = len(inst["rads"])
    tabs = [np.array([[float.fromhex(x) for x in q] for q in g["nodes"]]) for g in fx["roadmaps"]]
    sol = np.array([[tabs[i][ids[i]] for i in range(N)] for ids in fx["solution"]])
    as_nodes = lambda arr: [[S.Node(State(*q), 0) for q in Q] for Q in arr]
    init = [State(*q) for q in inst["config_init"]]
    assert S.validate(init, connect, collide, check_goal, as_nodes(sol)) is True
    assert _oracle_validate(orc, N, sol) is None
    assert S.validate(init, connect, collide, check_goal, None) is True
    assert S.is_valid_instance(init, [State(*q) for q in inst["config_goal"]], connect, collide)
    rng = np.random.default_rng(8)
    seen = set()
    for trial in range(12):
        bad = sol.copy()
        t = int(rng.integers(1, len(sol) - 1)) if len(sol) > 2 else 1
        i = int(rng.integers(0, N))
        if trial % 2 == 0:      # teleport one agent for one step: usually disconnects or collides
            bad[t, i] = tabs[i][int(rng.integers(0, len(tabs[i])))]
        else:                   # put agent i on top of agent j's state
            j = (i + 1) % N
            bad[t:, i] = bad[t, j] if inst["model"] != O.ARM22 else bad[t, i] + 1.0
        want = _oracle_validate(orc, N, bad)
        got = S.validate(init, connect, collide, lambda Q: True, as_nodes(bad))
        assert got is (want is None)
        if want is not None:
            assert S.validate.last_failure == want
            seen.add(want[0])
    assert seen, "no corrupted solution was rejected"
    # a wrong first configuration is caught before any GPU work
    bad = sol.copy()
    bad[0, 0] = bad[0, 0] + 0.25
    assert S.validate(init, connect, collide, check_goal, as_nodes(bad)) is False


def _oracle_first_conflict(orc, N, cfg):
    """get_constraints (cbs.jl:384-408) as the reference's loop over the oracle closures."""
    for t in range(1, len(cfg)):
        for i in range(N):
            for j in range(i + 1, N):
                if orc.collide_static(cfg[t][i], cfg[t][j], i, j):
                    return (t, "vertex", i, j)
                if orc.collide_pair(cfg[t - 1][i], cfg[t][i], cfg[t - 1][j], cfg[t][j], i, j):
                    return (t, "edge", i, j)
    return None


@pytest.mark.parametrize("name", ["point2d_many", "point3d", "arm22"])
def test_first_conflict_matches_get_constraints(name):
    """Independent single-agent plans (random walks on the golden roadmaps) conflict somewhere:
    the batched scan must name the same first (t, kind, i, j) as the reference's loop."""
    fx, inst, orc, base = load([p for p in GOLD if name in p][0])
    ctx = S.Context(inst["model"], inst["rads"], np.array(inst["obstacles"]), base_pos=base,
                    aux=inst.get("aux"))
    N = len(inst["rads"])
    tabs = [np.array([[float.fromhex(x) for x in q] for q in g["nodes"]]) for g in fx["roadmaps"]]
    nbrs = [g["neighbors"] for g in fx["roadmaps"]]
    rng = np.random.default_rng(31)
    kinds = set()
    for trial in range(10):
        T = 12
        ids = np.zeros((T, N), int)
        for i in range(N):
            v = int(rng.integers(0, len(tabs[i])))
            for t in range(T):
                ids[t, i] = v
                if nbrs[i][v] and rng.random() < 0.8:
                    v = int(nbrs[i][v][int(rng.integers(0, len(nbrs[i][v])))])
        cfg = np.array([[tabs[i][ids[t, i]] for i in range(N)] for t in range(T)])
        want = _oracle_first_conflict(orc, N, cfg)
        assert ctx.first_conflict(cfg) == want
        if want:
            kinds.add(want[1])
    # the golden SSSP solution itself is conflict free
    sol = np.array([[tabs[i][q[i]] for i in range(N)] for q in fx["solution"]])
    assert ctx.first_conflict(sol) is None and _oracle_first_conflict(orc, N, sol) is None
    assert kinds, "no conflict was exercised"


def test_distance_tables_with_stay_penalty_edge_cost():
    """mrmp_distance_tables_g: the edge cost of gen_g_func(stay_penalty) (solvers/utils.jl:34-36)
    -- dist + stay_penalty between bit-identical states -- as PP / CBS build their tables
    (pp.jl:161).  Duplicated vertices make the penalty reachable through the neighbour lists."""
    from helpers import make_point_instance
    from oracle import pp_ref
    from oracle.sssp_ref import Node as ONode
    N, nv = 3, 120
    rads, obs = make_point_instance(O.POINT2D, N, 6, seed=9)
    ctx, orc = S.Context(O.POINT2D, rads, obs), O.Oracle(O.POINT2D, rads, obs)
    nodes = np.stack([orc.sample_free(a, 7, nv)[0] for a in range(N)])
    nodes[:, 40:60] = nodes[:, 60:80]            # exact duplicates: dist 0, q_from == q_to
    nodes[:, 1] = nodes[:, 41]                   # the goal itself has a twin
    rp, col, nnz = orc.prm_build(0, nodes, 0.2, nthreads=0)
    graphs = [(nodes[a], [col[a, rp[a, v]:rp[a, v + 1]].tolist() for v in range(nv)])
              for a in range(N)]
    for pen in (0.1, 0.0, 2.5):
        g = pp_ref.gen_g_func(orc, False, pen)
        got = ctx.distance_tables(graphs, goal=1, stay_penalty=pen)
        for a in range(N):
            rm = [ONode(nodes[a, v], v, graphs[a][1][v]) for v in range(nv)]
            want = np.array(pp_ref.get_distance_table(rm, g, 1))
            assert np.array_equal(got[a].view(np.uint64), want.view(np.uint64)), (pen, a)
    twin = ctx.distance_tables(graphs, goal=1, stay_penalty=0.1)[0][41]
    assert twin == 0.1                           # one hop goal -> its twin costs the penalty


def test_validate_runs_collide_checks_on_the_collide_closure():
    """connect and collide are independent closures in the reference (arm22.jl:29-37 vs :93-99):
    a collide built with its own safety_dist must be the one that decides, and SSSP refuses a
    pair of closures it cannot evaluate on one context."""
    A = S.StateArm22
    base, rads = [[0.35, 0.5], [0.65, 0.5]], [0.1, 0.1]
    connect = S.gen_connect(A(0, 0), [], base, rads)
    loose = S.gen_collide(A(0, 0), base, rads)                       # safety_dist 0.01
    strict = S.gen_collide(A(0, 0), base, rads, safety_dist=0.2)
    # arm 0 points right (tip at x = 0.55), arm 1 points up: closest approach 0.10
    q0, q1 = A(0.0, 0.0), A(math.pi / 2, 0.0)
    sol = [[S.Node(q0, 0), S.Node(q1, 0)], [S.Node(q0, 0), S.Node(q1, 0)]]
    ok = lambda Q: True
    assert loose(q0, q1, 0, 1) is False and strict(q0, q1, 0, 1) is True
    assert S.validate([q0, q1], connect, loose, ok, sol) is True
    assert S.validate([q0, q1], connect, strict, ok, sol) is False
    assert S.validate.last_failure == ("colliding", 1, 0, 1)
    with pytest.raises(ValueError):
        S.SSSP([q0, q1], [q0, q1], connect, strict, S.gen_check_goal([q0, q1]), TIME_LIMIT=1)
    # goal radius on a non-point model: needs the model's metric
    with pytest.raises(ValueError):
        S.gen_check_goal([q0, q1], goal_rad=0.1)
    near = S.gen_check_goal([q0, q1], goal_rad=0.1, ctx=connect.ctx)
    assert near(A(2 * math.pi - 0.05, 0.0), 0) is True     # wraps around: dist = 0.05 / pi
    assert near(A(1.0, 0.0), 0) is False


@pytest.mark.parametrize("model", [O.POINT2D, O.POINT3D, O.ARM22])
def test_rrt_extend_equals_the_per_call_sequence(model):
    """mrmp_rrt_extend (one launch) == extend! of sssp.jl:404-440 spelled out with one oracle
    closure call per step: nearest vertex, connection test, bisection, from both trees."""
    path = [p for p in GOLD if {O.POINT2D: "point2d_many_n10", O.POINT3D: "point3d", O.ARM22: "arm22"}[model] in p][0]
    fx, inst, orc, base = load(path)
    ctx = S.Context(inst["model"], inst["rads"], np.array(inst["obstacles"]), base_pos=base,
                    aux=inst.get("aux"))
    N, d = len(inst["rads"]), ctx.d
    rng = np.random.default_rng(12)
    seen = set()
    for trial in range(120):
        i = int(rng.integers(0, N))
        eps = [None, 0.2, 0.5][trial % 3]
        depth = [0, 2, 4][trial % 3]
        from_start = bool(trial % 2)
        V = [orc.sample_free(i, 100 + trial, 1)[0][0] for _ in range(int(rng.integers(1, 12)))]
        q_rand = orc.sample_state(7, i, trial)
        conn2 = lambda a, b: (eps is None or orc.state_dist(a, b) <= eps) and orc.connect_edge(a, b, i)
        conn1 = lambda q: orc.connect_point(q, i)
        dists = [orc.state_dist(v, q_rand) if from_start else orc.state_dist(q_rand, v) for v in V]
        near = int(np.argmin(dists))
        q_near = V[near]
        flg, q_l, q_h = "reached", q_near, q_rand
        linked = (lambda x: conn2(q_near, x)) if from_start else (lambda x: conn1(x) and conn2(x, q_near))
        if not linked(q_h):
            flg = "trapped"
            for _ in range(depth):
                q = orc.mid_status(q_l, q_h) if from_start else orc.mid_status(q_h, q_l)
                if linked(q):
                    flg, q_l = "advanced", q
                else:
                    q_h = q
            q_h = q_l
        got_flg, got_q, got_near = ctx.rrt_extend(i, np.stack(V), q_rand, from_start, depth, eps)
        assert (got_flg, got_near) == (flg, near), trial
        assert np.array_equal(np.asarray(got_q).view(np.uint64), np.asarray(q_h, float).view(np.uint64)), trial
        seen.add(flg)
    assert len(seen) >= 2
