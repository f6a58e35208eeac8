"""GPU parity of the count / first reductions of the cross-product collide
(mrmp_collide_cross_reduce) against the oracle's restatement of their reference callers:
CBS's soft conflict count (cbs.jl:336-347, get_num_collisions :449-457) and the type-2
dependency scan of the temporal plan graph (smoother.jl:78-106)."""
import numpy as np
import pytest

import sssp_b200 as S
from oracle import oracle as O
from helpers import make_point_instance, make_arm_instance, make_chain_instance, chain_states, random_edges

pytestmark = pytest.mark.gpu


def _edges(rng, n, d, N, spread=0.5, max_len=0.1):
    a = rng.integers(0, N, n).astype(np.int32)
    qf, qt = random_edges(rng, n, d, max_len)
    return a, qf * spread + (1 - spread) * 0.5, qt * spread + (1 - spread) * 0.5


@pytest.mark.parametrize("model,d", [(O.POINT2D, 2), (O.POINT3D, 3)])
@pytest.mark.parametrize("cbs", [False, True])
def test_count_per_time_step_matches_oracle(model, d, cbs):
    """rows = candidate edges of one agent (a roadmap's moves + stay moves), columns = the other
    agents' path edges, t-major: group = time step.  cbs=True is g_func_i's count verbatim."""
    N, T = 10, 37
    rads, obs = make_point_instance(model, N, 0, seed=5 + d, rad=(0.02, 0.05))
    ctx, orc = S.Context(model, rads, obs), O.Oracle(model, rads, obs)
    rng = np.random.default_rng(11 + d)
    n_rows = 1500
    _, rf, rt = _edges(rng, n_rows, d, N)
    rt[::7] = rf[::7]                                   # stay moves
    ra = np.full(n_rows, 3, np.int32)
    ca = np.tile(np.arange(N, dtype=np.int32), T)
    _, cf, ct = _edges(rng, N * T, d, N)
    ct[::5] = cf[::5]                                   # agents waiting at their goal
    g = ctx.collide_cross_reduce(ra, rf, rt, ca, cf, ct, "count", group_size=N, cbs=cbs)
    o = orc.collide_cross_reduce(ra, rf, rt, ca, cf, ct, "count", group_size=N, cbs=cbs, nthreads=0)
    assert g.shape == (n_rows, T) and np.array_equal(g, o)
    assert o.max() >= 2 and (o == 0).any()
    if cbs:   # the vertex test adds conflicts the edge test alone does not always see
        e = orc.collide_cross_reduce(ra, rf, rt, ca, cf, ct, "count", group_size=N, nthreads=0)
        assert (o >= e).all()


@pytest.mark.parametrize("model,d", [(O.POINT2D, 2), (O.POINT3D, 3)])
def test_first_later_action_per_agent_matches_oracle(model, d):
    """TPG type-2 scan: actions of all agents as rows and columns, group = column agent
    (explicit col_group, ragged), key = time step; first colliding later action per agent."""
    N = 8
    rads, obs = make_point_instance(model, N, 0, seed=9 + d, rad=(0.03, 0.06))
    ctx, orc = S.Context(model, rads, obs), O.Oracle(model, rads, obs)
    rng = np.random.default_rng(21 + d)
    n_act = rng.integers(5, 60, N)
    agent = np.repeat(np.arange(N, dtype=np.int32), n_act)
    t = np.concatenate([np.sort(rng.choice(200, k, replace=False)) for k in n_act]).astype(np.int32)
    _, qf, qt = _edges(rng, agent.size, d, N, spread=0.35)
    args = (agent, qf, qt, agent, qf, qt, "first")
    kw = dict(col_group=agent, n_groups=N, row_key=t, col_key=t)
    g = ctx.collide_cross_reduce(*args, **kw)
    o = orc.collide_cross_reduce(*args, **kw, nthreads=0)
    assert g.shape == (agent.size, N) and np.array_equal(g, o)
    assert (o >= 0).any() and (o == -1).any()
    # without keys: first colliding action of every other agent
    g = ctx.collide_cross_reduce(*args, col_group=agent, n_groups=N)
    o = orc.collide_cross_reduce(*args, col_group=agent, n_groups=N, nthreads=0)
    assert np.array_equal(g, o)
    # empty column set
    e = ctx.collide_cross_reduce(agent, qf, qt, agent[:0], qf[:0], qt[:0], "first", col_group=agent[:0],
                                 n_groups=N)
    assert (e == -1).all()


@pytest.mark.parametrize("cbs", [False, True])
def test_arm22_count_and_first_match_oracle(cbs):
    N, M = 6, 2
    rads, obs, base = make_arm_instance(N, M, seed=4, link=(0.15, 0.25))
    ctx = S.Context(S.MODEL_ARM22, rads, obs, base_pos=base)
    orc = O.Oracle(O.ARM22, rads, obs, base=base)
    rng = np.random.default_rng(8)
    n_rows, T = 24, 5
    ra = rng.integers(0, N, n_rows).astype(np.int32)
    rf = rng.uniform(0, 2 * np.pi, (n_rows, 2))
    rt = rf + rng.normal(scale=0.4, size=(n_rows, 2))
    ca = np.tile(np.arange(N, dtype=np.int32), T)
    cf = rng.uniform(0, 2 * np.pi, (N * T, 2))
    ct = cf + rng.normal(scale=0.4, size=(N * T, 2))
    g = ctx.collide_cross_reduce(ra, rf, rt, ca, cf, ct, "count", group_size=N, cbs=cbs)
    o = orc.collide_cross_reduce(ra, rf, rt, ca, cf, ct, "count", group_size=N, cbs=cbs, nthreads=0)
    assert np.array_equal(g, o) and o.any()
    key_r = rng.integers(0, T, n_rows).astype(np.int32)
    key_c = np.repeat(np.arange(T, dtype=np.int32), N)
    kw = dict(col_group=ca, n_groups=N, row_key=key_r, col_key=key_c)
    g = ctx.collide_cross_reduce(ra, rf, rt, ca, cf, ct, "first", **kw)
    o = orc.collide_cross_reduce(ra, rf, rt, ca, cf, ct, "first", **kw, nthreads=0)
    assert np.array_equal(g, o) and (o >= 0).any()


def test_chain_model_count_matches_oracle():
    model = O.LINE2D
    inst = make_chain_instance(model, 5, 1, seed=3)
    ctx = S.Context(model, inst["rads"], inst["obs"], base_pos=inst["base"], aux=inst["aux"])
    orc = O.Oracle(model, inst["rads"], inst["obs"], base=inst["base"], aux=inst["aux"])
    rng = np.random.default_rng(13)
    n_rows, T, N = 20, 4, 5
    ra = rng.integers(0, N, n_rows).astype(np.int32)
    rf = chain_states(rng, model, n_rows)
    rt = chain_states(rng, model, n_rows, near=rf)
    ca = np.tile(np.arange(N, dtype=np.int32), T)
    cf = chain_states(rng, model, N * T)
    ct = chain_states(rng, model, N * T, near=cf)
    for cbs in (False, True):
        g = ctx.collide_cross_reduce(ra, rf, rt, ca, cf, ct, "count", group_size=N, cbs=cbs)
        o = orc.collide_cross_reduce(ra, rf, rt, ca, cf, ct, "count", group_size=N, cbs=cbs, nthreads=0)
        assert np.array_equal(g, o)
    assert o.any()


def test_reduce_argument_errors():
    rads = np.array([0.02, 0.03])
    ctx = S.Context(O.POINT2D, rads, np.zeros((0, 3)))
    q = np.zeros((2, 2))
    a = np.array([0, 1], np.int32)
    with pytest.raises(S.MRMPError):   # keys must come in pairs
        ctx.collide_cross_reduce(a, q, q, a, q, q, "first", group_size=1, row_key=a)
    with pytest.raises(S.MRMPError):   # group id beyond n_groups
        ctx.collide_cross_reduce(a, q, q, a, q, q, "count", col_group=np.array([0, 5], np.int32), n_groups=2)
    with pytest.raises(S.MRMPError):   # n_groups too small for group_size
        ctx.collide_cross_reduce(a, q, q, a, q, q, "count", group_size=1, n_groups=1)
