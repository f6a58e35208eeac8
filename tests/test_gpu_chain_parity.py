"""GPU parity for line2d, snake2d, arm33 and capsule3d (src/models/*.jl): every C-ABI entry
point against the CPU oracle (oracle/orc_chain.c) on the same seeded inputs -- verdicts,
samples, roadmap edge sets and conflict tables bit-exact."""
import math

import numpy as np
import pytest

import sssp_b200 as S
from sssp_b200 import lib as L
from oracle import oracle as O
from helpers import CHAIN_MODELS, chain_states, make_chain_instance

pytestmark = pytest.mark.gpu
NAMES = {O.LINE2D: "line2d", O.SNAKE2D: "snake2d", O.ARM33: "arm33", O.CAPSULE3D: "capsule3d"}


def _pair(model, N, M, seed, **kw):
    ins = make_chain_instance(model, N, M, seed)
    ctx = S.Context(model, ins["rads"], ins["obs"], base_pos=ins["base"], aux=ins["aux"], **kw)
    okw = {k: v for k, v in kw.items() if v is not None}
    orc = O.Oracle(model, ins["rads"], ins["obs"], base=ins["base"], aux=ins["aux"], **okw)
    return ctx, orc, ins


def _moves(rng, model, n, spread=0.12):
    q1 = chain_states(rng, model, n)
    q2 = chain_states(rng, model, n, near=q1, spread=spread)
    q2[:16] = q1[:16]                                     # D == 0: NaN first step
    q2[16:48] = chain_states(rng, model, 32)              # long moves (many postures)
    return q1, q2


@pytest.mark.parametrize("model", CHAIN_MODELS, ids=NAMES.values())
@pytest.mark.parametrize("max_dist", [None, 0.5])
def test_connect_parity(model, max_dist):
    N, M = 8, 3
    ctx, orc, ins = _pair(model, N, M, seed=model, max_dist=max_dist)
    rng = np.random.default_rng(60 + model)
    n = 6000
    ag = rng.integers(0, N, n).astype(np.int32)
    q1, q2 = _moves(rng, model, n)
    g, o = ctx.connect_points(ag, q1), orc.connect_points(ag, q1, nthreads=0)
    assert np.array_equal(g, o) and 0 < o.sum() < n
    g, o = ctx.connect_edges(ag, q1, q2), orc.connect_edges(ag, q1, q2, nthreads=0)
    assert np.array_equal(g, o) and 0 < o.sum() < n
    # single calls and the empty batch
    assert ctx.connect_edges(ag[:1], q1[:1], q2[:1])[0] == orc.connect_edge(q1[0], q2[0], ag[0])
    assert ctx.connect_points(ag[:0], q1[:0]).size == 0


@pytest.mark.parametrize("model", CHAIN_MODELS, ids=NAMES.values())
def test_collide_parity(model):
    N = 8
    ctx, orc, ins = _pair(model, N, 0, seed=model)
    rng = np.random.default_rng(70 + model)
    n = 1500
    ai = rng.integers(0, N, n).astype(np.int32)
    aj = ((ai + 1 + rng.integers(0, N - 1, n)) % N).astype(np.int32)
    a, b = _moves(rng, model, n)
    c = chain_states(rng, model, n, near=a, spread=0.1)
    e = chain_states(rng, model, n, near=c, spread=0.12)
    g, o = ctx.collide_pairs(ai, aj, a, b, c, e), orc.collide_pairs(ai, aj, a, b, c, e, nthreads=0)
    assert np.array_equal(g, o) and 0 < o.sum() < n
    g, o = ctx.collide_static_pairs(ai, aj, a, c), orc.collide_static_pairs(ai, aj, a, c, nthreads=0)
    assert np.array_equal(g, o) and 0 < o.sum() < n
    # two stay moves == the static test
    assert np.array_equal(ctx.collide_pairs(ai, aj, a, a, c, c), o)
    assert ctx.stats()[1] > 0
    # configurations: verdict + first colliding pair; candidate moves of one agent
    n_cfg = 120
    Qf = chain_states(rng, model, n_cfg * N).reshape(n_cfg, N, -1)
    Qt = chain_states(rng, model, n_cfg * N, near=Qf.reshape(n_cfg * N, -1), spread=0.1).reshape(Qf.shape)
    g, gfp = ctx.collide_configs(Qf, Qt)
    o, ofp = orc.collide_configs_batch(Qf, Qt, nthreads=0)
    assert np.array_equal(g, o) and np.array_equal(gfp, ofp) and 0 < o.sum() < n_cfg
    for i in (0, N - 1):
        cand = chain_states(rng, model, 300, near=np.repeat(Qf[3, i][None], 300, 0), spread=0.15)
        g = ctx.collide_config_moves(i, Qf[3], cand)
        o = orc.collide_config_moves(i, Qf[3], cand, nthreads=0)
        assert np.array_equal(g, o)


@pytest.mark.parametrize("model", CHAIN_MODELS, ids=NAMES.values())
def test_roadmaps_cross_and_helpers_parity(model):
    N, M, nv = 4, 2, 80
    ctx, orc, ins = _pair(model, N, M, seed=model)
    rng = np.random.default_rng(80 + model)
    d = ctx.d
    # sampler stream + free-space sampling in try order
    g = ctx.sample_states(agent=1, seed=5, t0=3, n=500)
    o = np.stack([orc.sample_state(5, 1, 3 + t) for t in range(500)])
    assert np.array_equal(g, o)
    init = np.stack([orc.sample_free(a, 900, 1)[0][0] for a in range(N)])
    goal = np.stack([orc.sample_free(a, 901, 1)[0][0] for a in range(N)])
    rm = S.Roadmaps(ctx, 0, N, nv)
    tries = rm.sample(nv, 77, init, goal)
    nodes = rm.nodes()
    for a in range(N):
        exp, t = orc.sample_free(a, 77, nv - 2)
        assert np.array_equal(nodes[a, 2:], exp) and tries[a] == t
    rad = {O.LINE2D: 0.25, O.SNAKE2D: 0.6, O.ARM33: 0.8, O.CAPSULE3D: 0.6}[model]
    nnz = rm.build(rad)
    row_ptr, col = rm.csr()
    orp, ocol, onnz = orc.prm_build(0, nodes, rad, nthreads=0)
    assert nnz == int(onnz.sum()) and nnz > 0
    for a in range(N):
        for v in range(nv):
            r = a * nv + v
            assert np.array_equal(col[row_ptr[r]:row_ptr[r + 1]], ocol[a, orp[a, v]:orp[a, v + 1]])
    # collide over roadmap edges by id, cross products, conflict table by ids
    rows = np.repeat(np.arange(N * nv), np.diff(row_ptr))
    n = 600
    ei, ej = rng.integers(0, nnz, n), rng.integers(0, nnz, n)
    ai, aj = (rows[ei] // nv).astype(np.int32), (rows[ej] // nv).astype(np.int32)
    keep = ai != aj
    ai, aj, ei, ej = ai[keep], aj[keep], ei[keep], ej[keep]
    args = (ai, aj, (rows[ei] % nv).astype(np.int32), col[ei], (rows[ej] % nv).astype(np.int32), col[ej])
    g = rm.collide_edges_indexed(*args)
    o = orc.collide_edges_indexed(*args, nodes, nv, nthreads=0)
    assert np.array_equal(g, o)
    ra, ca = ai[:24], aj[:64]
    rf, rt = nodes[ra, args[2][:24]], nodes[ra, args[3][:24]]
    cf, ct = nodes[ca, args[4][:64]], nodes[ca, args[5][:64]]
    for gs in (0, 8):
        assert np.array_equal(ctx.collide_cross(ra, rf, rt, ca, cf, ct, gs),
                              orc.collide_cross(ra, rf, rt, ca, cf, ct, gs, nthreads=0))
    g = rm.edge_conflicts(ca, args[4][:64], args[5][:64], group_size=8)
    o = orc.collide_cross((rows // nv).astype(np.int32)[::40], nodes[rows // nv, rows % nv][::40],
                          nodes[rows // nv, col][::40], ca, cf, ct, 8, nthreads=0)
    assert np.array_equal(g[::40], o)
    # model helpers
    a, b = chain_states(rng, model, 300), chain_states(rng, model, 300)
    dist, mid = ctx.state_dist(a, b), ctx.mid_status(a, b)
    for k in range(300):
        assert dist[k].view(np.uint64) == np.float64(orc.state_dist(a[k], b[k])).view(np.uint64)
        assert np.array_equal(mid[k].view(np.uint64), orc.mid_status(a[k], b[k]).view(np.uint64))
    idx, dd = ctx.nearest(a, b[0])
    ds = [orc.state_dist(v, b[0]) for v in a]
    assert idx == int(np.argmin(ds)) and dd == min(ds)


def test_kats_through_the_closures():
    # derived by hand from line2d.jl:30-68: link of length 0.2 from (x, y) along theta
    Ln = S.StateLine2D
    connect = S.gen_connect(Ln(0, 0, 0), [S.CircleObstacle2D(0.5, 0.5, 0.1)], [0.2])
    assert connect(Ln(0.3, 0.5, 0.0), 0) is False          # tip reaches the obstacle centre
    assert connect(Ln(0.3, 0.2, 0.0), 0) is True           # root exactly at rads[i]: not < rads[i]
    assert connect(Ln(0.1, 0.2, 0.0), 0) is False          # point form keeps the rads[i] margin ...
    assert connect(Ln(0.1, 0.2, 0.0), Ln(0.15, 0.2, 0.0), 0) is True   # ... the edge loop does not
    collide = S.gen_collide(Ln(0, 0, 0), [0.2, 0.2])
    assert collide(Ln(0.3, 0.5, 0.0), Ln(0.4, 0.4, math.pi / 2), 0, 1) is True     # crossing links
    assert collide(Ln(0.3, 0.5, 0.0), Ln(0.4, 0.6, math.pi / 2), 0, 1) is False
    # capsule3d.jl:70-79: capsule radius enters the obstacle and pair thresholds
    Cp = S.StateCapsule3D
    connect = S.gen_connect(Cp(0, 0, 0, 0, 0, 0), [S.CircleObstacle3D(0.5, 0.5, 0.5, 0.1)], [0.05], [0.2])
    assert connect(Cp(0.2, 0.5, 0.45, 0, 0, 0), 0) is False     # tip (0.4, 0.5, 0.45): 0.112 < 0.1 + 0.05
    assert connect(Cp(0.2, 0.5, 0.36, 0, 0, 0), 0) is True      # tip (0.4, 0.5, 0.36): 0.172 >= 0.15
