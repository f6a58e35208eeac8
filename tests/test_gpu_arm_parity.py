"""GPU parity for the StateArm22 model (src/models/arm22.jl): every C-ABI entry point
against the CPU oracle on the same seeded inputs.  The device trig is the same msun
restatement as the oracle's (trig.cuh vs oracle/orc_trig.c), so verdicts, samples and edge
sets must be BIT-EXACT: zero mismatches."""
import math

import numpy as np
import pytest

import sssp_b200 as S
from sssp_b200 import lib as L
from oracle import oracle as O
from helpers import make_arm_instance

pytestmark = pytest.mark.gpu
TWO_PI = 2 * math.pi


# BASELINE configs[3] / scripts/config/eval/arm22.yaml:11-20: link length 0.15-0.25, up to 3
# obstacles of radius 0.05-0.10 (the instance bench.py measures)
LINK = (0.15, 0.25)


def _pair(N, M, seed, link=LINK, **kw):
    rads, obs, base = make_arm_instance(N, M, seed, link=link)
    ctx = S.Context(L.MODEL_ARM22, rads, obs, base_pos=base, **kw)
    okw = dict(kw)
    if okw.get("max_dist") is None:
        okw.pop("max_dist", None)
    orc = O.Oracle(O.ARM22, rads, obs, base=base, **okw)
    return ctx, orc, rads, obs, base


def _moves(rng, n):
    q1 = rng.uniform(0, TWO_PI, (n, 2))
    q2 = rng.uniform(0, TWO_PI, (n, 2))
    k = n // 4
    q2[:k] = q1[:k] + rng.normal(size=(k, 2)) * 0.05     # short moves (few steps)
    q2[k:k + 16] = q1[k:k + 16]                           # D == 0: NaN first step
    q2[k + 16:k + 32] = q1[k + 16:k + 32] + np.array([0.29 * math.pi, 0.0])   # decimal-looking D
    return q1, q2


def test_kats_through_the_closures():
    A = S.StateArm22
    base, rads = [[0.5, 0.5], [0.5, 0.8]], [0.2, 0.2]
    connect = S.gen_connect(A(0, 0), [S.CircleObstacle2D(0.75, 0.5, 0.06)], base, rads)
    assert connect(A(math.pi / 2, 0.0), 0) is True            # up, then right: free
    assert connect(A(0.0, math.pi / 2), 0) is False           # first link runs into the obstacle
    assert connect(A(math.pi / 2, math.pi / 2), 0) is True    # straight up to y = 0.9
    assert connect(A(math.pi / 2, math.pi / 2), 1) is False   # tip leaves the unit square
    assert connect(A(math.pi, 0.0), 0) is False               # folded back: self collision
    # q_from == q_to must equal the static test (NaN-step rule, arm22.jl:65)
    for q in (A(math.pi / 2, 0.0), A(0.0, math.pi / 2)):
        assert connect(q, q, 0) is connect(q, 0)
    collide = S.gen_collide(A(0, 0), base, rads)
    up, down = A(math.pi / 2, math.pi / 2), A(-math.pi / 2, -math.pi / 2)
    assert collide(up, down, 0, 1) is True                    # arm 0 up meets arm 1 down
    assert collide(down, up, 0, 1) is False
    assert collide(down, up, down, down, 0, 1) is True        # sweeping up into arm 1
    Q = [up, down]
    assert collide(Q, Q) is True and collide([down, up], [down, up]) is False
    assert collide([down, down], up, 0) is True and collide([down, up], A(0.0, 0.0), 0) is False


@pytest.mark.parametrize("max_dist", [None, 0.6])
def test_connect_parity(max_dist):
    N, M = 10, 3
    ctx, orc, rads, obs, base = _pair(N, M, seed=11, max_dist=max_dist)
    rng = np.random.default_rng(21)
    n = 20000
    ag = rng.integers(0, N, n).astype(np.int32)
    q1, q2 = _moves(rng, n)
    g, o = ctx.connect_points(ag, q1), orc.connect_points(ag, q1, nthreads=0)
    assert np.array_equal(g, o) and 0 < o.sum() < n
    g, o = ctx.connect_edges(ag, q1, q2), orc.connect_edges(ag, q1, q2, nthreads=0)
    assert np.array_equal(g, o) and 0 < o.sum() < n
    # edges between free postures (what PRM actually tests)
    free = np.flatnonzero(orc.connect_points(ag, q1, nthreads=0))
    qa = q1[free]
    qb = qa + rng.normal(size=qa.shape) * 0.15
    g, o = ctx.connect_edges(ag[free], qa, qb), orc.connect_edges(ag[free], qa, qb, nthreads=0)
    assert np.array_equal(g, o) and 0 < o.sum() < free.size


def test_connect_parity_fine_steps_and_single_calls():
    # step_dist 0.003: up to ~470 postures per edge; 0.05 exercises another rational step
    for step in (0.003, 0.05):
        ctx, orc, rads, obs, base = _pair(6, 2, seed=12, step_dist=step)
        rng = np.random.default_rng(22)
        n = 3000
        ag = rng.integers(0, 6, n).astype(np.int32)
        q1, q2 = _moves(rng, n)
        assert np.array_equal(ctx.connect_edges(ag, q1, q2), orc.connect_edges(ag, q1, q2, nthreads=0))
    assert ctx.connect_edges(ag[:1], q1[:1], q2[:1])[0] == orc.connect_edge(q1[0], q2[0], ag[0])
    assert ctx.connect_edges(ag[:0], q1[:0], q2[:0]).size == 0


def test_collide_pairs_parity():
    N = 10
    ctx, orc, rads, obs, base = _pair(N, 3, seed=13)
    rng = np.random.default_rng(23)
    n = 3000
    ai = rng.integers(0, N, n).astype(np.int32)
    aj = ((ai + 1 + rng.integers(0, N - 1, n)) % N).astype(np.int32)
    a, b = _moves(rng, n)
    c, e = _moves(rng, n)
    g, o = ctx.collide_pairs(ai, aj, a, b, c, e), orc.collide_pairs(ai, aj, a, b, c, e, nthreads=0)
    assert np.array_equal(g, o) and 0 < o.sum() < n
    g, o = ctx.collide_static_pairs(ai, aj, a, c), orc.collide_static_pairs(ai, aj, a, c, nthreads=0)
    assert np.array_equal(g, o) and 0 < o.sum() < n
    # stay actions on both sides (D == 0 twice) must equal the static test
    g = ctx.collide_pairs(ai, aj, a, a, c, c)
    assert np.array_equal(g, orc.collide_pairs(ai, aj, a, a, c, c, nthreads=0))
    assert np.array_equal(g, o)
    assert ctx.stats()[1] > 0     # the exact phase ran on the GPU (arm link-pair counter)


def test_collide_pairs_parity_fine_steps():
    # step_dist 0.004: schedules longer than one 128-posture table tile on both sides
    N = 6
    ctx, orc, rads, obs, base = _pair(N, 0, seed=14, step_dist=0.004)
    rng = np.random.default_rng(24)
    n = 400
    ai = rng.integers(0, N, n).astype(np.int32)
    aj = ((ai + 1 + rng.integers(0, N - 1, n)) % N).astype(np.int32)
    a = rng.uniform(0, TWO_PI, (n, 2))
    b = rng.uniform(0, TWO_PI, (n, 2))
    c = rng.uniform(0, TWO_PI, (n, 2))
    e = rng.uniform(0, TWO_PI, (n, 2))
    g, o = ctx.collide_pairs(ai, aj, a, b, c, e), orc.collide_pairs(ai, aj, a, b, c, e, nthreads=0)
    assert np.array_equal(g, o) and 0 < o.sum() < n


def test_collide_configs_and_moves_parity():
    N = 6
    rads, obs, base = make_arm_instance(N, 0, 18, link=(0.07, 0.1))   # sparse enough for both verdicts
    ctx = S.Context(L.MODEL_ARM22, rads, obs, base_pos=base)
    orc = O.Oracle(O.ARM22, rads, obs, base=base)
    rng = np.random.default_rng(25)
    n_cfg = 300
    Qf = rng.uniform(0, TWO_PI, (n_cfg, N, 2))
    Qt = Qf + rng.normal(size=Qf.shape) * 0.2
    g, gfp = ctx.collide_configs(Qf, Qt)
    o, ofp = orc.collide_configs_batch(Qf, Qt, nthreads=0)
    assert np.array_equal(g, o) and np.array_equal(gfp, ofp) and 0 < o.sum() < n_cfg
    assert np.unique(ofp).size > 5          # first hits at several different pairs
    for i in (0, 3, N - 1):
        cand = rng.uniform(0, TWO_PI, (500, 2))
        cand[:50] = Qf[1, i] + rng.normal(size=(50, 2)) * 0.05
        g = ctx.collide_config_moves(i, Qf[1], cand)
        o = orc.collide_config_moves(i, Qf[1], cand, nthreads=0)
        assert np.array_equal(g, o) and 0 < o.sum() < 500


def test_roadmaps_and_cross_parity():
    # arm22.yaml: PRM num_vertices 100, rad 0.3 (scripts/config/eval/arm22.yaml:26-30)
    N, M, nv, rad = 6, 2, 100, 0.3
    ctx, orc, rads, obs, base = _pair(N, M, seed=17, link=(0.1, 0.15))   # every agent has free postures
    rng = np.random.default_rng(26)
    init = np.stack([orc.sample_free(a, 900, 1)[0][0] for a in range(N)])
    goal = np.stack([orc.sample_free(a, 901, 1)[0][0] for a in range(N)])
    rm = S.Roadmaps(ctx, 0, N, nv)
    tries = rm.sample(nv, 77, init, goal)
    nodes = rm.nodes()
    for a in range(N):
        exp, t = orc.sample_free(a, 77, nv - 2)
        assert np.array_equal(nodes[a, 2:], exp) and tries[a] == t
        assert np.array_equal(nodes[a, 0], init[a]) and np.array_equal(nodes[a, 1], goal[a])
    nnz = rm.build(rad)
    row_ptr, col = rm.csr()
    orp, ocol, onnz = orc.prm_build(0, nodes, rad, nthreads=0)
    assert nnz == int(onnz.sum()) and nnz > 0
    for a in range(N):
        for v in range(nv):
            r = a * nv + v
            assert np.array_equal(col[row_ptr[r]:row_ptr[r + 1]], ocol[a, orp[a, v]:orp[a, v + 1]])
    # collide over roadmap edges by id
    n = 1500
    rows = np.repeat(np.arange(N * nv), np.diff(row_ptr))
    ei = rng.integers(0, nnz, n)
    ej = rng.integers(0, nnz, n)
    ai, aj = (rows[ei] // nv).astype(np.int32), (rows[ej] // nv).astype(np.int32)
    keep = ai != aj
    ai, aj, ei, ej = ai[keep], aj[keep], ei[keep], ej[keep]
    args = (ai, aj, (rows[ei] % nv).astype(np.int32), col[ei], (rows[ej] % nv).astype(np.int32), col[ej])
    g = rm.collide_edges_indexed(*args)
    o = orc.collide_edges_indexed(*args, nodes, nv, nthreads=0)
    assert np.array_equal(g, o) and 0 < o.sum() < ai.size
    # cross product: 40 row edges x 96 column edges, full matrix and OR-reduced groups of 8
    ra, ca = ai[:40], aj[:96]
    rf, rt = nodes[ra, args[2][:40]], nodes[ra, args[3][:40]]
    cf, ct = nodes[ca, args[4][:96]], nodes[ca, args[5][:96]]
    for gs in (0, 8):
        g = ctx.collide_cross(ra, rf, rt, ca, cf, ct, gs)
        o = orc.collide_cross(ra, rf, rt, ca, cf, ct, gs, nthreads=0)
        assert np.array_equal(g, o) and o.any()
    # conflict table by ids against the roadmap set itself
    g = rm.edge_conflicts(ca, args[4][:96], args[5][:96], group_size=8)
    rows_f = nodes[rows // nv, rows % nv]
    rows_t = nodes[rows // nv, col]
    o = orc.collide_cross((rows // nv).astype(np.int32)[::50], rows_f[::50], rows_t[::50], ca, cf, ct, 8,
                          nthreads=0)
    assert np.array_equal(g[::50], o)


def test_config4_conflict_table_full_parity():
    """The whole conflict table of the BASELINE configs[3] instance (N=10, M=3, links 0.15-0.25):
    every roadmap edge (PP's num_vertices / rad of arm22.yaml:26-30, smaller vertex count) against
    the path edges of the other agents over T steps, OR-reduced per step -- every row, not a sample."""
    N, M, nv, rad, T = 10, 3, 40, 0.3, 3
    ctx, orc, rads, obs, base = _pair(N, M, seed=41)
    nodes = np.stack([orc.sample_free(a, 77, nv)[0] for a in range(N)])
    rm = S.Roadmaps(ctx, 0, N, nv)
    rm.set_nodes(nodes)
    nnz = rm.build(rad)
    row_ptr, col = rm.csr()
    orp, ocol, onnz = orc.prm_build(0, nodes, rad, nthreads=0)
    assert nnz == int(onnz.sum()) and nnz > 500
    rng = np.random.default_rng(1)
    ca = np.tile(np.arange(N, dtype=np.int32), T)
    vf = rng.integers(0, nv, N * T).astype(np.int32)
    vt = rng.integers(0, nv, N * T).astype(np.int32)
    rows = np.repeat(np.arange(N * nv), np.diff(row_ptr))
    ra = (rows // nv).astype(np.int32)
    g = rm.edge_conflicts(ca, vf, vt, group_size=N)
    o = orc.collide_cross(ra, nodes[ra, rows % nv], nodes[ra, col], ca, nodes[ca, vf], nodes[ca, vt], N,
                          nthreads=0)
    assert g.shape == o.shape == (nnz, 1) and np.array_equal(g, o)
    assert 0 < (o != 0).sum() and (o != 7).any()
