"""GPU parity of the cross-product collide kernels (broad phase + compaction + exact path):
bit matrices must equal the exhaustive CPU oracle -- the broad phase may never change a
verdict."""
import numpy as np
import pytest

import sssp_b200 as S
from oracle import oracle as O
from helpers import make_point_instance, random_edges

pytestmark = pytest.mark.gpu


def _edges(rng, n, d, N, spread=1.0, max_len=0.1):
    a = rng.integers(0, N, n).astype(np.int32)
    qf, qt = random_edges(rng, n, d, max_len)
    return a, qf * spread + (1 - spread) * 0.5, qt * spread + (1 - spread) * 0.5


@pytest.mark.parametrize("model,d", [(O.POINT2D, 2), (O.POINT3D, 3)])
@pytest.mark.parametrize("n_rows,n_cols,gs", [(1000, 700, 0), (513, 64, 0), (300, 1000, 50),
                                               (257, 6000, 7), (1, 1, 0), (2000, 3, 1)])
def test_cross_matches_exhaustive_oracle(model, d, n_rows, n_cols, gs):
    N = 12
    rads, obs = make_point_instance(model, N, 0, seed=n_rows + n_cols, rad=(0.015, 0.05))
    ctx, orc = S.Context(model, rads, obs), O.Oracle(model, rads, obs)
    rng = np.random.default_rng(n_rows * 7 + n_cols)
    # a dense cluster so that a few percent of the pairs truly collide
    ra, rf, rt = _edges(rng, n_rows, d, N, spread=0.5)
    ca, cf, ct = _edges(rng, n_cols, d, N, spread=0.5)
    g = ctx.collide_cross(ra, rf, rt, ca, cf, ct, gs)
    o = orc.collide_cross(ra, rf, rt, ca, cf, ct, gs, nthreads=0)
    assert g.shape == o.shape and np.array_equal(g, o)
    if n_rows * n_cols > 10000:
        assert o.any()


def test_cross_degenerate_and_near_threshold():
    """Stay edges (zero length), parallel edges, edges exactly at the contact threshold and
    edges just outside the broad-phase radius."""
    rads = np.array([0.02, 0.03, 0.025])
    ctx, orc = S.Context(O.POINT2D, rads, np.zeros((0, 3))), O.Oracle(O.POINT2D, rads, np.zeros((0, 3)))
    rng = np.random.default_rng(5)
    n = 400
    ra = rng.integers(0, 3, n).astype(np.int32)
    rf = rng.uniform(0.4, 0.6, (n, 2))
    rt = rf.copy()
    rt[::2] += rng.normal(scale=0.02, size=(n // 2, 2))          # half stay, half move
    ca = rng.integers(0, 3, n).astype(np.int32)
    thr = rads[ra] + rads[ca]
    eps = rng.choice([0.0, 1e-16, -1e-16, 1e-13, -1e-13, 1e-9, -1e-9], n)
    cf = rf + np.stack([thr * (1 + eps), np.zeros(n)], axis=1)   # column k grazes row k
    ct = cf + (rt - rf)                                          # parallel to the row edge
    g = ctx.collide_cross(ra, rf, rt, ca, cf, ct, 0)
    o = orc.collide_cross(ra, rf, rt, ca, cf, ct, 0, nthreads=0)
    assert np.array_equal(g, o)
    diag = (o[np.arange(n), np.arange(n) // 32] >> (np.arange(n) % 32)) & 1
    assert 0 < diag.sum() < n


def test_roadmap_edge_conflict_table():
    """All roadmap edges x path edges of the other agents, OR-reduced per time step: the
    table PP consults (pp.jl:179-189)."""
    N, nv, T = 6, 150, 9
    rads, obs = make_point_instance(O.POINT2D, N, 5, seed=2, rad=(0.02, 0.04))
    ctx, orc = S.Context(O.POINT2D, rads, obs), O.Oracle(O.POINT2D, rads, obs)
    rng = np.random.default_rng(3)
    nodes = rng.random((N, nv, 2))
    rm = S.Roadmaps(ctx, 0, N, nv)
    rm.set_nodes(nodes)
    nnz = rm.build(0.2)
    row_ptr, col = rm.csr()
    # columns: for every time step t, one path edge per agent (t-major, group = N columns)
    ca = np.tile(np.arange(N, dtype=np.int32), T)
    vf = rng.integers(0, nv, N * T).astype(np.int32)
    vt = rng.integers(0, nv, N * T).astype(np.int32)
    # explicit row arrays in packed CSR order
    rows = np.repeat(np.arange(N * nv), np.diff(row_ptr))
    ra = (rows // nv).astype(np.int32)
    rf = nodes[ra, rows % nv]
    rt = nodes[ra, col]
    cf, ct = nodes[ca, vf], nodes[ca, vt]
    for gs in (0, N):
        g = rm.edge_conflicts(ca, vf, vt, gs)
        o = orc.collide_cross(ra, rf, rt, ca, cf, ct, gs, nthreads=0)
        assert g.shape == (nnz, o.shape[1]) and np.array_equal(g, o)
    assert o.any()


def test_full_size_bench_step_matches_the_oracle_bit_for_bit():
    """BASELINE configs[1] at full size -- point2d_many N=50, PRMs(500, rad=0.1), all 396 166
    roadmap edges x 64 time steps of the other agents' path edges (1.24e9 collide checks) --
    the whole step of bench.py: roadmaps (vertices + CSR) and the OR-reduced conflict table
    are compared in full with the CPU restatement (about 10 s on the host cores)."""
    from sssp_b200 import workload as W
    T = 64
    inst = W.point2d_many(50, seed=0, nv=500, rad=0.1)
    N, nv = 50, inst.nv
    ctx = S.Context(inst.model, inst.rads, inst.obstacles)
    orc = O.Oracle(O.POINT2D, inst.rads, inst.obstacles)
    rm = S.Roadmaps(ctx, 0, N, nv)
    rm.sample(nv, inst.seed, inst.q_init, inst.q_goal, tries=False)
    nnz = rm.build(inst.rad)
    nodes, (row_ptr, col) = rm.nodes(), rm.csr()
    # roadmaps: the oracle draws the same counter-based stream and runs the all-pairs pass
    o_nodes = np.empty_like(nodes)
    for a in range(N):
        o_nodes[a, 0], o_nodes[a, 1] = inst.q_init[a], inst.q_goal[a]
        o_nodes[a, 2:] = orc.sample_free(a, inst.seed, nv - 2)[0]
    assert np.array_equal(nodes, o_nodes)
    orp, ocol, onnz = orc.prm_build(0, o_nodes, inst.rad, nthreads=0)
    assert nnz == int(onnz.sum()) == 396166
    for a in range(N):
        lo, hi = row_ptr[a * nv], row_ptr[(a + 1) * nv]
        assert np.array_equal(col[lo:hi], ocol[a, :onnz[a]]), a
        assert np.array_equal(row_ptr[a * nv:(a + 1) * nv + 1] - lo, orp[a]), a
    # conflict table
    _, ca, vf, vt = W.random_walk_paths(inst, row_ptr, col, T, seed=4321)
    got = rm.edge_conflicts(ca, vf, vt, group_size=N)
    rows = np.repeat(np.arange(N * nv), np.diff(row_ptr))
    ra = (rows // nv).astype(np.int32)
    rf, rt = nodes[ra, rows % nv], nodes[ra, col]
    want = orc.collide_cross(ra, rf, rt, ca, nodes[ca, vf], nodes[ca, vt], N, nthreads=0)
    assert got.shape == want.shape == (nnz, 2)
    assert np.array_equal(got, want)
    density = np.unpackbits(want.view(np.uint8)).mean()
    assert 0.05 < density < 0.95


def test_point3d_config_roadmaps_and_conflict_table_match_the_oracle():
    """BASELINE configs[2]: StatePoint3D N=20 with sphere obstacles, PRMs(200, rad=0.3)
    (point3d.yaml:26-30; radii per SURVEY 8d), every roadmap edge x 32 time steps of the other
    agents' random-walk path edges, compared in full."""
    from sssp_b200 import workload as W
    N, nv, T, rad = 20, 200, 32, 0.3
    rng = np.random.default_rng(5)
    obs = np.concatenate([rng.random((10, 3)), rng.uniform(0.05, 0.1, (10, 1))], axis=1)
    rads = rng.uniform(0.04, 0.05, N)
    ctx, orc = S.Context(O.POINT3D, rads, obs), O.Oracle(O.POINT3D, rads, obs)
    ends = np.stack([orc.sample_free(a, 99, 2)[0] for a in range(N)])
    rm = S.Roadmaps(ctx, 0, N, nv)
    rm.sample(nv, 7, ends[:, 0], ends[:, 1], tries=False)
    nnz = rm.build(rad)
    nodes, (row_ptr, col) = rm.nodes(), rm.csr()
    o_nodes = np.empty_like(nodes)
    for a in range(N):
        o_nodes[a, :2] = ends[a]
        o_nodes[a, 2:] = orc.sample_free(a, 7, nv - 2)[0]
    assert np.array_equal(nodes, o_nodes)
    orp, ocol, onnz = orc.prm_build(0, o_nodes, rad, nthreads=0)
    assert nnz == int(onnz.sum()) and nnz > 10 * N * nv
    for a in range(N):
        lo, hi = row_ptr[a * nv], row_ptr[(a + 1) * nv]
        assert np.array_equal(col[lo:hi], ocol[a, :onnz[a]]), a
    inst = W.Instance(O.POINT3D, rads, obs, ends[:, 0], ends[:, 1], nv, rad, 7)
    _, ca, vf, vt = W.random_walk_paths(inst, row_ptr, col, T, seed=11)
    rows = np.repeat(np.arange(N * nv), np.diff(row_ptr))
    ra = (rows // nv).astype(np.int32)
    rf, rt = nodes[ra, rows % nv], nodes[ra, col]
    for gs in (N, 0):
        got = rm.edge_conflicts(ca, vf, vt, group_size=gs)
        want = orc.collide_cross(ra, rf, rt, ca, nodes[ca, vf], nodes[ca, vt], gs, nthreads=0)
        assert np.array_equal(got, want), gs
    assert want.any()
