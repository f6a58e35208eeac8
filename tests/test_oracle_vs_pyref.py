"""Two independent restatements of the reference (C, scalarised; Python, literal vector
form) must agree bit for bit on random and near-threshold inputs."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import pyref as R
from helpers import make_point_instance, near_threshold_connect, near_threshold_pairs, random_edges


@pytest.mark.parametrize("d", [2, 3])
def test_primitives_bitwise(d):
    rng = np.random.default_rng(10 + d)
    for _ in range(3000):
        a, b, c, e = rng.random((4, d))
        if rng.random() < 0.2:
            b = a + (b - a) * 1e-3
        assert O.dist_pp(a, b) == R.dist_pp(list(a), list(b))
        assert O.dist_sp(a, b, c) == R.dist_sp(list(a), list(b), list(c))
        assert O.dist_ss(a, b, c, e) == R.dist_ss(list(a), list(b), list(c), list(e))


@pytest.mark.parametrize("model,d", [(O.POINT2D, 2), (O.POINT3D, 3)])
def test_closures_bitwise(model, d):
    rng = np.random.default_rng(20 + d)
    N, M = 6, 5
    rads, obs = make_point_instance(model, N, M, seed=5 + d, rad=(0.02, 0.06), rad_obs=(0.05, 0.1))
    o = O.Oracle(model, rads, obs)
    r = R.PointInstance(rads, obs)
    n = 1500
    ag = rng.integers(0, N, n)
    qf, qt = random_edges(rng, n, d, 0.3)
    nf, nt = near_threshold_connect(rng, n, rads, obs, ag)
    for k in range(n):
        i = int(ag[k])
        assert o.connect_point(qf[k], i) == r.connect_point(qf[k], i)
        assert o.connect_edge(qf[k], qt[k], i) == r.connect_edge(qf[k], qt[k], i)
        assert o.connect_edge(nf[k], nt[k], i) == r.connect_edge(nf[k], nt[k], i)
    aj = (ag + 1 + rng.integers(0, N - 1, n)) % N
    p = near_threshold_pairs(rng, n, rads, ag, aj, d)
    q2f, q2t = random_edges(rng, n, d, 0.3)
    hits = 0
    for k in range(n):
        i, j = int(ag[k]), int(aj[k])
        v = o.collide_pair(p[0][k], p[1][k], p[2][k], p[3][k], i, j)
        assert v == r.collide_pair(p[0][k], p[1][k], p[2][k], p[3][k], i, j)
        hits += v
        assert o.collide_pair(qf[k], qt[k], q2f[k], q2t[k], i, j) == \
            r.collide_pair(qf[k], qt[k], q2f[k], q2t[k], i, j)
        assert o.collide_static(qf[k], q2f[k], i, j) == r.collide_static(qf[k], q2f[k], i, j)
    assert 0 < hits < n  # the adversarial generator straddles the threshold
    for _ in range(200):
        Qf, Qt = random_edges(rng, N, d, 0.2)
        assert o.collide_configs(Qf, Qt)[0] == r.collide_configs(Qf, Qt)
        i = int(rng.integers(0, N))
        assert o.collide_config_move(Qf, Qt[i], i) == r.collide_config_move(Qf, Qt[i], i)


def test_prm_neighbors_bitwise():
    rng = np.random.default_rng(7)
    rads, obs = make_point_instance(O.POINT2D, 2, 6, seed=9, rad_obs=(0.05, 0.1))
    o = O.Oracle(O.POINT2D, rads, obs)
    r = R.PointInstance(rads, obs)
    nodes = rng.random((2, 60, 2))
    for rad in (0.25, None):
        row_ptr, col, nnz = o.prm_build(0, nodes, rad)
        for a in range(2):
            nb = r.prm_neighbors(nodes[a], rad, a)
            for v in range(60):
                assert col[a, row_ptr[a, v]:row_ptr[a, v + 1]].tolist() == nb[v]


def test_dot_fma_sensitivity_is_inside_1e9_band():
    """The 2-3 element `dot` goes to OpenBLAS ddot in the reference, whose tail may or may not
    fuse (hazard b).  Verdicts of the fused and un-fused oracles may only differ
    when the distance is within 1e-9 of the threshold."""
    rng = np.random.default_rng(11)
    N, M, d = 8, 10, 2
    rads, obs = make_point_instance(O.POINT2D, N, M, seed=3)
    a, b = O.Oracle(O.POINT2D, rads, obs), O.Oracle(O.POINT2D, rads, obs, dot_fma=True)
    n = 200000
    ag = rng.integers(0, N, n)
    nf, nt = near_threshold_connect(rng, n, rads, obs, ag)
    va, vb = a.connect_edges(ag, nf, nt, nthreads=0), b.connect_edges(ag, nf, nt, nthreads=0)
    bad = np.nonzero(va != vb)[0]
    for k in bad[:200]:
        dmin = min(abs(O.dist_sp(nf[k], nt[k], o[:2]) - (o[2] + rads[ag[k]])) for o in obs)
        assert dmin <= 1e-9
    assert bad.size < n * 0.05
