"""GPU parity for roadmap construction: sampler stream, free-space sampling order, and
bit-exact CSR edge sets (rows ascending) against the CPU oracle."""
import numpy as np
import pytest

import sssp_b200 as S
from oracle import oracle as O
from helpers import make_point_instance

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("model,d", [(O.POINT2D, 2), (O.POINT3D, 3)])
def test_sampler_stream_matches_oracle(model, d):
    rads, obs = make_point_instance(model, 4, 5, seed=1, rad_obs=(0.05, 0.1))
    ctx, orc = S.Context(model, rads, obs), O.Oracle(model, rads, obs)
    g = ctx.sample_states(agent=2, seed=46, t0=5, n=4096)
    o = np.stack([orc.sample_state(46, 2, 5 + t) for t in range(4096)])
    assert np.array_equal(g, o)
    assert 0.45 < g.mean() < 0.55 and g.min() >= 0 and g.max() < 1
    gq, gt = ctx.sample_free(agent=1, seed=7, n_wanted=700)
    oq, ot = orc.sample_free(1, 7, 700)
    assert gt == ot and np.array_equal(gq, oq)


@pytest.mark.parametrize("model,d,nv,rad", [(O.POINT2D, 2, 500, 0.1), (O.POINT3D, 3, 200, 0.3),
                                            (O.POINT2D, 2, 97, None), (O.POINT2D, 2, 2, 0.1)])
def test_prm_csr_bit_exact(model, d, nv, rad):
    N = 6
    rads, obs = make_point_instance(model, N, 10, seed=5 + nv)
    ctx, orc = S.Context(model, rads, obs), O.Oracle(model, rads, obs)
    rng = np.random.default_rng(nv)
    nodes = rng.random((N, nv, d)) * 0.9 + 0.05
    row_ptr, col = ctx.prm_build(nodes, rad)
    orp, ocol, onnz = orc.prm_build(0, nodes, rad, nthreads=0)
    assert row_ptr[-1] == onnz.sum() == col.size
    for a in range(N):
        for v in range(nv):
            r = a * nv + v
            assert np.array_equal(col[row_ptr[r]:row_ptr[r + 1]],
                                  ocol[a, orp[a, v]:orp[a, v + 1]]), (a, v)
    if nv > 2:
        assert col.size > 0


def test_prms_end_to_end_matches_sequential_reference_loop():
    """PRMs(config_init, config_goal, connect, 300; rad=0.1) (prm.jl:251-276): vertices 0/1
    are init/goal, the rest the first free samples in try order, then the neighbour pass."""
    N, nv, rad = 5, 300, 0.1
    rads, obs = make_point_instance(O.POINT2D, N, 10, seed=46)
    orc = O.Oracle(O.POINT2D, rads, obs)
    rng = np.random.default_rng(46)
    init, goal = rng.random((N, 2)) * 0.8 + 0.1, rng.random((N, 2)) * 0.8 + 0.1
    connect = S.gen_connect(S.StatePoint2D(0, 0), obs, rads)
    roadmaps = S.PRMs(init, goal, connect, nv, rad=rad, seed=1234)
    nodes = np.empty((N, nv, 2))
    for a in range(N):
        nodes[a, 0], nodes[a, 1] = init[a], goal[a]
        nodes[a, 2:], _ = orc.sample_free(a, 1234, nv - 2)
    orp, ocol, _ = orc.prm_build(0, nodes, rad, nthreads=0)
    for a in range(N):
        assert len(roadmaps[a]) == nv
        for v in range(nv):
            assert tuple(roadmaps[a][v].q) == tuple(nodes[a, v])
            assert roadmaps[a][v].id == v
            assert roadmaps[a][v].neighbors == ocol[a, orp[a, v]:orp[a, v + 1]].tolist()


def test_roadmaps_indexed_collide_and_shards():
    N, nv = 8, 120
    rads, obs = make_point_instance(O.POINT2D, N, 4, seed=3, rad=(0.02, 0.04))
    ctx, orc = S.Context(O.POINT2D, rads, obs), O.Oracle(O.POINT2D, rads, obs)
    rng = np.random.default_rng(8)
    nodes = rng.random((N, nv, 2))
    rm = S.Roadmaps(ctx, 0, N, nv)
    rm.set_nodes(nodes)
    assert np.array_equal(rm.nodes(), nodes)
    n = 200_000
    ai = rng.integers(0, N, n)
    aj = (ai + 1 + rng.integers(0, N - 1, n)) % N
    v = rng.integers(0, nv, (4, n))
    g = rm.collide_edges_indexed(ai, aj, *v)
    o = orc.collide_edges_indexed(ai, aj, *v, nodes, nv, nthreads=0)
    assert np.array_equal(g, o) and 0 < o.sum() < n
    # an agent shard [3, 6) builds the same rows as the full set (multi-GPU partitioning)
    nnz_full = rm.build(0.2)
    rp, col = rm.csr()
    sh = S.Roadmaps(ctx, 3, 3, nv)
    sh.set_nodes(nodes[3:6])
    sh.build(0.2)
    srp, scol = sh.csr()
    lo, hi = rp[3 * nv], rp[6 * nv]
    assert np.array_equal(scol, col[lo:hi]) and np.array_equal(srp, rp[3 * nv:6 * nv + 1] - lo)
    assert nnz_full == col.size


@pytest.mark.parametrize("model,d,nv,rads_prm", [(O.POINT2D, 2, 3000, (0.05, 0.031)),
                                                 (O.POINT3D, 3, 2500, (0.12, 0.2))])
def test_prm_grid_hash_path_bit_exact(model, d, nv, rads_prm):
    """nv >= 2048: neighbour search through the uniform-grid spatial hash (counting sort by
    cell, 3^d cell scan); the CSR must still equal the reference's all-pairs row-major loop,
    including vertices outside the unit cube and different radii per agent."""
    N = 2
    rads, obs = make_point_instance(model, N, 10, seed=77)
    ctx, orc = S.Context(model, rads, obs), O.Oracle(model, rads, obs)
    rng = np.random.default_rng(nv)
    nodes = rng.random((N, nv, d))
    nodes[:, :40] = rng.uniform(-0.2, 1.2, (N, 40, d))     # clamped into the border cells
    nodes[0, 40:60] = nodes[0, 60:80]                        # duplicates (distance 0)
    rm = S.Roadmaps(ctx, 0, N, nv)
    rm.set_nodes(nodes)
    nnz = rm.build(np.array(rads_prm))
    assert rm.used_grid()
    row_ptr, col = rm.csr()
    orp, ocol, onnz = orc.prm_build(0, nodes, np.array(rads_prm), nthreads=0)
    assert nnz == int(onnz.sum()) and nnz > nv
    for a in range(N):
        for v in range(nv):
            r = a * nv + v
            assert np.array_equal(col[row_ptr[r]:row_ptr[r + 1]], ocol[a, orp[a, v]:orp[a, v + 1]]), (a, v)
    # a second build on the same set (buffers reused) and the small-nv path on the same ctx
    assert rm.build(np.array(rads_prm)) == nnz
    rm2 = S.Roadmaps(ctx, 0, N, 500)
    rm2.set_nodes(nodes[:, :500])
    rm2.build(rads_prm[0])
    assert not rm2.used_grid()


def test_csr_shards_pack_and_stitch_equal_a_full_build():
    """The multi-GPU CSR exchange (pack fixed-size slots -> all-gather -> stitch), with the ranks
    emulated on one GPU: 7 agents in blocks of cap = 3 (the last block is short)."""
    import ctypes as C
    import torch
    from sssp_b200 import lib as L
    N, nv, rad, world, cap = 7, 200, 0.12, 3, 3
    rads, obs = make_point_instance(O.POINT2D, N, 10, seed=9)
    ctx = S.Context(O.POINT2D, rads, obs)
    nodes = np.random.default_rng(2).random((N, nv, 2))
    full = S.Roadmaps(ctx, 0, N, nv)
    full.set_nodes(nodes)
    nnz = full.build(rad)
    rp_full, col_full = full.csr()
    lib = L.load()
    cap_nnz = nnz          # generous
    slots = torch.full((world, cap * nv + cap_nnz), -7, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for r in range(world):
        a0, a1 = r * cap, min(N, (r + 1) * cap)
        sh = S.Roadmaps(ctx, a0, a1 - a0, nv)
        sh.set_nodes(nodes[a0:a1])
        sh.build(rad)
        L.check(lib.mrmp_roadmaps_pack_csr_shard_dev(sh._h, cap, cap_nnz, slots[r].data_ptr(), st))
        torch.cuda.synchronize()
        # a slot that is too small is refused
        assert lib.mrmp_roadmaps_pack_csr_shard_dev(sh._h, cap, 3, slots[r].data_ptr(), st) == L.ERR_CAPACITY
    adopt = S.Roadmaps(ctx, 0, N, nv)
    adopt.set_nodes(nodes)
    out = C.c_int64(0)
    L.check(lib.mrmp_roadmaps_set_csr_shards_dev(adopt._h, world, cap, cap_nnz, slots.data_ptr(), st,
                                                 C.byref(out)))
    assert out.value == nnz
    rp, col = adopt.csr()
    assert np.array_equal(rp, rp_full) and np.array_equal(col, col_full)


def _assert_csr_equal(N, nv, row_ptr, col, orp, ocol):
    for a in range(N):
        for v in range(nv):
            r = a * nv + v
            assert np.array_equal(col[row_ptr[r]:row_ptr[r + 1]], ocol[a, orp[a, v]:orp[a, v + 1]]), (a, v)


def _edge_case_nodes(N, nv, d, seed):
    rng = np.random.default_rng(seed)
    nodes = rng.random((N, nv, d))
    nodes[:, :30] = rng.uniform(-0.3, 1.3, (N, 30, d))       # outside the field
    nodes[:, 30:45] = nodes[:, 45:60]                          # duplicates: distance 0, a == b edges
    nodes[:, 60:64] = nodes[:, 64:68] + 1e-170                 # below the squared-norm underflow
    nodes[:, 70] = 3.0e6                                       # beyond the cull's coordinate range
    nodes[:, 71] = 3.0e6 + 0.05
    nodes[0, 72, 0] = np.nan
    return nodes


@pytest.mark.parametrize("model,d", [(O.POINT2D, 2), (O.POINT3D, 3)])
def test_prm_all_pairs_edge_cases(model, d):
    """The all-pairs neighbour pass (branch-free gate, obstacle cull + exact stage) on inputs
    that leave the cull's comfort zone: vertices outside the field, duplicates, coordinates
    beyond 2^20, a NaN vertex, and per-agent radii down to 1e-160 and up to 'connect all'."""
    N, nv = 4, 300
    rads, obs = make_point_instance(model, N, 13, seed=31, rad_obs=(0.03, 0.12))
    obs[12, :d] = 3.0e6 + 0.02                                 # an obstacle next to the far pair
    ctx, orc = S.Context(model, rads, obs), O.Oracle(model, rads, obs)
    nodes = _edge_case_nodes(N, nv, d, 5)
    rads_prm = np.array([0.1, 1e-160, 0.45, 0.2])
    rm = S.Roadmaps(ctx, 0, N, nv)
    rm.set_nodes(nodes)
    nnz = rm.build(rads_prm)
    assert not rm.used_grid()
    row_ptr, col = rm.csr()
    orp, ocol, onnz = orc.prm_build(0, nodes, rads_prm, nthreads=0)
    assert nnz == int(onnz.sum())
    _assert_csr_equal(N, nv, row_ptr, col, orp, ocol)
    assert onnz[1] > 0 and onnz[2] > 20 * nv
    # rad = nothing: every pair is a candidate
    nnz = rm.build(None)
    row_ptr, col = rm.csr()
    orp, ocol, onnz = orc.prm_build(0, nodes, None, nthreads=0)
    assert nnz == int(onnz.sum())
    _assert_csr_equal(N, nv, row_ptr, col, orp, ocol)


def test_prm_all_pairs_table_too_large_for_shared_memory():
    """nv = 6200 with rad = 0.26 (fewer than 4 grid cells per axis -> all pairs): the vertex
    table (99 KB) is not staged, the kernel instantiation that reads it from global runs."""
    N, nv, d = 1, 6200, 2
    rads, obs = make_point_instance(O.POINT2D, N, 10, seed=3)
    ctx, orc = S.Context(O.POINT2D, rads, obs), O.Oracle(O.POINT2D, rads, obs)
    nodes = np.random.default_rng(8).random((N, nv, d))
    rm = S.Roadmaps(ctx, 0, N, nv)
    rm.set_nodes(nodes)
    nnz = rm.build(0.26)
    assert not rm.used_grid()
    row_ptr, col = rm.csr()
    orp, ocol, onnz = orc.prm_build(0, nodes, 0.26, nthreads=0)
    assert nnz == int(onnz.sum()) and nnz > 100 * nv
    assert np.array_equal(row_ptr, orp[0]) and np.array_equal(col, ocol[0, :nnz])


def test_prm_all_pairs_with_more_obstacles_than_the_padded_copy_holds():
    """1500 small obstacles: the re-packed obstacle copy of the neighbour pass (32 KB) is not
    used, the cull reads the ctx arrays at clamped indices; many (edge, obstacle) pairs reach
    the exact stage."""
    N, nv, M = 2, 220, 1500
    rng = np.random.default_rng(17)
    obs = np.concatenate([rng.random((M, 2)), rng.uniform(0.0005, 0.003, (M, 1))], axis=1)
    rads = np.array([0.002, 0.004])
    ctx, orc = S.Context(O.POINT2D, rads, obs), O.Oracle(O.POINT2D, rads, obs)
    nodes = rng.random((N, nv, 2)) * 0.96 + 0.02
    rm = S.Roadmaps(ctx, 0, N, nv)
    rm.set_nodes(nodes)
    nnz = rm.build(np.array([0.15, 0.2]))
    row_ptr, col = rm.csr()
    orp, ocol, onnz = orc.prm_build(0, nodes, np.array([0.15, 0.2]), nthreads=0)
    assert nnz == int(onnz.sum()) and 0 < nnz
    _assert_csr_equal(N, nv, row_ptr, col, orp, ocol)
    # a fair share of the gated pairs is blocked by an obstacle, the rest is connected
    gated = sum(int((np.linalg.norm(nodes[a][:, None] - nodes[a][None], axis=2) <= r).sum()) - nv
                for a, r in ((0, 0.15), (1, 0.2)))
    assert 0.05 * gated < nnz < 0.95 * gated


def test_roadmaps_async_read_back_equals_the_blocking_getters():
    import ctypes as C
    from sssp_b200 import lib as L
    N, nv = 5, 150
    rads, obs = make_point_instance(O.POINT2D, N, 8, seed=4)
    ctx = S.Context(O.POINT2D, rads, obs)
    rng = np.random.default_rng(2)
    rm = S.Roadmaps(ctx, 0, N, nv)
    rm.sample(nv, 9, rng.random((N, 2)) * 0.8 + 0.1, rng.random((N, 2)) * 0.8 + 0.1, tries=False)
    nnz = rm.build(0.2)
    nodes, (row_ptr, col) = rm.nodes(), rm.csr()
    n2, rp2, c2 = np.empty_like(nodes), np.empty_like(row_ptr), np.empty(nnz + 7, np.int32)
    got = C.c_int64(-1)
    lib = L.load()
    L.check(lib.mrmp_roadmaps_read_async(rm._h, n2.ctypes.data, rp2.ctypes.data, c2.ctypes.data,
                                         c2.size, C.byref(got)))
    bits = rm.edge_conflicts(np.zeros(4, np.int32), np.arange(4, dtype=np.int32),
                             np.arange(1, 5, dtype=np.int32))        # a batch in between
    L.check(lib.mrmp_roadmaps_read_wait(rm._h))
    assert got.value == nnz and bits.shape[0] == nnz
    assert np.array_equal(n2, nodes) and np.array_equal(rp2, row_ptr)
    assert np.array_equal(c2[:nnz], col)
    with pytest.raises(S.MRMPError):
        lib_rc = lib.mrmp_roadmaps_read_async(rm._h, n2.ctypes.data, rp2.ctypes.data,
                                              c2.ctypes.data, nnz - 1, C.byref(got))
        L.check(lib_rc)


def test_edge_conflicts_written_in_place_into_pinned_host_memory():
    """OR-reduced conflict table of a point model: with a pinned output buffer the kernel
    stores the bits straight into host memory; the result must equal the staged copy."""
    import ctypes as C
    import torch
    from sssp_b200 import lib as L
    N, nv, T = 6, 120, 40
    rads, obs = make_point_instance(O.POINT2D, N, 8, seed=14)
    ctx = S.Context(O.POINT2D, rads, obs)
    rng = np.random.default_rng(3)
    rm = S.Roadmaps(ctx, 0, N, nv)
    rm.sample(nv, 3, rng.random((N, 2)) * 0.8 + 0.1, rng.random((N, 2)) * 0.8 + 0.1, tries=False)
    nnz = rm.build(0.25)
    row_ptr, col = rm.csr()
    # columns: T time steps x N agents, random roadmap edges
    ca = np.tile(np.arange(N, dtype=np.int32), T)
    vf = rng.integers(0, nv, ca.size).astype(np.int32)
    vt = np.array([col[rng.integers(row_ptr[a * nv + v], max(row_ptr[a * nv + v + 1], row_ptr[a * nv + v] + 1))]
                   if row_ptr[a * nv + v + 1] > row_ptr[a * nv + v] else v
                   for a, v in zip(ca, vf)], np.int32)
    want = rm.edge_conflicts(ca, vf, vt, group_size=N)                  # pageable: staged copy
    wpr = want.shape[1]
    for odd in (0, 1):                                                  # 8- and 4-byte aligned rows
        pinned = torch.full((nnz * wpr + 1,), -1, dtype=torch.int32).pin_memory()
        L.check(L.load().mrmp_roadmaps_edge_conflicts(
            rm._h, None, ca.size, ca.ctypes.data, vf.ctypes.data, vt.ctypes.data, N,
            pinned.data_ptr() + 4 * odd, nnz * wpr))
        got = pinned.numpy()[odd:odd + nnz * wpr].view(np.uint32).reshape(nnz, wpr)
        assert np.array_equal(got, want)
        assert pinned[nnz * wpr if odd == 0 else 0] == -1
    assert want.any()


def test_registered_numpy_buffer_takes_the_in_place_path():
    """mrmp_host_register pins an ordinary host array (what a Julia caller has): the OR-reduced
    table is then written in place and must equal the staged result."""
    from sssp_b200 import lib as L
    N, nv, T = 5, 100, 20
    rads, obs = make_point_instance(O.POINT2D, N, 6, seed=21)
    ctx = S.Context(O.POINT2D, rads, obs)
    rng = np.random.default_rng(4)
    rm = S.Roadmaps(ctx, 0, N, nv)
    rm.sample(nv, 5, rng.random((N, 2)) * 0.8 + 0.1, rng.random((N, 2)) * 0.8 + 0.1, tries=False)
    nnz = rm.build(0.3)
    ca = np.tile(np.arange(N, dtype=np.int32), T)
    vf = rng.integers(0, nv, ca.size).astype(np.int32)
    vt = rng.integers(0, nv, ca.size).astype(np.int32)
    want = rm.edge_conflicts(ca, vf, vt, group_size=N)
    out = np.full(want.shape, 0xdeadbeef, np.uint32)
    lib = L.load()
    L.check(lib.mrmp_host_register(out.ctypes.data, out.nbytes))
    try:
        L.check(lib.mrmp_roadmaps_edge_conflicts(rm._h, None, ca.size, ca.ctypes.data,
                                                 vf.ctypes.data, vt.ctypes.data, N,
                                                 out.ctypes.data, out.size))
    finally:
        L.check(lib.mrmp_host_unregister(out.ctypes.data))
    assert np.array_equal(out, want) and want.any()
