"""One conflict table strong-scaled over several ranks (emulated on one GPU): every rank draws
all vertices, builds the roadmaps of its agent block (prm.jl:265-275), pushes its CSR shard into
every rank's exchange buffer, evaluates the table rows of ITS agents against all agents' path
edges (pp.jl:179-189) and writes them into rank 0's buffer.  The stitched roadmaps and the
gathered table must equal the single-GPU results byte for byte."""
import numpy as np
import pytest

import sssp_b200 as S
from sssp_b200 import peer as P, sharding as SH, workload as W
from oracle import oracle as O

pytestmark = pytest.mark.gpu
L_ERR_CAPACITY = -3


@pytest.mark.parametrize("world,via", [(1, "direct"), (3, "direct"), (4, "direct"), (3, "push"), (5, "push"),
                                       (1, "async"), (4, "async")])
def test_row_sharded_table_equals_single_gpu_table(world, via):
    import torch
    N, nv, T = 10, 160, 17
    inst = W.point2d_many(N=N, seed=3, nv=nv, rad=0.17)
    st = torch.cuda.current_stream().cuda_stream
    # ---- single GPU
    ctx0 = S.Context(inst.model, inst.rads, inst.obstacles)
    ctx0.set_stream(st)
    full = S.Roadmaps(ctx0, 0, N, nv)
    full.sample(nv, inst.seed, inst.q_init, inst.q_goal)
    nnz = full.build(inst.rad)
    rp_ref, col_ref = full.csr()
    _, ca, vf, vt = W.random_walk_paths(inst, rp_ref, col_ref, T, seed=5)
    bits_ref = full.edge_conflicts(ca, vf, vt, N)
    wpr = bits_ref.shape[1]
    assert bits_ref.any() and nnz == bits_ref.shape[0]
    # the oracle agrees with the single-GPU table (so the sharded one is pinned to it as well)
    nodes = full.nodes()
    rows = np.repeat(np.arange(N * nv), np.diff(rp_ref))
    ra = (rows // nv).astype(np.int32)
    sel = slice(0, None, 9)
    obits = O.Oracle(O.POINT2D, inst.rads, inst.obstacles).collide_cross(
        ra[sel], nodes[ra, rows % nv][sel], nodes[ra, col_ref][sel], ca, nodes[ca, vf], nodes[ca, vt], N)
    assert np.array_equal(bits_ref[sel], obits)

    # ---- `world` ranks on this GPU, each with its own context
    cols_d = [torch.from_numpy(x).cuda() for x in (ca, vf, vt)]
    ranks = []
    nnz_sh = []
    for r in range(world):
        a0, n_local, cap = SH.agent_block(N, world, r)
        ctx = S.Context(inst.model, inst.rads, inst.obstacles)
        ctx.set_stream(st)
        allrm = S.Roadmaps(ctx, 0, N, nv)
        allrm.sample(nv, inst.seed, inst.q_init, inst.q_goal)
        sh = None
        if n_local > 0:
            sh = S.Roadmaps(ctx, a0, n_local, nv)
            sh.bind_nodes(nv, allrm.device_ptrs()["nodes"] + a0 * nv * 2 * 8)
            nnz_sh.append(sh.build(inst.rad))
        else:
            nnz_sh.append(0)
        ranks.append((ctx, allrm, sh, a0, n_local, cap))
    cap = ranks[0][5]
    cap_nnz = P.csr_slot_words(cap, nv, max(nnz_sh))
    tab_words = (max(nnz_sh) + 64) * wpr
    tab_words += (-tab_words) % 4
    xs = [P.Exchange(ranks[r][0], world, r, [4 * (cap * nv + cap_nnz), 4 * tab_words])
          for r in range(world)]
    P.Exchange.connect_local(xs)
    for round_ in range(3):            # several rounds: epochs and double buffering
        for r, (ctx, allrm, sh, a0, n_local, _) in enumerate(ranks):
            if via == "async":     # no host wait between the kernels: the edge count stays on the device
                allrm.sample(nv, inst.seed, inst.q_init, inst.q_goal, tries=False)
                sh.build_async(inst.rad)
            xs[r].push_csr_shard(sh, 0, cap, cap_nnz, st)
            if via == "async":
                local = torch.full((tab_words,), -1, dtype=torch.int32, device="cuda")
                sh.edge_conflicts_async(ca.size, cols_d[0].data_ptr(), cols_d[1].data_ptr(),
                                        cols_d[2].data_ptr(), N, local.data_ptr(), tab_words // wpr, st,
                                        cols=allrm)
                xs[r].push_table_rows(sh, 1, 0, local.data_ptr(), wpr, tab_words // wpr, st)
            elif via == "direct":  # the table kernel stores straight into rank 0's buffer
                sh.edge_conflicts_dev(ca.size, cols_d[0].data_ptr(), cols_d[1].data_ptr(),
                                      cols_d[2].data_ptr(), N, xs[r].ptr(1, 0), st, cols=allrm)
            else:                  # rows into local memory (finer work items), then a copy kernel
                local = torch.full((tab_words,), -1, dtype=torch.int32, device="cuda")
                sh.edge_conflicts_dev(ca.size, cols_d[0].data_ptr(), cols_d[1].data_ptr(),
                                      cols_d[2].data_ptr(), N, local.data_ptr(), st, cols=allrm)
                nb = nnz_sh[r] * wpr * 4
                xs[r].push(1, 0, local.data_ptr(), nb + (-nb) % 16, st)
            xs[r].signal(1, st, dst_mask=1)
        for r, (ctx, allrm, sh, a0, n_local, _) in enumerate(ranks):
            if via == "async":
                xs[r].adopt_csr_shards_async(allrm, 0, cap, cap_nnz, st)
                assert sh.finish() == nnz_sh[r] and allrm.finish() == nnz
            else:
                assert xs[r].adopt_csr_shards(allrm, 0, cap, cap_nnz, st) == nnz
            rp, col = allrm.csr()
            assert np.array_equal(rp, rp_ref) and np.array_equal(col, col_ref)
        xs[0].wait(1, st)
        torch.cuda.synchronize()
        assert xs[0].error() == 0
        slots = torch.as_tensor(_CudaArray(xs[0].slots(1), (world, xs[0].slot_bytes(1) // 4), "<i4"),
                                device="cuda").cpu().numpy().view(np.uint32)
        got = np.concatenate([slots[r, : nnz_sh[r] * wpr].reshape(-1, wpr) for r in range(world)])
        assert got.shape == bits_ref.shape and np.array_equal(got, bits_ref)
        for x in xs:
            x.advance(0)
            x.advance(1)


@pytest.mark.parametrize("world", [1, 2, 5])
def test_strong_step_in_one_call(world):
    """mrmp_roadmaps_strong_step: the whole step of a rank enqueued by one call (own stream per
    emulated rank, adoption on the ctx's second stream) gives the single-GPU roadmaps and table."""
    import torch
    N, nv, T = 10, 160, 17
    inst = W.point2d_many(N=N, seed=4, nv=nv, rad=0.17)
    ctx0 = S.Context(inst.model, inst.rads, inst.obstacles)
    full = S.Roadmaps(ctx0, 0, N, nv)
    full.sample(nv, inst.seed, inst.q_init, inst.q_goal)
    nnz = full.build(inst.rad)
    rp_ref, col_ref = full.csr()
    _, ca, vf, vt = W.random_walk_paths(inst, rp_ref, col_ref, T, seed=6)
    bits_ref = full.edge_conflicts(ca, vf, vt, N)
    wpr = bits_ref.shape[1]
    cols_d = [torch.from_numpy(x).cuda() for x in (ca, vf, vt)]
    torch.cuda.synchronize()
    ranks, nnz_sh, streams = [], [], []
    for r in range(world):
        a0, n_local, cap = SH.agent_block(N, world, r)
        assert n_local > 0
        ctx = S.Context(inst.model, inst.rads, inst.obstacles)
        streams.append(torch.cuda.Stream())
        ctx.set_stream(streams[-1].cuda_stream)
        allrm = S.Roadmaps(ctx, 0, N, nv)
        allrm.sample(nv, inst.seed, inst.q_init, inst.q_goal)
        sh = S.Roadmaps(ctx, a0, n_local, nv)
        sh.bind_nodes(nv, allrm.device_ptrs()["nodes"] + a0 * nv * 2 * 8)
        nnz_sh.append(sh.build(inst.rad))
        ranks.append((ctx, allrm, sh, a0, n_local, cap))
    cap = ranks[0][5]
    cap_nnz = P.csr_slot_words(cap, nv, max(nnz_sh))
    tab_words = (max(nnz_sh) + 64) * wpr
    tab_words += (-tab_words) % 4
    xs = [P.Exchange(ranks[r][0], world, r, [4 * (cap * nv + cap_nnz), 4 * tab_words])
          for r in range(world)]
    P.Exchange.connect_local(xs)
    local = [torch.full((tab_words,), -1, dtype=torch.int32, device="cuda") for _ in range(world)]
    torch.cuda.synchronize()
    steps = [P.StrongStep(xs[r], ranks[r][1], ranks[r][2], nv, inst.seed, inst.q_init, inst.q_goal,
                          inst.rad, cap, cap_nnz, ca.size, cols_d[0].data_ptr(), cols_d[1].data_ptr(),
                          cols_d[2].data_ptr(), N, local[r].data_ptr(), wpr, tab_words // wpr, table_dst=0)
             for r in range(world)]
    for round_ in range(3):
        for r in range(world):
            steps[r].enqueue()
        for r in range(world):
            assert steps[r].finish() == nnz
            rp, col = ranks[r][1].csr()
            assert np.array_equal(rp, rp_ref) and np.array_equal(col, col_ref)
        torch.cuda.synchronize()
        slots = torch.as_tensor(_CudaArray(xs[0].slots(1), (world, xs[0].slot_bytes(1) // 4), "<i4"),
                                device="cuda").cpu().numpy().view(np.uint32)
        got = np.concatenate([slots[r, : nnz_sh[r] * wpr].reshape(-1, wpr) for r in range(world)])
        assert got.shape == bits_ref.shape and np.array_equal(got, bits_ref)


class _CudaArray:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 3}


def test_exchange_wait_gives_up_instead_of_hanging():
    """A rank that never publishes must not hang the stream: the wait times out and reports."""
    import torch
    rads = np.array([0.02, 0.03])
    ctx = S.Context(O.POINT2D, rads, np.zeros((0, 3)))
    st = torch.cuda.current_stream().cuda_stream
    xs = [P.Exchange(ctx, 2, r, [64]) for r in range(2)]
    P.Exchange.connect_local(xs)
    xs[0].signal(0, st)                  # rank 0 publishes to both, rank 1 never does
    xs[0].wait(0, st)
    torch.cuda.synchronize()
    assert xs[0].error() == 2            # 1 + the missing rank


def test_async_build_reports_capacity_overflow():
    """A table computed behind an asynchronous build into too little room is truncated on the
    device and rejected by finish(); host-side queries are refused while the build is pending."""
    import torch
    N, nv = 4, 120
    inst = W.point2d_many(N=N, seed=1, nv=nv, rad=0.2)
    ctx = S.Context(inst.model, inst.rads, inst.obstacles)
    st = torch.cuda.current_stream().cuda_stream
    ctx.set_stream(st)
    rm = S.Roadmaps(ctx, 0, N, nv)
    rm.sample(nv, inst.seed, inst.q_init, inst.q_goal)
    nnz = rm.build(inst.rad)
    rp, col = rm.csr()
    _, ca, vf, vt = W.random_walk_paths(inst, rp, col, 5, seed=2)
    ref = rm.edge_conflicts(ca, vf, vt, N)
    cols_d = [torch.from_numpy(x).cuda() for x in (ca, vf, vt)]
    out = torch.full((nnz + 64, ref.shape[1]), -1, dtype=torch.int32, device="cuda")
    rm.build_async(inst.rad)
    with pytest.raises(S.MRMPError):
        rm.csr()                                      # pending: nnz is not on the host yet
    rm.edge_conflicts_async(ca.size, cols_d[0].data_ptr(), cols_d[1].data_ptr(), cols_d[2].data_ptr(),
                            N, out.data_ptr(), nnz + 64, st)
    assert rm.finish() == nnz
    assert np.array_equal(out[:nnz].cpu().numpy().view(np.uint32), ref)
    out.fill_(-1)
    rm.build_async(inst.rad)
    rm.edge_conflicts_async(ca.size, cols_d[0].data_ptr(), cols_d[1].data_ptr(), cols_d[2].data_ptr(),
                            N, out.data_ptr(), nnz // 2, st)
    with pytest.raises(S.MRMPError) as e:
        rm.finish()
    assert e.value.code == L_ERR_CAPACITY
    assert (out[nnz // 2 + 32:].cpu().numpy() == -1).all()    # nothing written beyond the room given


def test_consumers_on_another_stream_are_ordered_behind_the_build():
    """The ctx runs on its own stream, the table is requested on the default stream: the library
    orders the consumer behind the sampling / build kernels (inputs change every round, so a table
    computed from the previous round's roadmaps would not match)."""
    import torch
    N, nv = 16, 300
    inst = W.point2d_many(N=N, seed=2, nv=nv, rad=0.15)
    ctx = S.Context(inst.model, inst.rads, inst.obstacles)        # own (non-blocking) stream
    ctx_ref = S.Context(inst.model, inst.rads, inst.obstacles)
    rm, rm_ref = S.Roadmaps(ctx, 0, N, nv), S.Roadmaps(ctx_ref, 0, N, nv)
    rng = np.random.default_rng(0)
    n_cols = 12 * N
    ca = np.repeat(np.arange(N, dtype=np.int32), 12)
    vf = rng.integers(0, nv, n_cols).astype(np.int32)
    vt = rng.integers(0, nv, n_cols).astype(np.int32)
    cols_d = [torch.from_numpy(x).cuda() for x in (ca, vf, vt)]
    cap = 200000
    out = torch.empty((cap, 1), dtype=torch.int32, device="cuda")
    for seed, rad in [(11, 0.15), (12, 0.2), (13, 0.11), (14, 0.18)]:
        rm_ref.sample(nv, seed, inst.q_init, inst.q_goal)
        nnz_ref = rm_ref.build(rad)
        ref = rm_ref.edge_conflicts(ca, vf, vt, N)
        assert nnz_ref <= cap
        out.fill_(-1)
        rm.sample(nv, seed, inst.q_init, inst.q_goal, tries=False)
        rm.build_async(rad)
        rm.edge_conflicts_async(n_cols, cols_d[0].data_ptr(), cols_d[1].data_ptr(), cols_d[2].data_ptr(),
                                N, out.data_ptr(), cap, 0)
        assert rm.finish() == nnz_ref
        torch.cuda.synchronize()
        assert np.array_equal(out[:nnz_ref].cpu().numpy().view(np.uint32), ref)


def test_read_back_queued_behind_an_asynchronous_build():
    """mrmp_roadmaps_read_async on a pending set: the copies are ordered behind the build only, the
    count arrives with finish(), too little room is reported there."""
    import ctypes as C
    import torch
    from sssp_b200 import lib as L
    N, nv = 6, 150
    inst = W.point2d_many(N=N, seed=5, nv=nv, rad=0.2)
    ctx = S.Context(inst.model, inst.rads, inst.obstacles)
    rm = S.Roadmaps(ctx, 0, N, nv)
    rm.sample(nv, inst.seed, inst.q_init, inst.q_goal)
    nnz = rm.build(inst.rad)
    rp_ref, col_ref = rm.csr()
    nodes_ref = rm.nodes()
    lib = ctx._lib
    nodes_h = torch.zeros((N, nv, 2), dtype=torch.float64).pin_memory()
    rp_h = torch.zeros(N * nv + 1, dtype=torch.int32).pin_memory()
    col_h = torch.full((nnz + 100,), -7, dtype=torch.int32).pin_memory()
    out = C.c_int64(0)
    for cap, ok in ((nnz + 100, True), (nnz, True), (nnz - 1, False)):
        rm.sample(nv, inst.seed, inst.q_init, inst.q_goal, tries=False)
        rm.build_async(inst.rad)
        L.check(lib.mrmp_roadmaps_read_async(rm._h, nodes_h.data_ptr(), rp_h.data_ptr(), col_h.data_ptr(),
                                             cap, C.byref(out)))
        assert out.value == -1
        if ok:
            assert rm.finish() == nnz
            L.check(lib.mrmp_roadmaps_read_wait(rm._h))
            assert np.array_equal(rp_h.numpy(), rp_ref) and np.array_equal(col_h.numpy()[:nnz], col_ref)
            assert np.array_equal(nodes_h.numpy(), nodes_ref)
        else:
            with pytest.raises(S.MRMPError) as e:
                rm.finish()
            assert e.value.code == L_ERR_CAPACITY
            L.check(lib.mrmp_roadmaps_read_wait(rm._h))
