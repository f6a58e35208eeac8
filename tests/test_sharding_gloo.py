"""Host-side multi-GPU logic at world_size 2 on CPU (gloo): agent block partition, in-place
all-gather of node tables, two-phase all-gather of per-shard CSR adjacency.  The per-shard
roadmaps come from the CPU oracle (test infrastructure); the gathered result must equal
the un-sharded build byte for byte."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sssp_b200 import sharding as SH


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, nv, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        for p in (root, os.path.join(root, "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        from oracle import oracle as O
        from helpers import make_point_instance
        rads, obs = make_point_instance(O.POINT2D, N, 6, seed=1)
        orc = O.Oracle(O.POINT2D, rads, obs)
        nodes = np.random.default_rng(5).random((N, nv, 2))
        a0, n_local, cap = SH.agent_block(N, world, rank)
        # node table all-gather, in place in the gathered buffer
        gathered = torch.zeros((world * cap, nv, 2), dtype=torch.float64)
        gathered[a0:a0 + n_local] = torch.from_numpy(nodes[a0:a0 + n_local])
        SH.allgather_nodes_inplace(gathered, cap, nv, 2)
        ok_nodes = bool(np.array_equal(gathered[:N].numpy(), nodes))
        # shard CSR -> packed local CSR
        rp, col, nnz = orc.prm_build(a0, nodes[a0:a0 + n_local], 0.3)
        row_ptr = np.zeros(n_local * nv + 1, np.int32)
        cols = []
        for a in range(n_local):
            row_ptr[a * nv + 1:(a + 1) * nv + 1] = row_ptr[a * nv] + rp[a, 1:]
            cols.append(col[a, :nnz[a]])
        cols = np.concatenate(cols) if cols else np.zeros(0, np.int32)
        g_rp, g_col = SH.allgather_csr(torch.from_numpy(row_ptr), torch.from_numpy(cols), N, nv,
                                       world, rank)
        # un-sharded build
        frp, fcol, fnnz = orc.prm_build(0, nodes, 0.3)
        exp_rp = np.zeros(N * nv + 1, np.int64)
        exp_cols = []
        for a in range(N):
            exp_rp[a * nv + 1:(a + 1) * nv + 1] = exp_rp[a * nv] + frp[a, 1:]
            exp_cols.append(fcol[a, :fnnz[a]])
        exp_cols = np.concatenate(exp_cols)
        ok_csr = bool(np.array_equal(g_rp.numpy(), exp_rp) and np.array_equal(g_col.numpy(), exp_cols))
        lo, hi = SH.slice_range(1001, world, rank)
        q.put((rank, ok_nodes, ok_csr, (lo, hi), int(exp_cols.size)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N", [5, 8])
def test_world2_allgather_matches_unsharded(N):
    world, nv = 2, 40
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, nv, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(r[1] for r in res), "node all-gather mismatch"
    assert all(r[2] for r in res), "CSR all-gather mismatch"
    assert res[0][3][0] == 0 and res[0][3][1] == res[1][3][0] and res[1][3][1] == 1001
    assert res[0][4] > 0


def test_partition_helpers():
    for N in (1, 5, 50, 64):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                a0, n, cap = SH.agent_block(N, world, r)
                assert a0 == min(N, r * cap) and n >= 0
                seen += list(range(a0, a0 + n))
            assert seen == list(range(N))
            lo_prev = 0
            for r in range(world):
                lo, hi = SH.slice_range(N * 7 + 3, world, r)
                assert lo == lo_prev and hi >= lo
                lo_prev = hi
            assert lo_prev == N * 7 + 3


def _worker_table(rank, world, port, N, nv, q):
    """Strong-scaled conflict table, host logic only (oracle as the compute): every rank builds the
    CSR of its agent block, computes the table rows of ITS agents against all agents' path edges
    and the rows are gathered on rank 0 in rank order == packed CSR order of the full set."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        for p in (root, os.path.join(root, "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        from oracle import oracle as O
        from helpers import make_point_instance
        from sssp_b200 import peer as P
        rads, obs = make_point_instance(O.POINT2D, N, 4, seed=2, rad=(0.03, 0.05))
        orc = O.Oracle(O.POINT2D, rads, obs)
        rng = np.random.default_rng(7)            # same stream on every rank: vertices are re-drawn,
        nodes = rng.random((N, nv, 2))            # not communicated (counter-based sampler in the product)
        T = 5
        ca = np.tile(np.arange(N, dtype=np.int32), T)
        vf = rng.integers(0, nv, N * T)
        vt = rng.integers(0, nv, N * T)
        cf, ct = nodes[ca, vf], nodes[ca, vt]

        def rows_of(a0, n_local):
            rp, col, nnz = orc.prm_build(a0, nodes[a0:a0 + n_local], 0.3)
            ra, rf, rt = [], [], []
            for a in range(n_local):
                for v in range(nv):
                    for k in col[a, rp[a, v]:rp[a, v + 1]]:
                        ra.append(a0 + a); rf.append(nodes[a0 + a, v]); rt.append(nodes[a0 + a, k])
            return np.array(ra, np.int32), np.array(rf).reshape(-1, 2), np.array(rt).reshape(-1, 2)

        a0, n_local, cap = SH.agent_block(N, world, rank)
        ra, rf, rt = rows_of(a0, n_local)
        mine = orc.collide_cross(ra, rf, rt, ca, cf, ct, N)            # rows of this rank's agents
        wpr = mine.shape[1]
        # fixed-size slots like the peer exchange (sssp_b200/peer.py): rows, then padding
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([mine.shape[0]], dtype=torch.int64))
        counts = [int(c.item()) for c in counts]
        cap_rows = max(counts) + 3
        slot = torch.zeros((cap_rows, wpr), dtype=torch.int64)
        slot[: mine.shape[0]] = torch.from_numpy(mine.astype(np.int64))
        gathered = [torch.zeros_like(slot) for _ in range(world)] if rank == 0 else None
        dist.gather(slot, gathered, dst=0)
        ok = True
        if rank == 0:
            table = np.concatenate([gathered[r][: counts[r]].numpy().astype(np.uint32) for r in range(world)])
            fa, ff, ft = rows_of(0, N)
            full = orc.collide_cross(fa, ff, ft, ca, cf, ct, N)
            ok = bool(table.shape == full.shape and np.array_equal(table, full) and full.any())
        # slot sizing helper: room for the largest shard, 16-byte multiple together with the counts
        cw = P.csr_slot_words(cap, nv, max(counts))
        ok = ok and cw >= max(counts) and (cap * nv + cw) % 4 == 0
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_world2_row_sharded_table_matches_full():
    world, N, nv = 2, 5, 24
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_table, args=(r, world, port, N, nv, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
