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
