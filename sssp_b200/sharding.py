"""Multi-GPU host logic: one process per GPU, torch.distributed for the plumbing.

The path shards without a data-path collective: agents are independent during
roadmap construction (prm.jl:265-275 is a `map` over agents) and every check of an explicit
list is independent.  The only exchange is the all-gather of the per-agent roadmaps (node
tables and CSR adjacency) that cross-agent collide batches need afterwards.

Everything here works on CPU tensors with the gloo backend (tests) and on CUDA tensors with
NCCL over NVLink (bench / production).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def agent_block(n_agents: int, world: int, rank: int) -> Tuple[int, int, int]:
    """Block partition of agents: returns (agent0, n_local, cap) with cap = ceil(N / world).
    Rank r owns [r*cap, min(N, (r+1)*cap)); with a fixed cap the gathered layout index of
    agent a is simply a."""
    cap = (n_agents + world - 1) // world
    a0 = min(n_agents, rank * cap)
    a1 = min(n_agents, (rank + 1) * cap)
    return a0, a1 - a0, cap


def slice_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of an n-element check list for this rank."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allgather_nodes_inplace(gathered: torch.Tensor, cap: int, nv: int, d: int, group=None) -> None:
    """`gathered` is [world*cap, nv, d]; this rank's shard already sits in its slot (the
    roadmap kernels wrote it there via mrmp_roadmaps_bind_nodes).  In-place all-gather."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return
    flat = gathered.view(world, cap * nv * d)
    dist.all_gather_into_tensor(gathered.view(-1), flat[rank], group=group)


def allgather_csr(row_ptr_local: torch.Tensor, col_local: torch.Tensor, n_agents: int, nv: int,
                  world: int, rank: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather of per-shard CSR adjacency (two-phase: sizes, then padded payload) for a host
    that holds its shard as plain tensors (any backend; tests/test_sharding_gloo.py runs it on
    gloo).  The GPU paths do not come through here: the peer exchange (sssp_b200/peer.py) stores
    fixed-size slots into the peers' memory, and its NCCL fallback all-gathers the same slots
    (mrmp_roadmaps_pack_csr_shard_dev -> one all_gather_into_tensor ->
    mrmp_roadmaps_set_csr_shards_dev), which needs one collective instead of three.

    row_ptr_local: int32 [n_local*nv + 1] with offsets into col_local (this rank's agents);
    returns the global packed CSR (row_ptr int64 [N*nv + 1], col int32 [nnz_total]) in agent
    order, identical on every rank."""
    a0, n_local, cap = agent_block(n_agents, world, rank)
    dev = row_ptr_local.device
    rows_cap = cap * nv
    # per-row counts, padded to the fixed shard size
    cnt = torch.zeros(rows_cap, dtype=torch.int32, device=dev)
    if n_local > 0:
        cnt[: n_local * nv] = row_ptr_local[1:] - row_ptr_local[:-1]
    nnz_local = torch.tensor([int(col_local.numel())], dtype=torch.int64, device=dev)
    if world == 1:
        all_cnt, all_nnz = cnt.view(1, -1), nnz_local.view(1)
    else:
        all_cnt = torch.empty(world * rows_cap, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(all_cnt, cnt, group=group)
        all_cnt = all_cnt.view(world, rows_cap)
        all_nnz = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(all_nnz, nnz_local, group=group)
    max_nnz = int(all_nnz.max().item())
    pad = torch.zeros(max(max_nnz, 1), dtype=torch.int32, device=dev)
    pad[: col_local.numel()] = col_local
    if world == 1:
        all_col = pad.view(1, -1)
    else:
        all_col = torch.empty(world * pad.numel(), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(all_col, pad, group=group)
        all_col = all_col.view(world, -1)
    counts = all_cnt.reshape(-1)[: n_agents * nv].to(torch.int64)
    row_ptr = torch.zeros(n_agents * nv + 1, dtype=torch.int64, device=dev)
    row_ptr[1:] = torch.cumsum(counts, 0)
    nnz_list = all_nnz.tolist()
    col = torch.cat([all_col[r, : nnz_list[r]] for r in range(world)]) if world > 1 \
        else all_col[0, : nnz_list[0]]
    return row_ptr, col
