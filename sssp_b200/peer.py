"""Peer-memory exchange between the GPUs of one box (include/mrmp_b200.h, "peer-memory exchange").

One process per GPU; torch.distributed only carries the 64-byte CUDA IPC handles at set-up.
After that the roadmap shards (CSR of each rank's agent block, prm.jl:265-275) and the
conflict-table rows (pp.jl:179-189) are written by the producing kernels straight into the
peers' memory over NVLink, followed by one flag per peer -- no collective call in the data path.

`Exchange.connect()` returns False when the peer mapping cannot be opened on this machine
(cudaIpcOpenMemHandle refused): callers then fall back to the NCCL all-gather path of
sssp_b200.sharding / mrmp_roadmaps_pack_csr_shard_dev.
"""
from __future__ import annotations

import ctypes as C
import sys
from typing import List, Optional, Sequence

import numpy as np

from . import lib as L


class Exchange:
    """n_chan channels of fixed-size per-rank slots, double buffered by epoch."""

    def __init__(self, ctx, world: int, rank: int, slot_bytes: Sequence[int]):
        self.ctx, self.world, self.rank = ctx, int(world), int(rank)
        self.n_chan = len(slot_bytes)
        self._lib = ctx._lib
        sb = (C.c_int64 * self.n_chan)(*[int(b) for b in slot_bytes])
        self.handle = np.zeros(int(self._lib.mrmp_exchange_handle_bytes()), np.uint8)
        h = C.c_void_p()
        L.check(self._lib.mrmp_exchange_create(ctx._h, self.world, self.rank, self.n_chan, sb,
                                               C.byref(h), self.handle.ctypes.data_as(L._vp)))
        self._h = h
        ctx._adopt()      # the context outlives this object (see Context.close)
        self.all_mask = (1 << self.world) - 1
        self.connected = self.world == 1
        if self.world == 1:   # a single rank is its own peer
            bufs = (C.c_void_p * 1)(self._lib.mrmp_exchange_local_buffer(self._h))
            L.check(self._lib.mrmp_exchange_connect_ptrs(self._h, bufs))

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.mrmp_exchange_destroy(h)
            self.ctx._release()

    def __del__(self):
        if sys.is_finalizing():   # the CUDA runtime may already be gone at interpreter shutdown
            return
        try:
            self.close()
        except Exception:
            pass

    # ---- set-up -----------------------------------------------------------------------------
    def connect(self, group=None) -> bool:
        """All-gather the IPC handles over torch.distributed (any backend) and open the peers'
        buffers.  Returns True when EVERY rank succeeded, False otherwise (nobody uses it then)."""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return True
        backend = dist.get_backend(group)
        dev = torch.device("cuda", self.ctx.device) if backend == "nccl" else torch.device("cpu")
        mine = torch.from_numpy(self.handle.copy()).to(dev)
        allh = torch.empty(self.world * mine.numel(), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, mine, group=group)
        handles = np.ascontiguousarray(allh.cpu().numpy())
        rc = self._lib.mrmp_exchange_connect(self._h, handles.ctypes.data_as(L._vp))
        ok = torch.tensor([1 if rc == L.OK else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        self.connected = bool(int(ok.item()))
        self.connect_error = None if rc == L.OK else self._lib.mrmp_last_error().decode("utf-8", "replace")
        return self.connected

    @staticmethod
    def connect_local(xs: List["Exchange"]) -> None:
        """Several ranks driven by one process on one device (tests): raw pointers."""
        bufs = (C.c_void_p * len(xs))(*[x._lib.mrmp_exchange_local_buffer(x._h) for x in xs])
        for x in xs:
            L.check(x._lib.mrmp_exchange_connect_ptrs(x._h, bufs))
            x.connected = True

    # ---- one round --------------------------------------------------------------------------
    def ptr(self, ch: int, dst: int) -> int:
        """Device address (in rank dst's buffer) where THIS rank writes its slot of channel ch."""
        p = self._lib.mrmp_exchange_ptr(self._h, int(ch), int(dst))
        if not p:
            raise L.MRMPError(L.ERR_INVALID, "exchange not connected to rank %d" % dst)
        return int(p)

    def slots(self, ch: int) -> int:
        """Base of the local [world][slot] block of channel ch (valid after wait)."""
        return int(self._lib.mrmp_exchange_slots(self._h, int(ch)))

    def slot_bytes(self, ch: int) -> int:
        return int(self._lib.mrmp_exchange_slot_bytes(self._h, int(ch)))

    def push(self, ch: int, dst: int, src_ptr: int, nbytes: int, stream: int) -> None:
        """Copy kernel: nbytes (multiple of 16) of this rank's device memory -> its slot in rank dst."""
        L.check(self._lib.mrmp_exchange_push(self._h, int(ch), int(dst), src_ptr, int(nbytes), stream))

    def signal(self, ch: int, stream: int, dst_mask: Optional[int] = None) -> None:
        L.check(self._lib.mrmp_exchange_signal(self._h, int(ch),
                                               self.all_mask if dst_mask is None else int(dst_mask),
                                               stream))

    def wait(self, ch: int, stream: int, src_mask: Optional[int] = None) -> None:
        L.check(self._lib.mrmp_exchange_wait(self._h, int(ch),
                                             self.all_mask if src_mask is None else int(src_mask),
                                             stream))

    def advance(self, ch: int) -> None:
        L.check(self._lib.mrmp_exchange_advance(self._h, int(ch)))

    def error(self) -> int:
        return int(self._lib.mrmp_exchange_error(self._h))

    # ---- roadmap shards -------------------------------------------------------------------------
    def push_csr_shard(self, rm_shard, ch: int, cap_agents: int, cap_nnz: int, stream: int) -> None:
        L.check(self._lib.mrmp_roadmaps_push_csr_shard(rm_shard._h, self._h, int(ch), self.world,
                                                       int(cap_agents), int(cap_nnz), stream))

    def adopt_csr_shards(self, rm_all, ch: int, cap_agents: int, cap_nnz: int, stream: int) -> int:
        out = C.c_int64(0)
        L.check(self._lib.mrmp_roadmaps_adopt_csr_shards(rm_all._h, self._h, int(ch), self.world,
                                                         int(cap_agents), int(cap_nnz), stream,
                                                         C.byref(out)))
        return int(out.value)


    def adopt_csr_shards_async(self, rm_all, ch: int, cap_agents: int, cap_nnz: int, stream: int) -> None:
        """No wait: rm_all stays pending until rm_all.finish()."""
        L.check(self._lib.mrmp_roadmaps_adopt_csr_shards_async(rm_all._h, self._h, int(ch), self.world,
                                                               int(cap_agents), int(cap_nnz), stream))

    def push_table_rows(self, rm, ch: int, dst: int, bits_ptr: int, words_per_row: int,
                        max_rows: int, stream: int) -> None:
        """Table rows of `rm` (device memory of this rank) -> this rank's slot in rank dst."""
        L.check(self._lib.mrmp_roadmaps_push_table_rows(rm._h, self._h, int(ch), int(dst), bits_ptr,
                                                        int(words_per_row), int(max_rows), stream))


class StrongStep:
    """One strong-scaled step of the roadmap + conflict-table path per foreign call
    (mrmp_roadmaps_strong_step): the arguments are marshalled once, `enqueue()` is a single call
    into the library, `finish()` the host waits of the step.  Channel 0 of `ex` carries the CSR
    shards, channel 1 the table rows (to rank `table_dst`)."""

    def __init__(self, ex: Exchange, rm_all, rm_shard, nv: int, seed: int, q_init, q_goal, rad_shard,
                 cap_agents: int, cap_nnz: int, n_cols: int, col_agent_ptr: int, col_vfrom_ptr: int,
                 col_vto_ptr: int, group_size: int, tab_local_ptr: int, words_per_row: int,
                 tab_rows_cap: int, table_dst: int = 0):
        d = rm_all.ctx.d
        self.ex, self.rm_all, self.rm_shard = ex, rm_all, rm_shard
        # host arrays the library reads during the call: kept alive here
        self._qi = np.ascontiguousarray(np.asarray(q_init, np.float64).reshape(-1, d))
        self._qg = np.ascontiguousarray(np.asarray(q_goal, np.float64).reshape(-1, d))
        self._rad = np.broadcast_to(np.asarray(rad_shard, np.float64), (rm_shard.n,)).copy()
        self._fn = ex._lib.mrmp_roadmaps_strong_step
        self._head = (rm_all._h, rm_shard._h, ex._h, C.c_int(ex.world))
        self._tail = (C.c_int(int(nv)), C.c_uint64(int(seed)), L.ptr(self._qi, L._dp),
                      L.ptr(self._qg, L._dp), L.ptr(self._rad, L._dp), C.c_int(int(cap_agents)),
                      C.c_int64(int(cap_nnz)), C.c_int64(int(n_cols)), C.c_void_p(int(col_agent_ptr)),
                      C.c_void_p(int(col_vfrom_ptr)), C.c_void_p(int(col_vto_ptr)),
                      C.c_int(int(group_size)), C.c_void_p(int(tab_local_ptr)),
                      C.c_int(int(words_per_row)), C.c_int64(int(tab_rows_cap)),
                      C.c_int(int(table_dst)), C.c_int(1 if ex.rank == int(table_dst) else 0))
        self._adv = (C.c_int(0), C.c_int(1))
        self.rounds = 0
        rm_all.nv = int(nv)

    def enqueue(self, advance=None) -> None:
        """advance: move both channels to the next epoch first (default: every round but the
        first one of this object)."""
        adv = self.rounds > 0 if advance is None else bool(advance)
        L.check(self._fn(*self._head, self._adv[1 if adv else 0], *self._tail))
        self.rounds += 1

    def finish(self) -> int:
        """The host waits of the step; returns the edge count of the full set."""
        self.rm_shard.finish()
        n = self.rm_all.finish()
        if self.ex.error():
            raise L.MRMPError(L.ERR_CUDA, "exchange: rank %d never published" % (self.ex.error() - 1))
        return n


def csr_slot_words(cap_agents: int, nv: int, nnz_max: int) -> int:
    """cap_nnz for a shard slot: head room over the largest shard, and the slot
    (cap_agents*nv + cap_nnz words) a multiple of 16 bytes."""
    cap_nnz = int(nnz_max) * 5 // 4 + 1024
    rows_cap = int(cap_agents) * int(nv)
    cap_nnz += (-(rows_cap + cap_nnz)) % 4
    return cap_nnz
