"""Copy-engine K|V exchange for sequence-parallel attention (csrc/peer_ring.cu).

Every rank of the sequence-parallel group needs every other rank's K|V shard once per layer.  Each rank owns one receive
buffer PER PEER plus a small flag page, allocated by the library and exported with CUDA IPC; at set-up every rank maps
every peer's buffers and flags.  Per layer, entirely stream-ordered and without SMs (the attention kernel owns them all):

    sender   (comm stream)     for the peers in the order they will consume me (rank+1, rank+2, ...):
                                 wait  own.free[peer]  >= id of my previous transfer into that peer  (buffer consumed)
                                 copy  local K|V -> peer.recv[me]      cudaMemcpyAsync peer copy: DMA engines over NVLink
                                 write peer.ready[me] = T               (cuStreamWriteValue32; ordered after the copy)
    receiver (attention kernel) ONE launch over [local, recv[rank-1], recv[rank-2], ...]; its TMA producer warp polls
                                 ready[src] >= T (ld.acquire.sys) before the first load from that shard
             (compute stream)   write src.free[me] = T for every source after the kernel

Transfer ids T increase monotonically and identically on every rank (one per attention call), so nothing is ever reset
and there is no host synchronisation.  Pushing the local shard straight to every peer (an all-gather on the copy
engines) instead of forwarding hop by hop removes the arrival -> forward dependency chain of a ring: NVSwitch gives every
pair full bandwidth, and the k-th shard a rank needs is the k-th one its source sends.  Control plane: one
`all_gather_object` of the IPC handles over whatever process group the caller has (NCCL or gloo).
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch
import torch.distributed as dist

from . import _C
from .ops import check

FLAG_BYTES = 256   # ready[slot] (u32) at 4 * slot, free[slot] at 128 + 4 * slot; slot < 32
FREE_OFF = 128


class _Raw:
    """Zero-copy view of library-owned device memory for torch.as_tensor."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _alloc(nbytes: int):
    lib = _C.load()
    ptr = C.c_void_p()
    handle = C.create_string_buffer(64)
    check(lib.ld_ipc_alloc(nbytes, C.byref(ptr), handle), "ld_ipc_alloc")
    return ptr.value, handle.raw


def _open(handle: bytes) -> int:
    ptr = C.c_void_p()
    check(_C.load().ld_ipc_open(handle, C.byref(ptr)), "ld_ipc_open")
    return ptr.value


class PeerGather:
    """K|V all-gather over the copy engines for a sequence-parallel group of n ranks (n <= 4 shards per attention launch,
    csrc/attn_tcgen05.cu).  Slot numbering: on rank r the shard of source s lives in recv[((r - s) % n) - 1], i.e. slot
    j - 1 holds the shard of rank r - j — the j-th shard the attention kernel consumes after the local one."""

    def __init__(self, group, group_ranks: List[int], my_rank: int, shard_shape, dtype, device, max_ranks: int = 4):
        self.lib = _C.load()
        self.device = device
        n = len(group_ranks)
        if not 2 <= n <= max_ranks or max_ranks > 32:   # the flag page holds 32 ready + 32 free words
            raise ValueError(f"groups of 2..{max_ranks} ranks are supported, got {n}")
        self.n = n
        self.me = group_ranks.index(my_rank)
        self.nbytes = int(torch.empty(shard_shape, dtype=dtype, device="meta").numel()) * torch.empty((), dtype=dtype).element_size()
        self.shape, self.dtype = tuple(shard_shape), dtype
        # own resources: one receive buffer per peer, one flag page
        self._recv_ptr, recv_handles = [], []
        for _ in range(n - 1):
            p, h = _alloc(self.nbytes)
            self._recv_ptr.append(p)
            recv_handles.append(h)
        self._flags_ptr, flags_handle = _alloc(FLAG_BYTES)
        self.recv = [torch.as_tensor(_Raw(p, self.nbytes), device=device).view(dtype).view(self.shape) for p in self._recv_ptr]
        everyone = [None] * n
        dist.all_gather_object(everyone, {"rank": my_rank, "recv": recv_handles, "flags": flags_handle}, group=group)
        self._opened = []
        # peer at distance d downstream (me + d): I write into its recv[(n - d) - 1]... see slot numbering above
        self._peer_recv, self._peer_flags = {}, {}
        for d in range(1, n):
            peer = everyone[(self.me + d) % n]
            self._peer_recv[d] = self._map(peer["recv"][d - 1])     # on the peer I am source (peer - d): slot d - 1
            self._peer_flags[d] = self._map(peer["flags"])
        self.next_id = 1
        self.last_sent = 0
        dist.barrier(group=group)  # every mapping exists before anyone pushes

    def _map(self, handle: bytes) -> int:
        p = _open(handle)
        self._opened.append(p)
        return p

    def push_all(self, src: torch.Tensor, comm_stream) -> int:
        """Enqueue on `comm_stream`: send the local shard to every peer, nearest consumer first.  Returns the transfer
        id T the receivers wait for."""
        assert src.is_contiguous() and src.numel() * src.element_size() == self.nbytes
        T = self.next_id
        self.next_id += 1
        s = comm_stream.cuda_stream
        for d in range(1, self.n):
            if self.last_sent:
                # the peer at distance d reports consumption of my previous shard in MY free[d - 1]
                check(self.lib.ld_stream_wait_geq_u32(self._flags_ptr + FREE_OFF + 4 * (d - 1), self.last_sent, s),
                      "ld_stream_wait_geq_u32")
            check(self.lib.ld_copy_async(self._peer_recv[d], src.data_ptr(), self.nbytes, s), "ld_copy_async")
            check(self.lib.ld_stream_write_u32(self._peer_flags[d] + 4 * (d - 1), T, s), "ld_stream_write_u32")
        self.last_sent = T
        return T

    def kernel_shards(self, kv_local: torch.Tensor, T: int):
        """Shard list for `ops.attention_shards`: the local shard, then recv[0], recv[1], ... each guarded by its
        arrival flag (the kernel polls it)."""
        out = [(kv_local[0], kv_local[1], None)]
        for j in range(1, self.n):
            buf = self.recv[j - 1]
            out.append((buf[0], buf[1], None, self._flags_ptr + 4 * (j - 1), T))
        return out

    def wait_all(self, T: int, stream) -> None:
        """Enqueue on `stream`: stream-level waits (no kernel) until every source's transfer T has landed — for consumers
        that are plain kernels (the attention kernel polls the flags itself instead)."""
        for j in range(1, self.n):
            check(self.lib.ld_stream_wait_geq_u32(self._flags_ptr + 4 * (j - 1), T, stream.cuda_stream), "ld_stream_wait_geq_u32")

    def source_rank_of_slot(self, j: int) -> int:
        """Index (within the group) of the rank whose block sits in recv[j - 1]: the rank at distance j upstream."""
        return (self.me - j) % self.n

    def release_all(self, T: int, stream) -> None:
        """Enqueue on `stream` (after the attention kernel that read the receive buffers): tell every source its shard
        has been consumed.  The source at distance j upstream is the peer at distance n - j downstream; on it I am the
        destination at distance j, i.e. its free[j - 1]."""
        for j in range(1, self.n):
            check(self.lib.ld_stream_write_u32(self._peer_flags[self.n - j] + FREE_OFF + 4 * (j - 1), T, stream.cuda_stream),
                  "ld_stream_write_u32")

    def close(self):
        torch.cuda.synchronize(self.device)
        for p in self._opened:
            self.lib.ld_ipc_close(p)
        self._opened = []
        self.recv = []
        for p in self._recv_ptr + [self._flags_ptr]:
            self.lib.ld_ipc_free(p)
        self._recv_ptr = []
