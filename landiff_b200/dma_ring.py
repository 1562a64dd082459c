"""Copy-engine ring transport for the sequence-parallel attention hop (csrc/peer_ring.cu).

Each rank owns two receive buffers for a K|V shard and a small flag array, both allocated by the library and exported
with CUDA IPC.  At set-up every rank maps its DOWNSTREAM neighbour's buffers and flags (it pushes into them) and its
UPSTREAM neighbour's flags (it reports consumption there).  A hop is then, entirely stream-ordered:

    sender   (comm stream)    wait  own.free[j]   >= id of the previous transfer into down.recv[j]   (buffer consumed)
                              copy  cur -> down.recv[j]          cudaMemcpyAsync peer copy: DMA engines, no SMs
                              write down.ready[j]  = T           (cuStreamWriteValue32; fences the copy)
    receiver (compute stream) wait  own.ready[j]  >= T           before the attention that reads recv[j]
                              write up.free[j]     = T           after that attention

Transfer ids T increase monotonically and identically on every rank (one per hop of every ring-attention call), so
nothing is ever reset and there is no host synchronisation.  Control plane: one `all_gather_object` of the IPC handles
over whatever process group the caller has (NCCL or gloo).
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch
import torch.distributed as dist

from . import _C
from .ops import check

FLAG_BYTES = 256   # ready[0], ready[1], free[0], free[1] (u32), padded


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


class PeerRing:
    def __init__(self, group, group_ranks: List[int], my_rank: int, shard_shape, dtype, device):
        self.lib = _C.load()
        self.device = device
        n = len(group_ranks)
        me = group_ranks.index(my_rank)
        self.nbytes = int(torch.empty(shard_shape, dtype=dtype, device="meta").numel()) * torch.empty((), dtype=dtype).element_size()
        self.shape, self.dtype = tuple(shard_shape), dtype
        # own resources
        self._recv_ptr, self._recv_handles = [], []
        for _ in range(2):
            p, h = _alloc(self.nbytes)
            self._recv_ptr.append(p)
            self._recv_handles.append(h)
        self._flags_ptr, flags_handle = _alloc(FLAG_BYTES)
        self.recv = [torch.as_tensor(_Raw(p, self.nbytes), device=device).view(dtype).view(self.shape) for p in self._recv_ptr]
        # exchange handles
        mine = {"rank": my_rank, "recv": self._recv_handles, "flags": flags_handle}
        everyone = [None] * n
        dist.all_gather_object(everyone, mine, group=group)
        down, up = everyone[(me + 1) % n], everyone[(me - 1) % n]
        self._opened = []
        self._down_recv = [self._map(h) for h in down["recv"]]
        self._down_flags = self._map(down["flags"])
        self._up_flags = self._down_flags if up["rank"] == down["rank"] else self._map(up["flags"])
        self.next_id = 1
        self.last_sent = [0, 0]   # id of the last transfer pushed into down.recv[j]
        dist.barrier(group=group)  # every mapping exists before anyone pushes

    def _map(self, handle: bytes) -> int:
        p = _open(handle)
        self._opened.append(p)
        return p

    # flag addresses: ready[j] at 4 j, free[j] at 8 + 4 j
    def push(self, src: torch.Tensor, j: int, comm_stream) -> int:
        """Enqueue on `comm_stream`: wait until the downstream rank has consumed recv[j], copy `src` into it, publish
        the transfer id.  Returns the id (the receiver waits for the same number)."""
        assert src.is_contiguous() and src.numel() * src.element_size() == self.nbytes
        T = self.next_id
        self.next_id += 1
        s = comm_stream.cuda_stream
        if self.last_sent[j]:
            check(self.lib.ld_stream_wait_geq_u32(self._flags_ptr + 8 + 4 * j, self.last_sent[j], s), "ld_stream_wait_geq_u32")
        check(self.lib.ld_copy_async(self._down_recv[j], src.data_ptr(), self.nbytes, s), "ld_copy_async")
        check(self.lib.ld_stream_write_u32(self._down_flags + 4 * j, T, s), "ld_stream_write_u32")
        self.last_sent[j] = T
        return T

    def wait_arrival(self, j: int, T: int, stream) -> None:
        """Enqueue on `stream`: block until transfer T has landed in recv[j]."""
        check(self.lib.ld_stream_wait_geq_u32(self._flags_ptr + 4 * j, T, stream.cuda_stream), "ld_stream_wait_geq_u32")

    def release(self, j: int, T: int, stream) -> None:
        """Enqueue on `stream` (after the kernels that read recv[j]): tell the upstream rank the buffer is free."""
        check(self.lib.ld_stream_write_u32(self._up_flags + 8 + 4 * j, T, stream.cuda_stream), "ld_stream_write_u32")

    def close(self):
        torch.cuda.synchronize(self.device)
        for p in self._opened:
            self.lib.ld_ipc_close(p)
        self._opened = []
        self.recv = []
        for p in self._recv_ptr + [self._flags_ptr]:
            self.lib.ld_ipc_free(p)
        self._recv_ptr = []
