"""Multi-GPU partitioning of the denoise step on one 8xB200 box (one process per GPU, torch.distributed/NCCL).

Two orthogonal axes:
  * CFG-parallel (x2): the uncond / cond batch rows never mix inside the network (guiders.py:46-55 builds them,
    nothing in dit_video_concat.py crosses batch rows), so each half of the ranks evaluates ONE row; per step the
    bf16 network outputs are exchanged (1.1 MB) and every rank applies the fused CFG + DPM++ update redundantly
    with identical RNG state.
  * ring sequence parallelism (x2 / x4) over the token axis: every op except attention is row-local; for attention
    the K|V shard [2, B, H, N/sp, 64] rotates around the ring with NCCL send/recv on a side stream, overlapped with
    the attention tiles of the shard already present, and the per-shard partial results are merged by their
    log-sum-exp (`ld_attention_merge`).  Ulysses is not applicable (30 heads do not divide by 4 or 8).
rank = cfg_rank * sp_size + sp_rank.   world 1: none · 2: CFG · 4: CFG x SP2 · 8: CFG x SP4.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


@dataclass
class Layout:
    world: int
    rank: int
    cfg_size: int
    sp_size: int

    @property
    def cfg_rank(self) -> int:
        return self.rank // self.sp_size

    @property
    def sp_rank(self) -> int:
        return self.rank % self.sp_size

    def sp_group_ranks(self, cfg_rank: Optional[int] = None) -> List[int]:
        c = self.cfg_rank if cfg_rank is None else cfg_rank
        return [c * self.sp_size + s for s in range(self.sp_size)]

    def cfg_group_ranks(self, sp_rank: Optional[int] = None) -> List[int]:
        s = self.sp_rank if sp_rank is None else sp_rank
        return [c * self.sp_size + s for c in range(self.cfg_size)]


def make_layout(world: int, rank: int, cfg_parallel: bool = True) -> Layout:
    if world < 1 or rank < 0 or rank >= world:
        raise ValueError(f"bad world/rank {world}/{rank}")
    cfg = 2 if (cfg_parallel and world >= 2) else 1
    if world % cfg:
        raise ValueError(f"world size {world} is not divisible by the CFG group size {cfg}")
    return Layout(world, rank, cfg, world // cfg)


def shard_bounds(n_total: int, sp_size: int, sp_rank: int) -> Tuple[int, int]:
    """Even token shards (17 776 = 16 x 1111 splits evenly by 2, 4, 8, 16).  Returns (start, count)."""
    if n_total % sp_size:
        raise ValueError(f"sequence of {n_total} tokens does not split evenly over {sp_size} ring ranks")
    c = n_total // sp_size
    return sp_rank * c, c


def new_subgroups(layout: Layout):
    """Create (collectively, same order on every rank) the SP ring groups and CFG pair groups; return this rank's."""
    my_sp = my_cfg = None
    for c in range(layout.cfg_size):
        g = dist.new_group(layout.sp_group_ranks(c))
        if c == layout.cfg_rank:
            my_sp = g
    for s in range(layout.sp_size):
        g = dist.new_group(layout.cfg_group_ranks(s))
        if s == layout.sp_rank:
            my_cfg = g
    return my_sp, my_cfg


def ring_schedule(sp_size: int, sp_rank: int) -> List[int]:
    """Which rank's K/V shard this rank holds at hop h when every rank sends to (r+1) and receives from (r-1)."""
    return [(sp_rank - h) % sp_size for h in range(sp_size)]


def ring_attention_generic(q, kv_local, sp_size: int, sp_rank: int, exchange: Callable, attend: Callable, merge: Callable):
    """The ring loop with injected primitives (so the schedule is unit-testable on CPU with gloo):
        exchange(send_buf) -> recv_buf            rotate one hop (may be asynchronous; returns a handle via .wait())
        attend(q, kv) -> (o, lse)                 partial attention against one shard
        merge((o, lse), (o2, lse2)) -> (o, lse)   log-sum-exp merge
    """
    cur = kv_local
    acc = None
    for hop in range(sp_size):
        pending = exchange(cur) if hop < sp_size - 1 else None
        part = attend(q, cur)
        acc = part if acc is None else merge(acc, part)
        if pending is not None:
            cur = pending.wait() if hasattr(pending, "wait") else pending
    return acc


class RingAttention:
    """GPU ring: attached to a DiffusionTransformer as `.ring`; called with the layer workspace.

    transport "nccl" (default): batch_isend_irecv on a high-priority side stream.  transport "dma" (or
    LD_RING_TRANSPORT=dma): copy-engine peer copies into IPC-mapped buffers ordered by stream memory operations
    (landiff_b200/dma_ring.py) — no SMs, which matters because the attention kernel leaves none free."""

    def __init__(self, layout: Layout, sp_group, device, transport: Optional[str] = None):
        import os

        self.layout, self.group, self.device = layout, sp_group, device
        self.transport = (transport or os.environ.get("LD_RING_TRANSPORT", "nccl")).lower()
        if self.transport not in ("nccl", "dma"):
            raise ValueError(f"unknown ring transport {self.transport!r}")
        # High priority: the attention kernel fills every SM (640 threads x ~100 registers), so NCCL's send/recv CTAs only
        # run when an SM drains; with priority they take the first free slots instead of queueing behind the
        # remaining attention CTAs (measured with the default priority: 34 MB hops at 129 GB/s, not hidden at sp = 4).
        self.comm_stream = torch.cuda.Stream(device=device, priority=-1)
        ranks = layout.sp_group_ranks()
        self.next_rank = ranks[(layout.sp_rank + 1) % layout.sp_size]
        self.prev_rank = ranks[(layout.sp_rank - 1) % layout.sp_size]
        self._bufs = {}
        self._peer = {}

    def _buffers(self, ws):
        key = id(ws)
        b = self._bufs.get(key)
        if b is None:
            kv = ws["kv"]
            B, H, R = ws["q"].shape[:3]
            f = lambda *s: torch.empty(*s, dtype=torch.float32, device=kv.device)
            b = dict(o_acc=f(B * H, R, 64), lse_acc=f(B * H, R), o_new=f(B * H, R, 64), lse_new=f(B * H, R))
            if self.transport == "nccl":
                b["ring"] = [torch.empty_like(kv), torch.empty_like(kv)]
            self._bufs[key] = b
        return b

    def _peer_ring(self, kv):
        key = (tuple(kv.shape), kv.dtype)
        pr = self._peer.get(key)
        if pr is None:
            from .dma_ring import PeerRing

            pr = PeerRing(self.group, self.layout.sp_group_ranks(), self.layout.rank, kv.shape, kv.dtype, kv.device)
            self._peer[key] = pr
        return pr

    def attention(self, ws, variant: int = 0):
        if self.transport == "dma":
            return self._attention_dma(ws, variant)
        from . import ops

        sp = self.layout.sp_size
        b = self._buffers(ws)
        q, out = ws["q"], ws["attn"]
        B, H, R = q.shape[:3]
        cur = ws["kv"]
        compute = torch.cuda.current_stream()
        for hop in range(sp):
            reqs = None
            nxt = None
            if hop < sp - 1:
                nxt = b["ring"][hop % 2]
                ready = torch.cuda.Event()
                ready.record(compute)  # `cur` is complete on the compute stream (QKV GEMM / previous receive)
                self.comm_stream.wait_event(ready)
                with torch.cuda.stream(self.comm_stream):
                    ops_ = [dist.P2POp(dist.isend, cur, self.next_rank, group=self.group),
                            dist.P2POp(dist.irecv, nxt, self.prev_rank, group=self.group)]
                    reqs = dist.batch_isend_irecv(ops_)
            if hop == 0:
                ops.attention(q, cur[0], cur[1], out=out, lse=b["lse_acc"], out_f32=b["o_acc"], variant=variant)
            else:
                ops.attention(q, cur[0], cur[1], out=out, lse=b["lse_new"], out_f32=b["o_new"], variant=variant)
                ops.attention_merge(b["o_acc"], b["lse_acc"], b["o_new"], b["lse_new"], out if hop == sp - 1 else None,
                                    B, H, R)
            if reqs is not None:
                with torch.cuda.stream(self.comm_stream):
                    for r in reqs:
                        r.wait()
                done = torch.cuda.Event()
                done.record(self.comm_stream)
                compute.wait_event(done)
                cur = nxt

    def _attention_dma(self, ws, variant: int):
        """Same schedule with the copy-engine transport: at hop h this rank forwards the shard it holds into the
        downstream rank's recv[h % 2] while it attends to it; the shard for hop h+1 arrives in its own recv[h % 2]."""
        from . import ops

        sp = self.layout.sp_size
        b = self._buffers(ws)
        q, out = ws["q"], ws["attn"]
        B, H, R = q.shape[:3]
        kv = ws["kv"]
        pr = self._peer_ring(kv)
        compute = torch.cuda.current_stream()
        cur, cur_j, cur_T = kv, -1, 0     # shard held at this hop; the recv buffer it lives in (-1: local kv)
        for hop in range(sp):
            sent_T = None
            if hop < sp - 1:
                j = hop % 2
                if hop == 0:
                    ready = torch.cuda.Event()
                    ready.record(compute)        # local K|V written by the QKV GEMM
                    self.comm_stream.wait_event(ready)
                else:
                    pr.wait_arrival(cur_j, cur_T, self.comm_stream)   # forward as soon as it has landed
                sent_T = pr.push(cur, j, self.comm_stream)
            if hop > 0:
                pr.wait_arrival(cur_j, cur_T, compute)
            if hop == 0:
                ops.attention(q, cur[0], cur[1], out=out, lse=b["lse_acc"], out_f32=b["o_acc"], variant=variant)
            else:
                ops.attention(q, cur[0], cur[1], out=out, lse=b["lse_new"], out_f32=b["o_new"], variant=variant)
                ops.attention_merge(b["o_acc"], b["lse_acc"], b["o_new"], b["lse_new"], out if hop == sp - 1 else None,
                                    B, H, R)
            if hop < sp - 1:
                # the next shard to attend to arrives in my recv[hop % 2] with the same id my own push carries
                # (every rank numbers its transfers identically)
                nxt_j, nxt_T = hop % 2, sent_T
            if hop > 0:
                # recv[cur_j] has been read by the attention above AND (if forwarded) by the push on the comm stream
                if hop < sp - 1:
                    fwd = torch.cuda.Event()
                    fwd.record(self.comm_stream)
                    compute.wait_event(fwd)
                pr.release(cur_j, cur_T, compute)
            if hop < sp - 1:
                cur, cur_j, cur_T = pr.recv[nxt_j], nxt_j, nxt_T
        # the QKV GEMM of the next layer overwrites ws["kv"]: the push of hop 0 must have read it
        done = torch.cuda.Event()
        done.record(self.comm_stream)
        compute.wait_event(done)


class CFGGroup:
    """Evaluates one CFG batch row per half of the ranks and exchanges the bf16 outputs (and, under sequence
    parallelism, assembles the token shards) with ONE all-reduce over all ranks: every rank writes its
    (row, shard) contribution into a zeroed [2, T, C, H, W] buffer; 0 + x is exact, so the sum is the assembly."""

    def __init__(self, layout: Layout, world_group=None):
        self.layout, self.group = layout, world_group
        self._buf = None

    def my_row(self) -> int:
        return self.layout.cfg_rank if self.layout.cfg_size == 2 else -1

    def assemble(self, net_local: torch.Tensor, row: int, owned_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """net_local: [1, T, C, H, W] (only this rank's token shard is meaningful when owned_mask is given)."""
        shape = (2,) + tuple(net_local.shape[1:])
        if self._buf is None or self._buf.shape != shape or self._buf.dtype != net_local.dtype:
            self._buf = torch.zeros(shape, dtype=net_local.dtype, device=net_local.device)
        buf = self._buf
        buf.zero_()
        if owned_mask is None:
            buf[row].copy_(net_local[0])
        else:
            buf[row].copy_(torch.where(owned_mask, net_local[0], torch.zeros_like(net_local[0])))
        if self.layout.world > 1:
            dist.all_reduce(buf, group=self.group)
        return buf

    def evaluate(self, network, x, timestep: float, ctx2, **kwargs):
        lay = self.layout
        if lay.cfg_size == 2:
            row = lay.cfg_rank
            t1 = torch.full((1,), timestep, dtype=torch.float32, device=x.device)
            net = network(x, t1, {"crossattn": ctx2[row:row + 1]}, idx=t1, **kwargs)
            mask = getattr(network, "owned_latent_mask", None)
            buf = self.assemble(net, row, mask(x) if mask is not None else None)
            return buf[0:1], buf[1:2]
        t2 = torch.full((2,), timestep, dtype=torch.float32, device=x.device)
        net = network(torch.cat([x, x]), t2, {"crossattn": ctx2}, idx=t2, **kwargs)
        if lay.sp_size > 1:
            # ring SP without CFG parallelism: both rows live here, but only this rank's token shard of each was written
            mask = getattr(network, "owned_latent_mask", None)
            if mask is None:
                raise RuntimeError("sequence-parallel network without `owned_latent_mask` (use landiff_b200.parallel.attach)")
            own = mask(x)
            if self._buf is None or self._buf.shape != net.shape or self._buf.dtype != net.dtype:
                self._buf = torch.zeros(net.shape, dtype=net.dtype, device=net.device)
            buf = self._buf
            buf.copy_(torch.where(own, net, torch.zeros_like(net)))
            dist.all_reduce(buf, group=self.group)
            return buf[0:1], buf[1:2]
        return net[0:1], net[1:2]


def owned_latent_mask(shape, text_len: int, start: int, count: int, device) -> torch.Tensor:
    """Boolean [T, C, H, W] mask of the latent pixels whose 2x2 patch token (text_len + g) lies in this rank's shard."""
    T, C, H, W = shape
    Hp, Wp = H // 2, W // 2
    g = torch.arange(T * Hp * Wp, device=device).view(T, 1, Hp, 1, Wp, 1)
    tok = g + text_len
    own = (tok >= start) & (tok < start + count)
    return own.expand(T, C, Hp, 2, Wp, 2).reshape(T, C, H, W)


def attach(warp, layout: Layout, sp_group, device) -> None:
    """Configure a ControlDiffWarp for this rank: token shard + ring on both networks."""
    if layout.sp_size == 1:
        return
    from .dit import SequenceShard

    for wrapper in (warp.control_model, warp.main_model):
        m = wrapper.diffusion_model
        m.ring = RingAttention(layout, sp_group, device)
        m.sp_layout = layout

    def mask_fn(x):
        m = warp.main_model.diffusion_model
        n_total = m.text_length + x.shape[1] * (x.shape[3] // 2) * (x.shape[4] // 2)
        start, count = shard_bounds(n_total, layout.sp_size, layout.sp_rank)
        return owned_latent_mask(tuple(x.shape[1:]), m.text_length, start, count, x.device)

    warp.owned_latent_mask = mask_fn
