"""Multi-GPU partitioning of the denoise step on one 8xB200 box (one process per GPU, torch.distributed/NCCL).

Two orthogonal axes:
  * CFG-parallel (x2): the uncond / cond batch rows never mix inside the network (guiders.py:46-55 builds them,
    nothing in dit_video_concat.py crosses batch rows), so each half of the ranks evaluates ONE row; per step the
    bf16 network outputs are exchanged (1.1 MB) and every rank applies the fused CFG + DPM++ update redundantly
    with identical RNG state.
  * ring sequence parallelism (x2 / x4) over the token axis: every op except attention is row-local; for attention
    every rank pushes its K|V shard [2, B, H, N/sp, 64] to its peers with copy-engine peer copies over NVLink on a side
    stream, and ONE attention launch walks the local shard and then the peers' shards as they arrive (the kernel polls
    the arrival flags; accumulation stays in TMEM, nothing is merged afterwards).  Ulysses is not applicable (30 heads
    do not divide by 4 or 8).
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
    """Sequence-parallel attention of one rank, attached to a DiffusionTransformer as `.ring` and called with the layer
    workspace.  Every rank needs every peer's K|V shard once per layer; the attention itself is ONE launch over all
    shards (`ops.attention_shards`, accumulating in TMEM — no per-hop launches, no partial-result merges).

    transport "dma" (default; LD_RING_TRANSPORT overrides): the local shard is pushed to every peer by copy-engine peer
    copies into IPC-mapped receive buffers (landiff_b200/dma_ring.py) on a side stream — no SMs, which matters because
    the attention kernel leaves none free — and the kernel polls the per-shard arrival flags itself, so it starts on the
    local shard immediately and the transfers hide behind it.
    transport "nccl": one `all_gather_into_tensor` of the shards on the side stream, the launch waits for all of it
    (portable fallback without CUDA IPC; the transfer is not overlapped)."""

    def __init__(self, layout: Layout, sp_group, device, transport: Optional[str] = None):
        import os

        self.layout, self.group, self.device = layout, sp_group, device
        self.transport = (transport or os.environ.get("LD_RING_TRANSPORT", "dma")).lower()
        if self.transport not in ("nccl", "dma"):
            raise ValueError(f"unknown ring transport {self.transport!r}")
        if not 2 <= layout.sp_size <= 4:
            raise ValueError(f"sequence-parallel groups of 2..4 ranks are supported, got {layout.sp_size}")
        self.comm_stream = torch.cuda.Stream(device=device, priority=-1)
        self._bufs = {}
        self._peer = {}

    def _peer_gather(self, kv):
        key = (tuple(kv.shape), kv.dtype)
        pg = self._peer.get(key)
        if pg is None:
            from .dma_ring import PeerGather

            pg = PeerGather(self.group, self.layout.sp_group_ranks(), self.layout.rank, kv.shape, kv.dtype, kv.device)
            self._peer[key] = pg
        return pg

    def attention(self, ws, variant: int = 0):
        if self.transport == "dma":
            return self._attention_dma(ws, variant)
        from . import ops

        sp, me = self.layout.sp_size, self.layout.sp_rank
        kv = ws["kv"]
        gathered = self._bufs.get(id(ws))
        if gathered is None:
            gathered = self._bufs[id(ws)] = torch.empty((sp,) + tuple(kv.shape), dtype=kv.dtype, device=kv.device)
        compute = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(compute)              # local K|V written by the QKV GEMM
        self.comm_stream.wait_event(ready)
        with torch.cuda.stream(self.comm_stream):
            dist.all_gather_into_tensor(gathered.view(-1), kv.view(-1), group=self.group)
        done = torch.cuda.Event()
        done.record(self.comm_stream)
        compute.wait_event(done)
        order = [(me - j) % sp for j in range(sp)]      # same key order as the dma transport: local, rank-1, rank-2, ...
        ops.attention_shards(ws["q"], [(gathered[r][0], gathered[r][1]) for r in order], out=ws["attn"], variant=variant)

    def _attention_dma(self, ws, variant: int):
        from . import ops

        kv = ws["kv"]
        pg = self._peer_gather(kv)
        compute = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(compute)              # local K|V written by the QKV GEMM
        self.comm_stream.wait_event(ready)
        T = pg.push_all(kv, self.comm_stream)
        ops.attention_shards(ws["q"], pg.kernel_shards(kv, T), out=ws["attn"], variant=variant)
        pg.release_all(T, compute)
        # the QKV GEMM of the next layer overwrites ws["kv"]: the pushes must have read it
        done = torch.cuda.Event()
        done.record(self.comm_stream)
        compute.wait_event(done)


def _all_reduce_sum(buf: torch.Tensor, group) -> None:
    """Sum all-reduce in place.  gloo (CPU tests, single-GPU multi-process tests) has no bf16 reduction: stage through
    fp32 there — every element is x + 0 + ... + 0, exact in either type."""
    if buf.dtype == torch.bfloat16 and dist.get_backend(group) == "gloo":
        tmp = buf.float()
        dist.all_reduce(tmp, group=group)
        buf.copy_(tmp)
    else:
        dist.all_reduce(buf, group=group)


class OutputGather:
    """Exchange of the network outputs under CFG / sequence parallelism without an all-reduce: every rank's final linear
    leaves its (batch row, image-token shard) block token-major in a fixed-size buffer; the copy engines push it into an
    IPC-mapped receive buffer on every other rank (the same `PeerGather` protocol as the K|V exchange: monotonically
    increasing transfer ids, ready / free flags, no SMs, no host synchronisation); each rank then waits for the arrival
    flags with stream memory operations and ONE small kernel scatters all blocks into the latent layout
    (`ops.unpatchify_blocks`).  Replaces zero-fill + mask + a 4.5 MB all-reduce (1.1 ms on 8 GPUs) per sampler step."""

    def __init__(self, layout: Layout, world_group, device, rows_local: int, tok_rows: int):
        from .dma_ring import PeerGather

        self.layout = layout
        self.rows_local, self.tok_rows = rows_local, tok_rows
        self.pg = PeerGather(world_group, list(range(layout.world)), layout.rank, (rows_local, tok_rows, 64),
                             torch.bfloat16, device, max_ranks=16)
        self.comm_stream = torch.cuda.Stream(device=device, priority=-1)
        self._sent = None     # event: the previous push has read the send buffer

    def blocks_of(self, rank: int, base, n_total: int, text_len: int):
        """(address, row, g0, count) of the blocks inside rank `rank`'s buffer at `base` (a tensor or a device address)."""
        lay = self.layout
        s = rank % lay.sp_size
        start, count = shard_bounds(n_total, lay.sp_size, s)
        g0 = max(start - text_len, 0)
        n_img = start + count - max(start, text_len)
        addr = base.data_ptr() if isinstance(base, torch.Tensor) else int(base)
        rows = [rank // lay.sp_size] if lay.cfg_size == 2 else list(range(self.rows_local))
        return [(addr + i * self.tok_rows * 64 * 2, r, g0, n_img) for i, r in enumerate(rows)]

    def gather(self, tok_local: torch.Tensor, out: torch.Tensor, n_total: int, text_len: int) -> torch.Tensor:
        """tok_local: this rank's [rows_local, tok_rows, 64] block (persistent buffer); out: [2, T, 16, H, W] bf16."""
        assert tuple(tok_local.shape) == (self.rows_local, self.tok_rows, 64) and tok_local.is_contiguous()
        compute = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(compute)                       # the final GEMM has written the block
        self.comm_stream.wait_event(ready)
        T = self.pg.push_all(tok_local, self.comm_stream)
        self._sent = torch.cuda.Event()
        self._sent.record(self.comm_stream)
        blocks = self.blocks_of(self.layout.rank, tok_local, n_total, text_len)
        for j in range(1, self.pg.n):
            blocks += self.blocks_of(self.pg.source_rank_of_slot(j), self.pg.recv[j - 1], n_total, text_len)
        self.pg.wait_all(T, compute)
        from . import ops

        ops.unpatchify_blocks(blocks, out)
        self.pg.release_all(T, compute)
        compute.wait_event(self._sent)              # the next step's final GEMM overwrites tok_local
        return out

    def close(self):
        self.pg.close()


class CFGGroup:
    """Evaluates one CFG batch row per half of the ranks and exchanges the bf16 outputs (and, under sequence
    parallelism, assembles the token shards) with ONE all-reduce over all ranks: every rank writes its
    (row, shard) contribution into a zeroed [2, T, C, H, W] buffer; 0 + x is exact, so the sum is the assembly."""

    def __init__(self, layout: Layout, world_group=None):
        self.layout, self.group = layout, world_group
        self._buf = None
        self._gather = None   # OutputGather, created on the first token-major network output

    def _gather_tokens(self, network, net: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        """net: the network's token-major block [rows_local, tok_rows, 64] (see DiffusionTransformer._final)."""
        m = network.main_model.diffusion_model if hasattr(network, "main_model") else network
        n_total = m.text_length + x.shape[1] * (x.shape[3] // 2) * (x.shape[4] // 2)
        if self._gather is None:
            self._gather = OutputGather(self.layout, self.group, x.device, net.shape[0], net.shape[1])
        shape = (2,) + tuple(x.shape[1:])
        if self._buf is None or self._buf.shape != shape or self._buf.dtype != torch.bfloat16:
            self._buf = torch.empty(shape, dtype=torch.bfloat16, device=x.device)
        return self._gather.gather(net, self._buf, n_total, m.text_length)

    def my_row(self) -> int:
        return self.layout.cfg_rank if self.layout.cfg_size == 2 else -1

    def close(self) -> None:
        """Release the IPC buffers of the copy-engine output exchange (collective-free; call before the process group goes)."""
        if self._gather is not None:
            self._gather.close()
            self._gather = None

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
            _all_reduce_sum(buf, self.group)
        return buf

    def evaluate(self, network, x, timestep: float, ctx2, **kwargs):
        lay = self.layout
        if lay.cfg_size == 2:
            row = lay.cfg_rank
            t1 = torch.full((1,), timestep, dtype=torch.float32, device=x.device)
            net = network(x, t1, {"crossattn": ctx2[row:row + 1]}, idx=t1, **kwargs)
            if net.dim() == 3:        # token-major block: peer-copy exchange instead of the all-reduce
                buf = self._gather_tokens(network, net, x)
                return buf[0:1], buf[1:2]
            mask = getattr(network, "owned_latent_mask", None)
            buf = self.assemble(net, row, mask(x) if mask is not None else None)
            return buf[0:1], buf[1:2]
        t2 = torch.full((2,), timestep, dtype=torch.float32, device=x.device)
        net = network(torch.cat([x, x]), t2, {"crossattn": ctx2}, idx=t2, **kwargs)
        if net.dim() == 3:
            buf = self._gather_tokens(network, net, x)
            return buf[0:1], buf[1:2]
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
            _all_reduce_sum(buf, self.group)
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

    for wrapper in (warp.control_model, warp.main_model):
        m = wrapper.diffusion_model
        m.ring = RingAttention(layout, sp_group, device)
        m.sp_layout = layout
    # with the copy-engine transport the network output is exchanged the same way (CFGGroup -> OutputGather): the final
    # linear leaves its block token-major; LD_OUTPUT_EXCHANGE=allreduce keeps the all-reduce assembly (A/B, NCCL-only setups)
    import os

    if warp.main_model.diffusion_model.ring.transport == "dma" and os.environ.get("LD_OUTPUT_EXCHANGE", "dma") == "dma":
        warp.main_model.diffusion_model.token_major_out = True

    def mask_fn(x):
        m = warp.main_model.diffusion_model
        n_total = m.text_length + x.shape[1] * (x.shape[3] // 2) * (x.shape[4] // 2)
        start, count = shard_bounds(n_total, layout.sp_size, layout.sp_rank)
        return owned_latent_mask(tuple(x.shape[1:]), m.text_length, start, count, x.device)

    warp.owned_latent_mask = mask_fn
