"""CPU (gloo, world_size 2): host-side logic of the multi-GPU partitioning — rank layout, token shards, the ring
attention schedule (generic loop with injected primitives == monolithic attention), CFG row exchange/assembly and the
owned-latent masks.  The GPU kernels themselves are covered by the `-m gpu` tests."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from landiff_b200 import parallel as P


def test_layouts():
    for world, (cfg, sp) in {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (2, 4)}.items():
        seen = set()
        for r in range(world):
            lay = P.make_layout(world, r)
            assert (lay.cfg_size, lay.sp_size) == (cfg, sp)
            assert lay.rank == lay.cfg_rank * sp + lay.sp_rank
            assert r in lay.sp_group_ranks() and r in lay.cfg_group_ranks()
            assert len(lay.sp_group_ranks()) == sp and len(lay.cfg_group_ranks()) == cfg
            seen.add((lay.cfg_rank, lay.sp_rank))
        assert len(seen) == world
    with pytest.raises(ValueError):
        P.make_layout(3, 0)


def test_shard_bounds_full_shape():
    n = 226 + 13 * 30 * 45
    assert n == 17776
    for sp in (1, 2, 4, 8, 16):
        spans = [P.shard_bounds(n, sp, r) for r in range(sp)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        for (s0, c0), (s1, _) in zip(spans, spans[1:]):
            assert s0 + c0 == s1
    with pytest.raises(ValueError):
        P.shard_bounds(n, 3, 0)


def test_ring_schedule_visits_every_shard_once():
    for sp in (1, 2, 4):
        for r in range(sp):
            assert sorted(P.ring_schedule(sp, r)) == list(range(sp))
            assert P.ring_schedule(sp, r)[0] == r


def test_owned_latent_masks_partition_the_latent():
    T, C, H, W, TL = 2, 16, 8, 12, 6
    n = TL + T * (H // 2) * (W // 2)
    for sp in (1, 2, 3):
        if n % sp:
            continue
        total = torch.zeros(T, C, H, W, dtype=torch.int32)
        for r in range(sp):
            s, c = P.shard_bounds(n, sp, r)
            total += P.owned_latent_mask((T, C, H, W), TL, s, c, "cpu").int()
        assert bool((total == 1).all())


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _attend(q, kv):
    k, v = kv[0], kv[1]
    s = (q @ k.transpose(-1, -2)) / math.sqrt(q.shape[-1])
    lse = torch.logsumexp(s, -1)
    return torch.softmax(s, -1) @ v, lse


def _merge(a, b):
    (oa, la), (ob, lb) = a, b
    m = torch.maximum(la, lb)
    wa, wb = torch.exp(la - m), torch.exp(lb - m)
    o = (oa * wa[..., None] + ob * wb[..., None]) / (wa + wb)[..., None]
    return o, m + torch.log(wa + wb)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)  # identical full tensors on every rank
        B, H, N, D = 1, 2, 24, 64
        q_full, k_full, v_full = torch.randn(B, H, N, D), torch.randn(B, H, N, D), torch.randn(B, H, N, D)
        # --- ring over the whole world (sp = world)
        start, count = P.shard_bounds(N, world, rank)
        q = q_full[:, :, start:start + count]
        kv = torch.stack([k_full[:, :, start:start + count], v_full[:, :, start:start + count]]).contiguous()
        nxt, prv = (rank + 1) % world, (rank - 1) % world

        def exchange(buf):
            recv = torch.empty_like(buf)
            reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, buf, nxt), dist.P2POp(dist.irecv, recv, prv)])
            for r in reqs:
                r.wait()
            return recv

        o, lse = P.ring_attention_generic(q, kv, world, rank, exchange, _attend, _merge)
        ref, ref_lse = _attend(q_full, torch.stack([k_full, v_full]))
        err = (o - ref[:, :, start:start + count]).abs().max().item()
        err_l = (lse - ref_lse[:, :, start:start + count]).abs().max().item()
        # --- CFG assembly: rank r contributes batch row r, all-reduce assembles [2, ...]
        lay = P.make_layout(world, rank)
        grp = P.CFGGroup(lay)
        rows = torch.randn(2, 3, 4, 6, 8)
        buf = grp.assemble(rows[lay.cfg_rank:lay.cfg_rank + 1], lay.cfg_rank)
        err_c = (buf - rows).abs().max().item()
        # --- ring SP WITHOUT CFG parallelism: evaluate() must assemble both rows from the ranks' token shards
        lay1 = P.make_layout(world, rank, cfg_parallel=False)
        assert (lay1.cfg_size, lay1.sp_size) == (1, world)
        T, C, Hh, Ww, TL = 2, 4, 4, 6, 4
        n_tok = TL + T * (Hh // 2) * (Ww // 2)
        s0, cnt = P.shard_bounds(n_tok, world, rank)
        own = P.owned_latent_mask((T, C, Hh, Ww), TL, s0, cnt, "cpu")
        x = torch.randn(1, T, C, Hh, Ww)

        def sharded_net(x2, t2, c, idx=None):
            full = x2 * 2.0 + t2.view(-1, 1, 1, 1, 1)
            return torch.where(own, full, torch.full_like(full, float("nan")))   # rows outside the shard: garbage

        sharded_net.owned_latent_mask = lambda xx: own
        u, c_ = P.CFGGroup(lay1).evaluate(sharded_net, x, 5.0, torch.zeros(2, 3, 8))
        err_sp = max((u - (x * 2 + 5)).abs().max().item(), (c_ - (x * 2 + 5)).abs().max().item())
        ret[rank] = (err, err_l, err_c, err_sp)
    finally:
        dist.destroy_process_group()


def test_ring_and_cfg_exchange_gloo_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        err, err_l, err_c, err_sp = ret[r]
        assert err_sp == 0.0, f"rank {r}: SP-only output assembly wrong ({err_sp})"
        assert err < 1e-5 and err_l < 1e-5, f"rank {r}: ring attention differs from monolithic ({err}, {err_l})"
        assert err_c == 0.0, f"rank {r}: CFG assembly not exact"


def test_output_gather_blocks_cover_every_token_once():
    """Host logic of the copy-engine output exchange: for every layout the (row, g0, count) blocks of all ranks tile the
    [2, n_img] output exactly once, and a rank's blocks sit at the right offsets of its buffer."""
    from landiff_b200.parallel import Layout, OutputGather

    text_len, n_img = 226, 17550
    n_total = text_len + n_img
    for world, cfg in [(2, 1), (4, 2), (8, 2), (4, 1), (16, 2)]:
        sp = world // cfg
        tok_rows = n_total // sp
        cover = torch.zeros(2, n_img, dtype=torch.int32)
        for rank in range(world):
            og = OutputGather.__new__(OutputGather)
            og.layout = Layout(world, 0, cfg, sp)
            og.rows_local, og.tok_rows = (1 if cfg == 2 else 2), tok_rows
            blocks = og.blocks_of(rank, 1 << 20, n_total, text_len)
            assert len(blocks) == og.rows_local
            for i, (addr, row, g0, count) in enumerate(blocks):
                assert addr == (1 << 20) + i * tok_rows * 64 * 2 and 0 < count <= tok_rows
                assert row == (rank // sp if cfg == 2 else i)
                cover[row, g0:g0 + count] += 1
        assert bool((cover == 1).all()), (world, cfg)
