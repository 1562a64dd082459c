"""Functional check of the copy-engine ring transport (landiff_b200/dma_ring.py + csrc/peer_ring.cu) with TWO processes
on ONE GPU (gloo control plane): the IPC mapping, the peer copies and the stream-memory-op ordering are exactly what
the multi-GPU ring uses; only the link underneath differs.  Each rank holds half of the sequence, runs several
ring-attention calls with fresh data and compares its rows with the monolithic attention over the full K/V.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dma_ring_check.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from landiff_b200 import ops, parallel

world, rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"])
multi = os.environ.get("LD_DMA_CHECK_MULTI", "0") == "1"      # 1: one GPU per rank (real NVLink peer copies)
lr = int(os.environ.get("LOCAL_RANK", "0")) if multi else 0
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("gloo")
layout = parallel.Layout(world, rank, 1, world)
ring = parallel.RingAttention(layout, None, dev, transport="dma")
B, H, N = 1, 4, 512 * world
R = N // world
ws = dict(q=torch.empty(B, H, R, 64, device=dev, dtype=torch.bfloat16), kv=torch.empty(2, B, H, R, 64, device=dev, dtype=torch.bfloat16),
          attn=torch.empty(B, R, H * 64, device=dev, dtype=torch.bfloat16))
worst = 0.0
for call in range(8):
    g = torch.Generator(device=dev).manual_seed(100 + call)      # identical full tensors on every rank
    qf = torch.randn(B, H, N, 64, device=dev, generator=g).bfloat16()
    kf = torch.randn(B, H, N, 64, device=dev, generator=g).bfloat16()
    vf = torch.randn(B, H, N, 64, device=dev, generator=g).bfloat16()
    sl = slice(rank * R, (rank + 1) * R)
    ws["q"].copy_(qf[:, :, sl]); ws["kv"][0].copy_(kf[:, :, sl]); ws["kv"][1].copy_(vf[:, :, sl])
    ring.attention(ws)
    if call % 3 == 2:
        torch.cuda._sleep(int(2e7) * (rank + 1))    # skew the ranks: the flags, not luck, must order the hops
    ref = torch.nn.functional.scaled_dot_product_attention(qf[:, :, sl].float(), kf.float(), vf.float())
    ref = ref.permute(0, 2, 1, 3).reshape(B, R, H * 64)
    r = ((ws["attn"].float() - ref).norm() / ref.norm()).item()
    worst = max(worst, r)
torch.cuda.synchronize()
t = torch.tensor([worst])
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"dma_ring_check world={world}: 8 ring-attention calls, worst rel-L2 vs monolithic {t.item():.3e} {'OK' if t.item() < 4e-3 else 'FAIL'}", flush=True)
# hop bandwidth: 68 MB (the sp = 2 shard of the full shape) pushed 20 times into the neighbour, flags included
big = torch.empty(2, 1, 30, 8888, 64, device=dev, dtype=torch.bfloat16)
from landiff_b200.dma_ring import PeerRing
pr2 = PeerRing(None, list(range(world)), rank, big.shape, big.dtype, dev)
st = torch.cuda.Stream(device=dev)
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    with torch.cuda.stream(st):
        e0.record(st)
        for i in range(20):
            T = pr2.push(big, i % 2, st)
            pr2.wait_arrival(i % 2, T, st)       # my upstream pushed the same id into me
            pr2.release(i % 2, T, st)
        e1.record(st)
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
if rank == 0:
    print(f"peer hop: {big.numel() * 2 / 1e6:.1f} MB in {ms:.3f} ms = {big.numel() * 2 / ms / 1e6:.0f} GB/s per direction ({'one GPU per rank' if multi else 'both ranks on one GPU'})", flush=True)
pr2.close()
for pr in ring._peer.values():
    pr.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if t.item() < 4e-3 else 1)
