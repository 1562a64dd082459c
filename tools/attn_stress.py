"""Hang / race stress of the attention kernels: many launches at the full shape and at ragged shapes with fresh inputs,
each checked for finiteness and (a subset) against torch SDPA.  Run under `timeout`."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from landiff_b200 import ops

dev = "cuda"
torch.manual_seed(0)
random.seed(0)
shapes = [(2, 30, 17776, 17776)] * 12 + [(1, 30, 4444, 4444)] * 10 + [(1, 30, 8888, 8888)] * 6
shapes += [(1, random.randint(1, 4), random.randint(1, 700), random.randint(1, 700)) for _ in range(60)]
worst = 0.0
for it, (B, H, nq, nkv) in enumerate(shapes):
    q = torch.randn(B, H, nq, 64, device=dev).bfloat16()
    k = torch.randn(B, H, nkv, 64, device=dev).bfloat16()
    v = torch.randn(B, H, nkv, 64, device=dev).bfloat16()
    for variant in (0, 1):
        out = ops.attention(q, k, v, variant=variant)
        assert torch.isfinite(out.float()).all(), (it, variant, B, H, nq, nkv)
        if nq * nkv <= 1 << 20 or it in (0, 12, 22):
            ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(B, nq, H * 64)
            r = ((out.float() - ref.float()).norm() / ref.float().norm()).item()
            worst = max(worst, r)
            assert r < 1e-2, (it, variant, B, H, nq, nkv, r)
torch.cuda.synchronize()
print(f"attn_stress ok: {len(shapes)} shapes x 2 variants, worst rel-L2 vs torch SDPA (bf16) {worst:.3e}")
