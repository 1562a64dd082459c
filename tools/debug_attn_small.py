import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from landiff_b200 import ops
dev = "cuda"
torch.manual_seed(1)
for (B, H, nq, nkv) in [(1, 1, 128, 128), (1, 1, 256, 128), (1, 2, 300, 300), (2, 3, 886, 886), (1, 4, 4444, 17776)]:
    q = torch.randn(B, H, nq, 64, device=dev).bfloat16()
    k = torch.randn(B, H, nkv, 64, device=dev).bfloat16()
    v = torch.randn(B, H, nkv, 64, device=dev).bfloat16()
    out = ops.attention(q, k, v, variant=int(os.environ.get("VAR", "0")))
    torch.cuda.synchronize()
    ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float()).permute(0, 2, 1, 3).reshape(B, nq, H * 64)
    print(B, H, nq, nkv, "rel", ((out.float() - ref).norm() / ref.norm()).item(), flush=True)
