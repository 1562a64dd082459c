import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from landiff_b200 import _C, ops
dev = "cuda"
torch.manual_seed(8)
B, H, nq, n0, n1 = 1, 2, 256, 192, 320
q = torch.randn(B, H, nq, 64, device=dev).bfloat16()
k = torch.randn(B, H, n0 + n1, 64, device=dev).bfloat16()
v = torch.randn(B, H, n0 + n1, 64, device=dev).bfloat16()
k0, v0 = k[:, :, :n0].contiguous(), v[:, :, :n0].contiguous()
ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float()).permute(0, 2, 1, 3).reshape(B, nq, H * 64)
def rel(a): return ((a.float() - ref).norm() / ref.norm()).item()
flag = torch.zeros(64, device=dev, dtype=torch.int32)
for variant in (0, 1):
    for mode in ("preset", "late", "late_finite_poison", "late_nofill"):
        poison = 1e4 if mode == "late_finite_poison" else float("nan")
        k1 = torch.full((B, H, n1, 64), poison, device=dev, dtype=torch.bfloat16)
        v1 = torch.full((B, H, n1, 64), poison, device=dev, dtype=torch.bfloat16)
        flag.zero_()
        side = torch.cuda.Stream()
        torch.cuda.synchronize()
        if mode == "preset":
            k1.copy_(k[:, :, n0:]); v1.copy_(v[:, :, n0:]); flag.fill_(7)
            torch.cuda.synchronize()
        if mode == "late_nofill":
            k1.copy_(k[:, :, n0:]); v1.copy_(v[:, :, n0:])
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = ops.attention_shards(q, [(k0, v0, None), (k1, v1, None, flag.data_ptr(), 7)], variant=variant)
        e1.record()
        if mode.startswith("late"):
            with torch.cuda.stream(side):
                torch.cuda._sleep(int(2e7))
                if mode != "late_nofill":
                    k1.copy_(k[:, :, n0:]); v1.copy_(v[:, :, n0:])
                _C.check(_C.load().ld_stream_write_u32(flag.data_ptr(), 7, side.cuda_stream), "w")
        torch.cuda.synchronize()
        print(f"variant {variant} {mode:20s}: kernel {e0.elapsed_time(e1):8.3f} ms  nan {int(torch.isnan(out.float()).sum())}/{out.numel()}  rel {rel(out):.3e}  status {ops.attention_status()}", flush=True)
