"""Per-phase cycle breakdown of the attention kernels' softmax warps and MMA issuers (PROF template variants).
usage: python tools/attn_phase_prof.py [N]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from landiff_b200 import _C, ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 17776
B, H = 1, 30
dev = "cuda"
q = torch.randn(B, H, N, 64, device=dev).bfloat16()
k = torch.randn(B, H, N, 64, device=dev).bfloat16()
v = torch.randn(B, H, N, 64, device=dev).bfloat16()
out = torch.empty(B, N, H * 64, device=dev, dtype=torch.bfloat16)
grid = B * H * ((N + 255) // 256)
VAR = int(os.environ.get("LD_ATTN_VARIANT", "0"))   # 0: attn4_kernel, 1: attn3_kernel
NW = 20
FIRST, LAST, MMAW = (0, 16, 19) if VAR == 0 else (4, NW, 1)
prof_all = torch.zeros(grid * NW * 8, device=dev, dtype=torch.int64)
prof = prof_all[:grid * NW * 8].view(grid, NW, 8)
lib = _C.load()
lib.ld_debug_attn_prof.argtypes = [ctypes.c_void_p]
lib.ld_debug_attn_prof.restype = None
ops.attention(q, k, v, out=out, variant=VAR)
torch.cuda.synchronize()
lib.ld_debug_attn_prof(prof.data_ptr())
ops.attention(q, k, v, out=out, variant=VAR)
torch.cuda.synchronize()
lib.ld_debug_attn_prof(None)
n_sub = (N + 63) // 64
per_warp = n_sub / 2   # every softmax warp serves the sub-blocks of one parity
p = prof.double().cpu()
sm = p[:, FIRST:LAST, :5].mean(dim=(0, 1)) / per_warp
names = ["wait s_full", "tcgen05.ld", "mask+max+rescale", "exp+sum+pack", "st+fence+arrive"]
print(f"variant {VAR} N={N} n_sub={n_sub}: softmax warp cycles per 64-key sub-block IT processes (each warp serves every other sub-block; mean over CTAs and warps)")
for n, c in zip(names, sm.tolist()):
    print(f"  {n:20s} {c:8.1f}")
print(f"  {'total':20s} {sm.sum().item():8.1f}")
for w in (17, 18, 19):
    mm = p[:, w, :7].mean(dim=0) / n_sub
    print("issuer warp %d (17: scores of both tiles; 18/19: PV + row sums) per sub-block: wait k_full %.1f  v_full %.1f  p_full %.1f  s_free %.1f | issue S %.1f  issue PV+L %.1f | loop total %.1f" % ((w,) + tuple(mm.tolist())))
# spread between the softmax warps of a CTA (the slowest of the 8 warps of a tile gates p_full / s_free)
tot = p[:, FIRST:LAST, :5].sum(dim=2) / per_warp
print("softmax loop cycles per sub-block by warp (mean over CTAs):", " ".join(f"{v:.0f}" for v in tot.mean(dim=0).tolist()))
ex = p[:, FIRST:LAST, 3] / per_warp
print("exp phase by warp:", " ".join(f"{v:.0f}" for v in ex.mean(dim=0).tolist()))
