"""Host-side cost of one sampler step: the TINY network has negligible GPU work, so its step time is the Python +
ctypes + launch overhead that the full-shape step must hide (1 GPU) or pays (8 GPUs, short kernels)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dataclasses
import torch
from landiff_b200 import _C, dit
from landiff_b200.factory import TINY, build_warp, random_init_
from landiff_b200.sampling import VPSDEDPMPP2MSampler

cfg = dataclasses.replace(TINY, main_layers=30, control_layers=15)   # full depth, tiny width
warp = build_warp(cfg, device="cuda")
random_init_(warp, seed=0)
x = torch.randn(1, cfg.latent_t, 16, cfg.latent_h, cfg.latent_w, device="cuda")
ctx = torch.randn(1, cfg.text_length, cfg.text_hidden, device="cuda").bfloat16()
dit.InferValueRegistry.clear()
dit.InferValueRegistry.register("semantic_feature", torch.zeros_like(x).bfloat16())
cond, uc = {"crossattn": ctx}, {"crossattn": torch.zeros_like(ctx)}
s = VPSDEDPMPP2MSampler(num_steps=50, device="cuda")
s.sample(warp, x, cond, uc, max_steps=5)
torch.cuda.synchronize()
l0 = _C.LAUNCHES[0]
t0 = time.perf_counter()
s.sample(warp, x, cond, uc, max_steps=20)
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t = time.perf_counter() - t0
n = (_C.LAUNCHES[0] - l0) / 20
print(f"host overhead per sampler step (30+15 layers, tiny width): issue {t_issue / 20 * 1e3:.2f} ms, wall {t / 20 * 1e3:.2f} ms, "
      f"{n:.0f} C-ABI launches/step -> {t_issue / 20 / n * 1e6:.1f} us per launch")
