"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/trajectory50_config1.pt: the 50-step DPM++(2M) SDE CFG
trajectory of the CPU oracle (fp32, plain PyTorch restatement of the reference graph + sampler, oracle/dit_oracle.py)
at BASELINE config 1 (5 frames 240x352 -> latent 2x16x30x44, N = 886 tokens, full 15 + 30 layers, d = 1920) on seeded
weights (O.random_state_dict seeds 10 / 11 — regenerated identically by the GPU test) and a seeded CPU noise stream.
The GPU test (tests/test_network_gpu.py::test_50_step_trajectory_psnr) replays the same noise through the CUDA
path and requires PSNR >= 35 dB on the final latent (BASELINE.json north_star).

Run (about 10-15 minutes on 8 cores):  python -m oracle.make_trajectory_golden
"""
from __future__ import annotations

import time
from pathlib import Path

import torch

from . import dit_oracle as O

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "trajectory50_config1.pt"
SEED_INPUT, SEED_NOISE = 1, 42


def inputs(cfg):
    g = torch.Generator().manual_seed(SEED_INPUT)
    x = torch.randn(1, cfg.latent_t, 16, cfg.latent_h, cfg.latent_w, generator=g)
    ctx = (torch.randn(1, cfg.text_length, cfg.text_hidden, generator=g) * 0.2).bfloat16().float()
    sem = (torch.randn(1, cfg.latent_t, 16, cfg.latent_h, cfg.latent_w, generator=g) * 0.1).bfloat16().float()
    return x, ctx, sem


def main():
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    cfg = O.CONFIG1
    sdc = O.cast_state_dict(O.random_state_dict(cfg, True, seed=10), torch.float32)
    sdm = O.cast_state_dict(O.random_state_dict(cfg, False, seed=11), torch.float32)
    x, ctx, sem = inputs(cfg)

    def network(x2, t2, ctx2):
        return O.warp_forward(sdc, sdm, cfg, x2, t2, ctx2, sem)

    sampler = O.OracleSampler(num_steps=50)
    gen = torch.Generator().manual_seed(SEED_NOISE)
    trace = []
    t0 = time.time()
    out = sampler(network, x, ctx, torch.zeros_like(ctx), gen, trace=trace)
    print(f"50 oracle steps in {time.time() - t0:.0f} s; final latent std {out.std():.4f}")
    torch.save({"final": out.contiguous(), "steps": {i: trace[i].contiguous() for i in (0, 9, 24, 39, 48)},
                "seed_input": SEED_INPUT, "seed_noise": SEED_NOISE, "weight_seeds": (10, 11),
                "shape": "CONFIG1 (latent 2x16x30x44, N=886, 15+30 layers)"}, OUT)
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
