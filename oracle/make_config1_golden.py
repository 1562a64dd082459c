"""TEST INFRASTRUCTURE ONLY.  Goldens at BASELINE config 1 (5 frames 240x352 -> latent 2x16x30x44, N = 886 tokens,
d = 1920, 30 heads) produced by the REFERENCE's own modules (dit_video_concat.py imported unchanged from
/root/reference on top of oracle/sat_shim.py) — the real width and depth, not a toy shape.  Only runnable in the build
container.  Run:  python -m oracle.make_config1_golden [net] [traj]

  tests/golden/config1_ref.pt              one CFG evaluation of ControlDiffusionTransformer -> DiffusionTransformer:
                                           "weak":   full 15 + 30 layers, N(0, 0.02^2) init (O.random_state_dict 10/11)
                                           "strong": 2 + 4 layers, O(1) modulations (strong init)
                                           Weights are regenerated from the seeds by the tests (not stored).
  tests/golden/trajectory50_config1_ref.pt the 50-step DPM++(2M) SDE CFG trajectory produced by the reference
                                           VPSDEDPMPP2MSampler + DiscreteDenoiser + DynamicCFG objects driving the
                                           reference network exactly like SATVideoDiffusionEngine.sample
                                           (diffusion_video.py:302-313), fp32 on CPU, global RNG seeded 42.
"""
from __future__ import annotations

import dataclasses
import sys
import time
from pathlib import Path

import torch

from . import dit_oracle as O
from . import ref_build as rb
from . import sat_shim
from .make_golden import reference_sampler
from .make_trajectory_golden import SEED_NOISE, inputs

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def build_loaded(cfg_o: O.OracleConfig, strong: bool):
    """Reference modules at config-1 shape carrying O.random_state_dict(seed 10 / 11) (bf16-representable values)."""
    cfg_r = rb.DiTConfig(**{k: getattr(cfg_o, k) for k in ("hidden_size", "num_heads", "main_layers", "control_layers",
                                                          "time_embed_dim", "text_hidden", "text_length", "latent_t",
                                                          "latent_h", "latent_w", "in_channels", "interp")})
    ctrl, main = rb.build_reference(cfg_r, seed=0)
    sdc = O.cast_state_dict(O.random_state_dict(cfg_o, True, seed=10, strong=strong), torch.float32)
    sdm = O.cast_state_dict(O.random_state_dict(cfg_o, False, seed=11, strong=strong), torch.float32)
    ctrl.load_state_dict(sdc, strict=True)
    main.load_state_dict(sdm, strict=True)
    return ctrl, main


def net_inputs(cfg_o):
    """Same inputs as tests/test_network_gpu.py::test_config1_*."""
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, cfg_o.latent_t, 16, cfg_o.latent_h, cfg_o.latent_w, generator=g)
    ctx = (torch.randn(2, cfg_o.text_length, cfg_o.text_hidden, generator=g) * 0.2).bfloat16().float()
    ctx[0] = 0
    sem = (torch.randn(1, cfg_o.latent_t, 16, cfg_o.latent_h, cfg_o.latent_w, generator=g) * 0.1).bfloat16().float()
    return x, ctx, sem, torch.tensor([519.0, 519.0])


def make_net():
    blob = {"shape": "CONFIG1 (latent 2x16x30x44, N=886, d=1920)", "weight_seeds": (10, 11)}
    for tag, strong, layers in (("weak", False, None), ("strong", True, (4, 2))):
        cfg_o = O.CONFIG1 if layers is None else dataclasses.replace(O.CONFIG1, main_layers=layers[0], control_layers=layers[1])
        ctrl, main = build_loaded(cfg_o, strong)
        x, ctx, sem, t = net_inputs(cfg_o)
        t0 = time.time()
        out, ctl = rb.reference_forward(ctrl, main, x, t, ctx, sem)
        print(f"config1_ref[{tag}]: reference forward {time.time() - t0:.1f} s, out abs mean {out.abs().mean():.4f}")
        blob[tag] = {"out": out.float().contiguous(), "main_layers": cfg_o.main_layers, "control_layers": cfg_o.control_layers,
                     "control_last": ctl[-1]["hidden_states"].float()[:, ::37].contiguous()}
    torch.save(blob, OUT / "config1_ref.pt")
    print("config1_ref.pt", (OUT / "config1_ref.pt").stat().st_size, "bytes")


def make_traj():
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    sat_shim.install()
    from landiff.diffusion.sgm.util import InferValueRegistry

    cfg_o = O.CONFIG1
    ctrl, main = build_loaded(cfg_o, False)
    x, ctx, sem = inputs(cfg_o)
    sampler, denoiser = reference_sampler(50)

    def warp(xx, t, c, **kw):     # ControlDiffWarp.forward over OpenAIWrapper (:1196-1200, wrappers.py:31-51)
        cl = ctrl(xx, timesteps=t, context=c["crossattn"], **kw)
        return main(xx, timesteps=t, context=c["crossattn"], control_layers_output=cl, **kw)

    den = lambda inp, sigma, c, **kw: denoiser(warp, inp, sigma, c, concat_images=None, **kw)
    InferValueRegistry.clear()
    InferValueRegistry.register("semantic_feature", sem)
    trace = {}
    step_fn = sampler.sampler_step
    counter = [0]

    def traced_step(*a, **k):      # record x after chosen steps without touching the reference loop
        r = step_fn(*a, **k)
        if counter[0] in (0, 9, 24, 39, 48):
            trace[counter[0]] = r[0].clone().contiguous()
        counter[0] += 1
        return r

    sampler.sampler_step = traced_step
    torch.manual_seed(SEED_NOISE)   # the reference sampler draws from the global generator (randn_like)
    t0 = time.time()
    with torch.no_grad():
        out = sampler(den, x.clone(), {"crossattn": ctx}, uc={"crossattn": torch.zeros_like(ctx)})
    InferValueRegistry.clear()
    print(f"50 reference steps in {time.time() - t0:.0f} s; final latent std {out.std():.4f}")
    torch.save({"final": out.float().contiguous(), "steps": trace, "seed_input": 1, "seed_noise": SEED_NOISE,
                "weight_seeds": (10, 11), "shape": "CONFIG1 (latent 2x16x30x44, N=886, 15+30 layers)",
                "made_by": "reference VPSDEDPMPP2MSampler + DiscreteDenoiser + DynamicCFG + reference DiT modules, fp32 CPU"},
               OUT / "trajectory50_config1_ref.pt")
    print("trajectory50_config1_ref.pt", (OUT / "trajectory50_config1_ref.pt").stat().st_size, "bytes")


if __name__ == "__main__":
    which = sys.argv[1:] or ["net", "traj"]
    if "net" in which:
        make_net()
    if "traj" in which:
        make_traj()
