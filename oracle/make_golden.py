"""TEST INFRASTRUCTURE ONLY.  Generates the committed golden vectors under tests/golden/ by running the REFERENCE's
own code (imported unchanged from /root/reference on top of oracle/sat_shim.py) — only possible in the build
container.  Run:  python -m oracle.make_golden

  tests/golden/tiny_warp.pt     seeded (strong init, see ref_build.seeded_init_) TINY ControlDiffWarp pair: bf16-representable state dicts under the reference's
                                key names, inputs, reference fp32 output + per-layer control hidden states
  tests/golden/small_warp_b.pt  a second shape that exercises what TINY cannot: 3 control layers feeding a 5-layer main net
                                (zero-linear chaining, control add only for i < control_layers), 3 heads, 3 latent frames,
                                a text length that is not a multiple of anything, the SURVEY 8d "weak" init (N(0, 0.02^2))
                                and a mid-range timestep; weights regenerated from the seed (not stored), outputs only
  tests/golden/schedule.json    ZeroSNRDDPMDiscretization(shift_scale=3) 50-step table + timesteps, the 1000-entry
                                denoiser table (head/tail + checksum), DynamicCFG scales as the sampler calls it,
                                DPM++(2M) SDE scalars for every step
  tests/golden/sampler_toy.pt   5-step trajectory of the reference VPSDEDPMPP2MSampler + DiscreteDenoiser + DynamicCFG
                                driving a toy analytic network (pins the sampler/denoiser/guider algebra + RNG order),
                                without and with fixed_frames=1 (the streaming prefix hook)
"""
from __future__ import annotations

import json
from pathlib import Path

import torch

from . import ref_build as rb
from . import sat_shim

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def reference_sampler(num_steps=50, device="cpu", fixed_frames=0):
    sat_shim.install()
    from landiff.diffusion.sgm.modules.diffusionmodules.denoiser import DiscreteDenoiser
    from landiff.diffusion.sgm.modules.diffusionmodules.sampling import VPSDEDPMPP2MSampler

    disc = {"target": "landiff.diffusion.sgm.modules.diffusionmodules.discretizer.ZeroSNRDDPMDiscretization",
            "params": {"shift_scale": 3.0}}
    sampler = VPSDEDPMPP2MSampler(
        num_steps=num_steps, verbose=False, device=device, discretization_config=disc, fixed_frames=fixed_frames,
        guider_config={"target": "landiff.diffusion.sgm.modules.diffusionmodules.guiders.DynamicCFG",
                       "params": {"scale": 6, "exp": 5, "num_steps": num_steps}})
    denoiser = DiscreteDenoiser(
        weighting_config={"target": "landiff.diffusion.sgm.modules.diffusionmodules.denoiser_weighting.EpsWeighting"},
        scaling_config={"target": "landiff.diffusion.sgm.modules.diffusionmodules.denoiser_scaling.VideoScaling"},
        num_idx=1000, discretization_config=disc, quantize_c_noise=False)
    return sampler, denoiser


def toy_network(x, t, cond, **kw):
    """analytic stand-in with batch-row, timestep and conditioning dependence (bf16 output like the real net)"""
    c = cond["crossattn"].mean(dim=(1, 2)).view(-1, 1, 1, 1, 1)
    return (torch.tanh(x * 0.5 + c) * (1.0 + t.view(-1, 1, 1, 1, 1) / 1000.0)).to(torch.bfloat16)


def make_tiny_warp():
    cfg = rb.TINY
    ctrl, main = rb.build_reference(cfg, seed=0, strong=True)
    # bf16-representable parameters so the CUDA bf16 path and the fp32 references see identical weights
    for m in (ctrl, main):
        for p in m.parameters():
            p.data.copy_(p.data.to(torch.bfloat16).float())
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, cfg.latent_t, 16, cfg.latent_h, cfg.latent_w, generator=g)
    ctx = (torch.randn(2, cfg.text_length, cfg.text_hidden, generator=g) * 0.2).to(torch.bfloat16).float()
    ctx[0] = 0  # uncond row is all-zero (force_uc_zero_embeddings)
    sem = (torch.randn(1, cfg.latent_t, 16, cfg.latent_h, cfg.latent_w, generator=g) * 0.1).to(torch.bfloat16).float()
    t = torch.tensor([999.0, 999.0])
    out, ctl = rb.reference_forward(ctrl, main, x, t, ctx, sem)
    t2 = torch.tensor([19.0, 19.0])
    out2, _ = rb.reference_forward(ctrl, main, x, t2, ctx, sem)
    blob = {
        "cfg": cfg.__dict__,
        "sd_ctrl": {k: v.to(torch.bfloat16) for k, v in ctrl.state_dict().items()},
        "sd_main": {k: v.to(torch.bfloat16) for k, v in main.state_dict().items()},
        "x": x, "context": ctx, "semantic_feature": sem, "t": t, "out": out.float(),
        "control_hidden": [c["hidden_states"].float() for c in ctl], "t2": t2, "out2": out2.float(),
    }
    torch.save(blob, OUT / "tiny_warp.pt")
    print("tiny_warp.pt", (OUT / "tiny_warp.pt").stat().st_size, "bytes; out abs mean", out.abs().mean().item())


SMALL_B = dict(hidden_size=192, num_heads=3, main_layers=5, control_layers=3, time_embed_dim=96, text_hidden=128,
               text_length=7, latent_t=3, latent_h=6, latent_w=10)


def small_b_inputs(cfg):
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, cfg.latent_t, 16, cfg.latent_h, cfg.latent_w, generator=g)
    ctx = (torch.randn(2, cfg.text_length, cfg.text_hidden, generator=g) * 0.2).to(torch.bfloat16).float()
    ctx[0] = 0
    sem = (torch.randn(1, cfg.latent_t, 16, cfg.latent_h, cfg.latent_w, generator=g) * 0.1).to(torch.bfloat16).float()
    return x, ctx, sem, torch.tensor([519.0, 519.0])


def make_small_warp_b():
    """Outputs of the reference modules for SMALL_B.  Weights are NOT stored: `ref_build.seeded_init_` draws them in
    sorted-parameter-name order from one seeded generator, which `dit_oracle.seeded_state_dict` restates, so the test
    regenerates bit-identical (bf16-rounded) parameters from the seed."""
    cfg = rb.DiTConfig(**SMALL_B)
    blob = {"cfg": SMALL_B, "seed": 5}
    for strong, tag in ((False, "weak"), (True, "strong")):
        ctrl, main = rb.build_reference(cfg, seed=5, strong=strong)
        for m in (ctrl, main):
            for p in m.parameters():
                p.data.copy_(p.data.to(torch.bfloat16).float())
        x, ctx, sem, t = small_b_inputs(cfg)
        out, ctl = rb.reference_forward(ctrl, main, x, t, ctx, sem)
        blob[tag] = {"out": out.float(), "control_hidden": [c["hidden_states"].float() for c in ctl],
                     "probe": {k: ctrl.state_dict()[k].flatten()[:4].clone() for k in
                               ("time_embed.0.weight", "mixins.adaln_layer.zero_linears.2.weight",
                                "transformer.layers.1.attention.query_key_value.bias")}}
        print(f"small_warp_b[{tag}] out abs mean", out.abs().mean().item())
    torch.save(blob, OUT / "small_warp_b.pt")
    print("small_warp_b.pt", (OUT / "small_warp_b.pt").stat().st_size, "bytes")


def make_schedule():
    sampler, denoiser = reference_sampler(50)
    x = torch.zeros(1, 1, 1, 1, 1)
    _, s_in, acs, num_sigmas, _, _, timesteps = sampler.prepare_sampling_loop(x, {}, None)
    table = denoiser.sigmas
    cfg = {int(t): sampler.guider.scale_schedule(None, 50 - int(t)) for t in timesteps[1:]}
    scal = []
    for i in range(num_sigmas - 1):
        prev = None if i == 0 else s_in * acs[i - 1]
        a, nxt = s_in * acs[i], s_in * acs[i + 1]
        h, r, _, _ = sampler.get_variables(a, nxt, prev)
        mult = sampler.get_mult(h, r, a, nxt, prev)
        mn = (1 - nxt ** 2) ** 0.5 * (1 - (-2 * h).exp()) ** 0.5
        q = denoiser.possibly_quantize_sigma(a)
        scal.append({"i": i, "mult": [float(m) for m in mult], "mult_noise": float(mn), "a_quantized": float(q)})
    blob = {
        "alphas_cumprod_sqrt": [float(v) for v in acs], "timesteps": [int(v) for v in timesteps],
        "table_head": [float(v) for v in table[:8]], "table_tail": [float(v) for v in table[-8:]],
        "table_sum": float(table.double().sum()), "table_len": int(table.numel()),
        "cfg_scale_by_timestep": cfg, "steps": scal,
    }
    (OUT / "schedule.json").write_text(json.dumps(blob, indent=1))
    print("schedule.json written;", blob["alphas_cumprod_sqrt"][:3], blob["timesteps"][:4])


def make_sampler_toy():
    sampler, denoiser = reference_sampler(5)
    g = torch.Generator().manual_seed(7)
    x0 = torch.randn(1, 2, 4, 6, 8, generator=g)
    cond = {"crossattn": torch.randn(1, 3, 8, generator=g)}
    uc = {"crossattn": torch.zeros(1, 3, 8)}
    torch.manual_seed(42)  # the sampler draws from the global generator (randn_like)
    den = lambda inp, sigma, c, **kw: denoiser(toy_network, inp, sigma, c, **kw)
    out = sampler(den, x0.clone(), cond, uc=uc)
    # the same loop with a fixed 1-frame prefix (the streaming hook, sampling.py:800-817, 834-835)
    sampler_ff, _ = reference_sampler(5, fixed_frames=1)
    torch.manual_seed(42)
    out_ff = sampler_ff(den, x0.clone(), cond, uc=uc)
    torch.save({"x0": x0, "cond": cond["crossattn"], "uc": uc["crossattn"], "seed": 42, "num_steps": 5, "out": out,
                "fixed_frames": 1, "out_fixed_frames": out_ff}, OUT / "sampler_toy.pt")
    print("sampler_toy.pt out abs mean", out.abs().mean().item(), "with prefix", out_ff.abs().mean().item())


if __name__ == "__main__":
    OUT.mkdir(parents=True, exist_ok=True)
    make_tiny_warp()
    make_small_warp_b()
    make_schedule()
    make_sampler_toy()
