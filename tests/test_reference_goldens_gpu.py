"""GPU parity against goldens produced by the REFERENCE's own modules at the real width/depth (BASELINE config 1:
d = 1920, 30 heads, 15 + 30 layers, N = 886 tokens; oracle/make_config1_golden.py), and of the reference-compatible
sampler entry point `VPSDEDPMPP2MSampler.__call__` driven exactly like SATVideoDiffusionEngine.sample does
(diffusion_video.py:302-313) against the reference-generated toy trajectory.

Tolerances (BASELINE.json north_star): per-step rel-L2 <= 1e-2, cosine >= 0.999; 50-step latent PSNR >= 35 dB.
"""
import dataclasses

import pytest
import torch

from conftest import GOLDEN
from landiff_b200 import dit
from landiff_b200 import sampling as S
from landiff_b200.factory import CONFIG1, build_warp
from oracle import dit_oracle as O
from oracle.make_golden import toy_network

pytestmark = pytest.mark.gpu
REL_TOL, COS_TOL = 1e-2, 0.999


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm()).item()


def cos(a, b):
    a, b = a.double().cpu().flatten(), b.double().cpu().flatten()
    return (a @ b / (a.norm() * b.norm())).item()


def _net_inputs(cfg_o):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, cfg_o.latent_t, 16, cfg_o.latent_h, cfg_o.latent_w, generator=g)
    ctx = (torch.randn(2, cfg_o.text_length, cfg_o.text_hidden, generator=g) * 0.2).bfloat16().float()
    ctx[0] = 0
    sem = (torch.randn(1, cfg_o.latent_t, 16, cfg_o.latent_h, cfg_o.latent_w, generator=g) * 0.1).bfloat16().float()
    return x, ctx, sem, torch.tensor([519.0, 519.0])


@pytest.mark.parametrize("tag,strong", [("weak", False), ("strong", True)])
def test_config1_against_reference_modules(tag, strong):
    """d = 1920 / N = 886 output of the reference ControlDiffusionTransformer -> DiffusionTransformer (fp32, CPU) on
    the seeded weights vs the CUDA path.  Also reports the distance of the oracle graph run in eager bf16 on this GPU
    (cuBLAS + SDPA: the reference's own rounding points) to the same golden, next to ours (SURVEY hard part 5)."""
    gold = torch.load(GOLDEN / "config1_ref.pt", weights_only=False)[tag]
    cfg_o = dataclasses.replace(O.CONFIG1, main_layers=gold["main_layers"], control_layers=gold["control_layers"])
    cfg_p = dataclasses.replace(CONFIG1, main_layers=gold["main_layers"], control_layers=gold["control_layers"])
    sdc = O.random_state_dict(cfg_o, True, seed=10, strong=strong)
    sdm = O.random_state_dict(cfg_o, False, seed=11, strong=strong)
    x, ctx, sem, t = _net_inputs(cfg_o)
    warp = build_warp(cfg_p, device="cuda", sd_ctrl=sdc, sd_main=sdm)
    dit.InferValueRegistry.clear()
    dit.InferValueRegistry.register("semantic_feature", sem.cuda())
    ctl = warp.control_model(x.cuda(), t.cuda(), {"crossattn": ctx.cuda()}, idx=t.cuda())
    ctl_last = ctl[-1]["hidden_states"].float().cpu()[:, ::37].clone()
    out = warp(x.cuda(), t.cuda(), {"crossattn": ctx.cuda()}, idx=t.cuda()).float().cpu()
    torch.cuda.synchronize()
    dit.InferValueRegistry.clear()
    r, c = rel(out, gold["out"]), cos(out, gold["out"])
    rc = rel(ctl_last, gold["control_last"])
    # the oracle graph in eager bf16 on this GPU against the same reference-generated golden
    dev_sd = lambda sd: {k: v.cuda() for k, v in sd.items()}
    eager = O.warp_forward(dev_sd(sdc), dev_sd(sdm), cfg_o, x.cuda(), t.cuda(), ctx.cuda().bfloat16(), sem.cuda().bfloat16())
    re_, rd = rel(eager.float().cpu(), gold["out"]), rel(out, eager.float().cpu())
    print(f"config1[{tag}] vs reference modules: ours rel-L2 {r:.3e} cos {c:.6f} (last control layer {rc:.3e}); "
          f"eager-bf16 oracle rel-L2 {re_:.3e}; ours vs eager-bf16 {rd:.3e}")
    assert r <= REL_TOL and c >= COS_TOL, f"rel-L2 {r:.3e} cos {c:.6f}"
    # the 15th control output is an INTERMEDIATE carried in bf16 from layer to layer like the reference does (the gate of
    # north_star is on the predicted latent above); measured 1.2e-2 on the weak init, where eager bf16 itself is 1.6e-2 off
    assert rc <= 2.5 * REL_TOL, f"last control layer rel-L2 {rc:.3e}"
    assert rd <= 2.5 * REL_TOL, f"CUDA path vs eager-bf16 oracle rel-L2 {rd:.3e}"


def test_50_step_trajectory_against_reference_objects():
    """The golden is the reference VPSDEDPMPP2MSampler + DiscreteDenoiser + DynamicCFG driving the reference network
    (fp32, CPU, global RNG seeded 42); the CUDA path replays the same CPU noise stream."""
    from oracle.make_trajectory_golden import SEED_NOISE, inputs

    gold = torch.load(GOLDEN / "trajectory50_config1_ref.pt", weights_only=False)
    cfg_o = O.CONFIG1
    sdc = O.random_state_dict(cfg_o, True, seed=10)
    sdm = O.random_state_dict(cfg_o, False, seed=11)
    x, ctx, sem = inputs(cfg_o)
    warp = build_warp(CONFIG1, device="cuda", sd_ctrl=sdc, sd_main=sdm)
    dit.InferValueRegistry.clear()
    dit.InferValueRegistry.register("semantic_feature", sem.cuda())
    gen = torch.Generator().manual_seed(SEED_NOISE)
    noise = lambda t: torch.randn(t.shape, generator=gen).to(t.device)
    sampler = S.VPSDEDPMPP2MSampler(num_steps=50, device="cuda")
    trace = {}
    cond = {"crossattn": ctx.cuda().bfloat16()}
    uc = {"crossattn": torch.zeros_like(cond["crossattn"])}
    out = sampler.sample(warp, x.cuda(), cond, uc, noise_fn=noise,
                         step_callback=lambda i, xs: trace.__setitem__(i, xs.float().cpu().clone()))
    torch.cuda.synchronize()
    dit.InferValueRegistry.clear()

    def psnr(a, b):
        mse = ((a.double() - b.double()) ** 2).mean()
        return float(10 * torch.log10(b.double().abs().max() ** 2 / mse))

    for i, ref in gold["steps"].items():
        p = psnr(trace[i], ref)
        assert p >= 35.0, f"step {i + 1}: PSNR {p:.1f} dB"
    p = psnr(out.float().cpu(), gold["final"])
    print(f"50-step latent vs the reference sampler/denoiser/guider/network objects: PSNR {p:.1f} dB, "
          f"rel-L2 {rel(out.float().cpu(), gold['final']):.3e}")
    assert p >= 35.0, f"final latent PSNR {p:.1f} dB < 35"


@pytest.mark.parametrize("fixed_frames,key", [(0, "out"), (1, "out_fixed_frames")])
def test_reference_compatible_sampler_call_entry(fixed_frames, key):
    """`sampler(denoiser_lambda, x, cond, uc=uc)` with `denoiser_lambda = lambda input, sigma, c, **kw:
    DiscreteDenoiser(network, input, sigma, c, **kw)` — the call the reference engine makes — on the GPU, against the
    trajectory the reference's own sampler + denoiser + guider produced for the toy network (sampler_toy.pt).  The global
    CUDA RNG cannot replay the CPU stream, so the noise draws are redirected to a CPU generator seeded like the golden."""
    g = torch.load(GOLDEN / "sampler_toy.pt", weights_only=False)
    ref_cfg = "landiff.diffusion.sgm.modules.diffusionmodules."
    disc = {"target": ref_cfg + "discretizer.ZeroSNRDDPMDiscretization", "params": {"shift_scale": 3.0}}
    sampler = S.VPSDEDPMPP2MSampler(num_steps=g["num_steps"], discretization_config=disc, fixed_frames=fixed_frames,
                                    guider_config={"target": ref_cfg + "guiders.DynamicCFG",
                                                   "params": {"scale": 6, "exp": 5, "num_steps": g["num_steps"]}})
    denoiser = S.DiscreteDenoiser(discretization_config=disc, quantize_c_noise=False,
                                  scaling_config={"target": ref_cfg + "denoiser_scaling.VideoScaling"})
    den = lambda inp, sigma, c, **kw: denoiser(toy_network, inp, sigma, c, **kw)
    cpu_gen = torch.Generator().manual_seed(g["seed"])
    orig = torch.randn_like
    torch.randn_like = lambda t, **k: torch.randn(t.shape, generator=cpu_gen).to(t.device)
    try:
        out = sampler(den, g["x0"].cuda(), {"crossattn": g["cond"].cuda()}, uc={"crossattn": g["uc"].cuda()})
    finally:
        torch.randn_like = orig
    torch.cuda.synchronize()
    r = rel(out.float().cpu(), g[key])
    assert out.is_cuda and out.shape == g[key].shape
    assert r <= 2e-3, f"__call__ entry (fixed_frames={fixed_frames}) rel-L2 {r:.3e} vs the reference trajectory"
