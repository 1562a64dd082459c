"""GPU parity of the whole drop-in network (through the plugin surface: ControlDiffWarp(x, t, cond) exactly as the
reference denoiser calls it, denoiser.py:38-41) against (a) the committed golden vectors produced by the reference's
own code and (b) the CPU oracle on seeded weights at BASELINE config 1.

Tolerances (BASELINE.json north_star): per-step rel-L2 <= 1e-2 and cosine >= 0.999 in bf16.
"""
import pytest
import torch

from conftest import GOLDEN
from landiff_b200 import dit
from landiff_b200.factory import CONFIG1, TINY, build_warp
from oracle import dit_oracle as O

pytestmark = pytest.mark.gpu

REL_TOL = 1e-2
COS_TOL = 0.999


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm()).item()


def cos(a, b):
    a, b = a.double().cpu().flatten(), b.double().cpu().flatten()
    return (a @ b / (a.norm() * b.norm())).item()


def run_warp(warp, x, t, ctx, sem):
    dit.InferValueRegistry.clear()
    dit.InferValueRegistry.register("semantic_feature", sem.cuda())
    out = warp(x.cuda(), t.cuda(), {"crossattn": ctx.cuda()}, idx=t.cuda())
    torch.cuda.synchronize()
    dit.InferValueRegistry.clear()
    return out


@pytest.fixture(scope="module")
def tiny():
    return torch.load(GOLDEN / "tiny_warp.pt", weights_only=False)


def test_tiny_golden_from_reference_code(tiny):
    warp = build_warp(TINY, device="cuda", sd_ctrl=tiny["sd_ctrl"], sd_main=tiny["sd_main"])
    out = run_warp(warp, tiny["x"], tiny["t"], tiny["context"], tiny["semantic_feature"]).float().clone()
    assert out.shape == tiny["out"].shape
    r, c = rel(out, tiny["out"]), cos(out, tiny["out"])
    assert r <= REL_TOL and c >= COS_TOL, f"rel-L2 {r:.3e} cos {c:.6f}"
    out2 = run_warp(warp, tiny["x"], tiny["t2"], tiny["context"], tiny["semantic_feature"]).float().clone()
    r2 = rel(out2, tiny["out2"])
    assert r2 <= REL_TOL, f"second timestep rel-L2 {r2:.3e}"
    assert rel(out2, tiny["out"]) > 5 * REL_TOL, "timestep must matter in the strong-init golden"


def test_tiny_control_hidden_states(tiny):
    warp = build_warp(TINY, device="cuda", sd_ctrl=tiny["sd_ctrl"], sd_main=tiny["sd_main"])
    dit.InferValueRegistry.clear()
    dit.InferValueRegistry.register("semantic_feature", tiny["semantic_feature"].cuda())
    ctl = warp.control_model(tiny["x"].cuda(), tiny["t"].cuda(), {"crossattn": tiny["context"].cuda()})
    torch.cuda.synchronize()
    dit.InferValueRegistry.clear()
    assert len(ctl) == len(tiny["control_hidden"])
    for i, (a, b) in enumerate(zip(ctl, tiny["control_hidden"])):
        r = rel(a["hidden_states"].float(), b)
        assert r <= REL_TOL, f"control layer {i}: rel-L2 {r:.3e}"


def test_tiny_batch_rows_are_independent(tiny):
    """CFG-parallel premise: nothing mixes batch rows — running the cond row alone gives the same row."""
    warp = build_warp(TINY, device="cuda", sd_ctrl=tiny["sd_ctrl"], sd_main=tiny["sd_main"])
    both = run_warp(warp, tiny["x"], tiny["t"], tiny["context"], tiny["semantic_feature"]).float().clone()
    one = run_warp(warp, tiny["x"][1:], tiny["t"][1:], tiny["context"][1:], tiny["semantic_feature"]).float().clone()
    assert rel(one[0], both[1]) < 2e-3


def test_missing_semantic_feature_raises(tiny):
    warp = build_warp(TINY, device="cuda", sd_ctrl=tiny["sd_ctrl"], sd_main=tiny["sd_main"])
    dit.InferValueRegistry.clear()
    with pytest.raises(RuntimeError, match="semantic_feature"):
        warp(tiny["x"].cuda(), tiny["t"].cuda(), {"crossattn": tiny["context"].cuda()})


@pytest.mark.parametrize("strong", [False, True])
def test_config1_against_cpu_oracle(strong):
    """BASELINE config 1 (5 frames, 240x352 -> N = 886 tokens, full 15+30 layers, d=1920): CUDA bf16 vs the fp32 CPU
    oracle on identical seeded (bf16-representable) weights.  `strong` shrinks the depth to keep O(1) modulations
    inside the bf16 tolerance."""
    cfg_o, cfg_p = O.CONFIG1, CONFIG1
    if strong:
        import dataclasses

        cfg_o = dataclasses.replace(cfg_o, main_layers=4, control_layers=2)
        cfg_p = dataclasses.replace(cfg_p, main_layers=4, control_layers=2)
    sdc = O.random_state_dict(cfg_o, True, seed=10, strong=strong)
    sdm = O.random_state_dict(cfg_o, False, seed=11, strong=strong)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, cfg_o.latent_t, 16, cfg_o.latent_h, cfg_o.latent_w, generator=g)
    ctx = (torch.randn(2, cfg_o.text_length, cfg_o.text_hidden, generator=g) * 0.2).bfloat16().float()
    ctx[0] = 0
    sem = (torch.randn(1, cfg_o.latent_t, 16, cfg_o.latent_h, cfg_o.latent_w, generator=g) * 0.1).bfloat16().float()
    t = torch.tensor([519.0, 519.0])
    warp = build_warp(cfg_p, device="cuda", sd_ctrl=sdc, sd_main=sdm)
    out = run_warp(warp, x, t, ctx, sem).float().cpu()
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    ref = O.warp_forward(O.cast_state_dict(sdc, torch.float32), O.cast_state_dict(sdm, torch.float32), cfg_o, x, t, ctx, sem)
    r, c = rel(out, ref), cos(out, ref)
    assert r <= REL_TOL and c >= COS_TOL, f"rel-L2 {r:.3e} cos {c:.6f}"


def test_50_step_trajectory_psnr():
    """BASELINE.json north_star: the 50-step end-to-end latent must match the reference GPU-free fp32 path within
    PSNR >= 35 dB.  Golden = the CPU oracle's 50-step DPM++(2M) SDE CFG trajectory at config 1 (full depth) on the
    seeded weights, generated by oracle/make_trajectory_golden.py; here the same noise stream is replayed through
    the CUDA path (drop-in network + fused sampler update)."""
    from landiff_b200.sampling import VPSDEDPMPP2MSampler
    from oracle.make_trajectory_golden import SEED_NOISE, inputs

    gold = torch.load(GOLDEN / "trajectory50_config1.pt", weights_only=False)
    cfg_o = O.CONFIG1
    sdc = O.random_state_dict(cfg_o, True, seed=10)
    sdm = O.random_state_dict(cfg_o, False, seed=11)
    x, ctx, sem = inputs(cfg_o)
    warp = build_warp(CONFIG1, device="cuda", sd_ctrl=sdc, sd_main=sdm)
    dit.InferValueRegistry.clear()
    dit.InferValueRegistry.register("semantic_feature", sem.cuda())
    gen = torch.Generator().manual_seed(SEED_NOISE)
    noise = lambda t: torch.randn(t.shape, generator=gen).to(t.device)
    sampler = VPSDEDPMPP2MSampler(num_steps=50, device="cuda")
    trace = {}
    cond = {"crossattn": ctx.cuda().bfloat16()}
    uc = {"crossattn": torch.zeros_like(cond["crossattn"])}
    out = sampler.sample(warp, x.cuda(), cond, uc, noise_fn=noise,
                         step_callback=lambda i, xs: trace.__setitem__(i, xs.float().cpu().clone()))
    torch.cuda.synchronize()
    dit.InferValueRegistry.clear()

    def psnr(a, b):
        mse = ((a.double() - b.double()) ** 2).mean()
        peak = b.double().abs().max()
        return float(10 * torch.log10(peak ** 2 / mse))

    for i, ref in gold["steps"].items():
        p = psnr(trace[i], ref)
        assert p >= 35.0, f"step {i + 1}: PSNR {p:.1f} dB"
    p = psnr(out.float().cpu(), gold["final"])
    print(f"50-step final latent PSNR {p:.1f} dB, rel-L2 {rel(out.float().cpu(), gold['final']):.3e}")
    assert p >= 35.0, f"final latent PSNR {p:.1f} dB < 35"


def test_full_shape_step_against_oracle_fp32_on_gpu():
    """BASELINE config 2 shape (49 frames 480x720: latent 13x16x60x90, N = 17 776 tokens, 15 + 30 layers, CFG batch 2):
    one network evaluation of the CUDA path against the oracle's plain-PyTorch fp32 graph executed on the same GPU
    (TF32 off; the oracle's SDPA call is served by an exact fp32 softmax(QK^T/8)V evaluated in 2048-query chunks so
    that the 17 776^2 score matrix is never materialised) on identical device-generated random-init weights.
    Gate (north_star): rel-L2 <= 1e-2, cosine >= 0.999."""
    from landiff_b200.factory import FULL, random_init_

    def chunked_sdpa(q, k, v):
        outs = []
        kt = k.transpose(-1, -2)
        for s0 in range(0, q.shape[2], 2048):
            p = torch.softmax((q[:, :, s0:s0 + 2048] @ kt) * 0.125, dim=-1)
            outs.append(p @ v)
        return torch.cat(outs, dim=2)

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    warp = build_warp(FULL, device="cuda")
    random_init_(warp, seed=0)
    sd = warp.state_dict()
    pick = lambda prefix: {k[len(prefix):]: v.float() for k, v in sd.items() if k.startswith(prefix)}
    sdc, sdm = pick("control_model.diffusion_model."), pick("main_model.diffusion_model.")
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 13, 16, 60, 90, generator=g).cuda()
    ctx = (torch.randn(1, 226, 4096, generator=g) * 0.2).bfloat16().float().cuda()
    sem = (torch.randn(1, 13, 16, 60, 90, generator=g) * 0.1).bfloat16().float().cuda()
    x2, ctx2 = torch.cat([x, x]), torch.cat([torch.zeros_like(ctx), ctx])
    t = torch.tensor([519.0, 519.0], device="cuda")
    out = run_warp(warp, x2, t, ctx2, sem).float()
    del warp
    torch.cuda.empty_cache()
    orig = O.F.scaled_dot_product_attention
    O.F.scaled_dot_product_attention = chunked_sdpa
    try:
        ref = O.warp_forward(sdc, sdm, O.FULL, x2, t, ctx2, sem)
    finally:
        O.F.scaled_dot_product_attention = orig
    torch.cuda.synchronize()
    r, c = rel(out, ref), cos(out, ref)
    print(f"full shape (N=17776, 15+30 layers, B=2): rel-L2 {r:.3e} cos {c:.6f}; uncond/cond rows differ by {rel(out[0], out[1]):.3e}")
    assert r <= REL_TOL and c >= COS_TOL, f"rel-L2 {r:.3e} cos {c:.6f}"


@pytest.mark.parametrize("tag,strong", [("weak", False), ("strong", True)])
def test_small_warp_b_golden_from_reference_code(tag, strong):
    """Second reference-generated golden (tests/golden/small_warp_b.pt): 3 control layers into a 5-layer main net,
    d = 192 (3 heads), 3 latent frames, text length 7, timestep 519 — weights regenerated from the stored seed."""
    from landiff_b200.factory import DiTShape
    from oracle.make_golden import small_b_inputs

    g = torch.load(GOLDEN / "small_warp_b.pt", weights_only=False)
    cfg_o = O.OracleConfig(**g["cfg"])
    sdc = O.seeded_state_dict(cfg_o, True, g["seed"], strong)
    sdm = O.seeded_state_dict(cfg_o, False, g["seed"] + 1, strong)
    warp = build_warp(DiTShape(**g["cfg"]), device="cuda", sd_ctrl=sdc, sd_main=sdm)
    x, ctx, sem, t = small_b_inputs(cfg_o)
    out = run_warp(warp, x, t, ctx, sem).float().cpu()
    r, c = rel(out, g[tag]["out"]), cos(out, g[tag]["out"])
    assert r <= REL_TOL and c >= COS_TOL, f"rel-L2 {r:.3e} cos {c:.6f}"


def test_cuda_graph_replay_is_bit_identical_to_eager(tiny):
    """the whole ControlDiffWarp forward captured once into a CUDA graph and replayed per step.  Replays
    with new latents / timesteps / text features / semantic features must equal the eager launches bit for bit (same
    kernels, same order), and a 4-step sampler trajectory through the graphed network must equal the eager one."""
    from landiff_b200.graph import GraphedWarp
    from landiff_b200.sampling import VPSDEDPMPP2MSampler

    warp = build_warp(TINY, device="cuda", sd_ctrl=tiny["sd_ctrl"], sd_main=tiny["sd_main"])
    gw = GraphedWarp(warp)
    g = torch.Generator().manual_seed(3)
    for k in range(3):
        x = (tiny["x"] + 0.3 * k * torch.randn(tiny["x"].shape, generator=g)).cuda()
        t = torch.tensor([999.0 - 300 * k] * 2).cuda()
        ctx = (tiny["context"] * (1 + 0.5 * k)).cuda()
        sem = (tiny["semantic_feature"] * (1 - 0.25 * k)).cuda()
        dit.InferValueRegistry.clear()
        dit.InferValueRegistry.register("semantic_feature", sem)
        eager = warp(x, t, {"crossattn": ctx}, idx=t).clone()
        got = gw(x, t, {"crossattn": ctx}, idx=t).clone()
        torch.cuda.synchronize()
        assert torch.equal(eager, got), f"replay {k} differs from the eager launch"
    assert gw.replays == 3 and len(gw._graphs) == 1
    # sampler trajectory: eager vs graphed network, same noise
    sampler = VPSDEDPMPP2MSampler(num_steps=50, device="cuda")
    x0 = tiny["x"][:1].cuda()
    cond = {"crossattn": tiny["context"][1:].cuda().bfloat16()}
    uc = {"crossattn": torch.zeros_like(cond["crossattn"])}
    dit.InferValueRegistry.register("semantic_feature", tiny["semantic_feature"].cuda())
    outs = []
    for net in (warp, gw):
        gen = torch.Generator().manual_seed(11)
        noise = lambda t_: torch.randn(t_.shape, generator=gen).to(t_.device)
        outs.append(sampler.sample(net, x0.clone(), cond, uc, max_steps=4, noise_fn=noise).clone())
    torch.cuda.synchronize()
    dit.InferValueRegistry.clear()
    assert torch.equal(outs[0], outs[1])
