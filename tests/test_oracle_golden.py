"""CPU: the oracle restatement (oracle/dit_oracle.py) against the committed golden vectors that were produced by the
reference's own code (oracle/make_golden.py)."""
import json

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import dit_oracle as O


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


@pytest.fixture(scope="module")
def tiny():
    return torch.load(GOLDEN / "tiny_warp.pt", weights_only=False)


def test_tiny_config_matches(tiny):
    for k, v in tiny["cfg"].items():
        if hasattr(O.TINY, k):
            assert getattr(O.TINY, k) == v, k


def test_state_dict_keys_match_reference(tiny):
    assert set(O.param_shapes(O.TINY, control=True)) == set(tiny["sd_ctrl"].keys())
    assert set(O.param_shapes(O.TINY, control=False)) == set(tiny["sd_main"].keys())
    for k, shp in O.param_shapes(O.TINY, control=False).items():
        assert tuple(tiny["sd_main"][k].shape) == tuple(shp), k


def test_pos_embedding_table_bit_exact(tiny):
    cfg = O.TINY
    tab = O.pos_embed_3d(cfg.hidden_size, cfg.latent_h // 2, cfg.latent_w // 2, cfg.latent_t, cfg.interp, cfg.interp)
    ref = tiny["sd_main"]["mixins.pos_embed.pos_embedding"][0, cfg.text_length:]
    assert torch.equal(tab.to(torch.bfloat16), ref)
    assert torch.count_nonzero(tiny["sd_main"]["mixins.pos_embed.pos_embedding"][0, :cfg.text_length]) == 0


def test_warp_forward_fp32_matches_reference(tiny):
    sdc, sdm = O.cast_state_dict(tiny["sd_ctrl"], torch.float32), O.cast_state_dict(tiny["sd_main"], torch.float32)
    out, ctrl = O.warp_forward(sdc, sdm, O.TINY, tiny["x"], tiny["t"], tiny["context"], tiny["semantic_feature"],
                               return_control=True)
    assert rel(out, tiny["out"]) < 1e-5
    for a, b in zip(ctrl, tiny["control_hidden"]):
        assert rel(a, b) < 1e-5
    out2 = O.warp_forward(sdc, sdm, O.TINY, tiny["x"], tiny["t2"], tiny["context"], tiny["semantic_feature"])
    assert rel(out2, tiny["out2"]) < 1e-5
    assert rel(out2, tiny["out"]) > 1e-2  # the timestep matters


def test_schedule_tables():
    g = json.loads((GOLDEN / "schedule.json").read_text())
    s = O.OracleSampler(50)
    acs = torch.cat([s.sigmas, s.sigmas.new_ones(1)])
    np.testing.assert_allclose(acs.numpy(), np.array(g["alphas_cumprod_sqrt"], dtype=np.float32), rtol=0, atol=1e-7)
    assert [-1] + [int(t) for t in s.timesteps] == g["timesteps"]
    np.testing.assert_allclose(s.table[:8].numpy(), np.array(g["table_head"], dtype=np.float32), rtol=0, atol=1e-7)
    np.testing.assert_allclose(s.table[-8:].numpy(), np.array(g["table_tail"], dtype=np.float32), rtol=0, atol=1e-7)
    assert s.table.numel() == g["table_len"]
    assert abs(float(s.table.double().sum()) - g["table_sum"]) < 1e-4
    for t, v in g["cfg_scale_by_timestep"].items():
        assert abs(O.dynamic_cfg_scale(int(t)) - v) < 1e-12
    # values pinned by executing the reference
    assert abs(O.dynamic_cfg_scale(999) - 2.291487174551452) < 1e-12
    assert abs(O.dynamic_cfg_scale(19) - 1.123397939968981) < 1e-12


def test_dpmpp_scalars_every_step():
    g = json.loads((GOLDEN / "schedule.json").read_text())
    s = O.OracleSampler(50)
    acs = torch.cat([s.sigmas, s.sigmas.new_ones(1)])
    for st in g["steps"]:
        i = st["i"]
        m = O.dpmpp2m_scalars(None if i == 0 else acs[i - 1], acs[i], acs[i + 1])
        got = [float(m["m1"]), float(m["m2"])] + ([] if i == 0 else [float(m["m3"]), float(m["m4"])])
        np.testing.assert_allclose(got, st["mult"], rtol=1e-6, atol=1e-7, err_msg=f"step {i}")
        exp_mn = st["mult_noise"]
        if np.isnan(exp_mn):
            assert np.isnan(float(m["mn"]))
        else:
            assert abs(float(m["mn"]) - exp_mn) <= 1e-6 * max(1.0, abs(exp_mn))
        assert abs(float(s.quantize(acs[i])) - st["a_quantized"]) < 1e-7
    # first step: zero terminal SNR -> h = +inf limits
    m0 = O.dpmpp2m_scalars(None, acs[0], acs[1])
    assert float(m0["m1"]) == 0.0 and abs(float(m0["m2"]) + float(acs[1])) < 1e-7


def test_sampler_trajectory_toy_network():
    g = torch.load(GOLDEN / "sampler_toy.pt", weights_only=False)

    def network(x2, t2, ctx2):
        c = ctx2.mean(dim=(1, 2)).view(-1, 1, 1, 1, 1)
        return (torch.tanh(x2 * 0.5 + c) * (1.0 + t2.view(-1, 1, 1, 1, 1) / 1000.0)).to(torch.bfloat16)

    s = O.OracleSampler(g["num_steps"])
    gen = torch.Generator().manual_seed(g["seed"])
    out = s(network, g["x0"].clone(), g["cond"], g["uc"], gen)
    assert rel(out, g["out"]) < 1e-6
    # fixed_frames (streaming prefix): the prefix frames come back untouched, the rest follows the reference
    s2 = O.OracleSampler(g["num_steps"], fixed_frames=g["fixed_frames"])
    gen = torch.Generator().manual_seed(g["seed"])
    out2 = s2(network, g["x0"].clone(), g["cond"], g["uc"], gen)
    assert torch.equal(out2[:, :g["fixed_frames"]], g["x0"][:, :g["fixed_frames"]])
    assert torch.equal(g["out_fixed_frames"][:, :g["fixed_frames"]], g["x0"][:, :g["fixed_frames"]])
    assert rel(out2, g["out_fixed_frames"]) < 1e-6
    assert rel(out2[:, :g["fixed_frames"]], g["out"][:, :g["fixed_frames"]]) > 1e-2   # without the hook they evolve


@pytest.mark.parametrize("tag,strong", [("weak", False), ("strong", True)])
def test_small_warp_b_golden(tag, strong):
    """Second reference-generated golden (oracle/make_golden.py make_small_warp_b): 3 control layers feeding a 5-layer
    main net (zero-linear chaining, control add only for the first 3 layers), 3 heads, 3 latent frames, text length 7,
    weak and strong init, timestep 519.  The file carries the seed, not the weights: `seeded_state_dict` restates the
    generator's init, and a few stored parameter values prove the regenerated weights are the reference's."""
    from oracle.make_golden import small_b_inputs

    g = torch.load(GOLDEN / "small_warp_b.pt", weights_only=False)
    cfg = O.OracleConfig(**g["cfg"])
    sdc = O.seeded_state_dict(cfg, True, g["seed"], strong)
    sdm = O.seeded_state_dict(cfg, False, g["seed"] + 1, strong)
    for k, v in g[tag]["probe"].items():
        assert torch.equal(sdc[k].float().flatten()[:4], v), k
    x, ctx, sem, t = small_b_inputs(cfg)
    out, ctl = O.warp_forward(O.cast_state_dict(sdc, torch.float32), O.cast_state_dict(sdm, torch.float32), cfg, x, t,
                              ctx, sem, return_control=True)
    assert len(ctl) == cfg.control_layers == 3
    for i, (a, b) in enumerate(zip(ctl, g[tag]["control_hidden"])):
        assert ((a - b).norm() / b.norm()).item() < 1e-5, f"control layer {i}"
    assert ((out - g[tag]["out"]).norm() / g[tag]["out"].norm()).item() < 1e-5
    # the control branch matters in this golden: dropping it changes the output far beyond the tolerance
    no_ctrl = O.main_forward(O.cast_state_dict(sdm, torch.float32), cfg, x, t, ctx, None)
    assert ((no_ctrl - g[tag]["out"]).norm() / g[tag]["out"].norm()).item() > 1e-2


@pytest.mark.parametrize("tag,strong", [("weak", False), ("strong", True)])
def test_config1_real_width_golden_from_reference_modules(tag, strong):
    """d = 1920, 30 heads, N = 886 (BASELINE config 1; `weak`: the full 15 + 30 layers): the oracle restatement against the
    output of the reference's own ControlDiffusionTransformer -> DiffusionTransformer (oracle/make_config1_golden.py)."""
    import dataclasses

    gold = torch.load(GOLDEN / "config1_ref.pt", weights_only=False)[tag]
    cfg = dataclasses.replace(O.CONFIG1, main_layers=gold["main_layers"], control_layers=gold["control_layers"])
    sdc = O.cast_state_dict(O.random_state_dict(cfg, True, seed=10, strong=strong), torch.float32)
    sdm = O.cast_state_dict(O.random_state_dict(cfg, False, seed=11, strong=strong), torch.float32)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, cfg.latent_t, 16, cfg.latent_h, cfg.latent_w, generator=g)
    ctx = (torch.randn(2, cfg.text_length, cfg.text_hidden, generator=g) * 0.2).bfloat16().float()
    ctx[0] = 0
    sem = (torch.randn(1, cfg.latent_t, 16, cfg.latent_h, cfg.latent_w, generator=g) * 0.1).bfloat16().float()
    out, ctl = O.warp_forward(sdc, sdm, cfg, x, torch.tensor([519.0, 519.0]), ctx, sem, return_control=True)
    r = ((out - gold["out"]).norm() / gold["out"].norm()).item()
    rc = ((ctl[-1][:, ::37] - gold["control_last"]).norm() / gold["control_last"].norm()).item()
    assert r < 2e-5 and rc < 2e-5, f"oracle vs reference modules at config 1 [{tag}]: rel-L2 {r:.3e} / control {rc:.3e}"


def test_oracle_trajectory_equals_reference_object_trajectory():
    """The oracle's own 50-step config-1 trajectory (trajectory50_config1.pt, OracleSampler driving the oracle network)
    against the trajectory the REFERENCE sampler + denoiser + guider + network objects produced from the same seeds
    (trajectory50_config1_ref.pt): the restatement of the whole sampling loop is pinned at the real width and depth."""
    a = torch.load(GOLDEN / "trajectory50_config1.pt", weights_only=False)
    b = torch.load(GOLDEN / "trajectory50_config1_ref.pt", weights_only=False)
    assert sorted(a["steps"]) == sorted(b["steps"])
    for i in a["steps"]:
        r = ((a["steps"][i] - b["steps"][i]).norm() / b["steps"][i].norm()).item()
        assert r < 1e-3, f"step {i + 1}: oracle vs reference trajectory rel-L2 {r:.3e}"
    r = ((a["final"] - b["final"]).norm() / b["final"].norm()).item()
    assert r < 1e-3, f"final latent: oracle vs reference trajectory rel-L2 {r:.3e}"
