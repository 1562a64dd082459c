"""CPU: the host-side logic of the PRODUCT sampler (landiff_b200/sampling.py — discretizer, DynamicCFG schedule,
DPM++(2M) SDE step scalars, DiscreteDenoiser table quantisation, acceptance of the reference's YAML blocks) against
tests/golden/schedule.json, which oracle/make_golden.py produced by executing the reference's own discretizer,
guider, denoiser and sampler.  (The per-element update itself is the fused CUDA kernel, covered by the -m gpu tests.)"""
import json

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from landiff_b200 import sampling as S

REF = "landiff.diffusion.sgm.modules.diffusionmodules."
YAML_DISC = {"target": REF + "discretizer.ZeroSNRDDPMDiscretization", "params": {"shift_scale": 3.0}}
YAML_GUIDER = {"target": REF + "guiders.DynamicCFG", "params": {"scale": 6, "exp": 5, "num_steps": 50}}


@pytest.fixture(scope="module")
def gold():
    return json.loads((GOLDEN / "schedule.json").read_text())


@pytest.fixture(scope="module")
def sampler():
    # the reference's YAML blocks, unchanged (…video_vq.yaml sampler_config): class paths are mapped to the local restatements
    return S.VPSDEDPMPP2MSampler(num_steps=50, discretization_config=YAML_DISC, guider_config=YAML_GUIDER, device="cpu")


def test_schedule_and_timesteps(sampler, gold):
    acs, timesteps = sampler.prepare_sampling_loop()
    np.testing.assert_allclose(acs.numpy(), np.array(gold["alphas_cumprod_sqrt"], dtype=np.float32), rtol=0, atol=1e-7)
    assert [int(t) for t in timesteps] == gold["timesteps"]
    assert float(acs[0]) == 0.0 and float(acs[-1]) == 1.0     # zero terminal SNR, appended 1


def test_denoiser_table(sampler, gold):
    tab = sampler._table
    assert tab.numel() == gold["table_len"]
    np.testing.assert_allclose(tab[:8].numpy(), np.array(gold["table_head"], dtype=np.float32), rtol=0, atol=1e-7)
    np.testing.assert_allclose(tab[-8:].numpy(), np.array(gold["table_tail"], dtype=np.float32), rtol=0, atol=1e-7)
    assert abs(float(tab.double().sum()) - gold["table_sum"]) < 1e-4
    den = S.DiscreteDenoiser(discretization_config=YAML_DISC, scaling_config={"target": REF + "denoiser_scaling.VideoScaling"},
                             quantize_c_noise=False)
    assert torch.equal(den.sigmas, tab)


def test_cfg_scale_schedule(sampler, gold):
    for t, v in gold["cfg_scale_by_timestep"].items():
        # the sampler calls scale_schedule(None, num_steps - timestep) (sampling.py:600-606)
        assert abs(sampler.guider.scale_schedule(None, 50 - int(t)) - v) < 1e-12, t


def test_step_scalars_and_quantisation_every_step(sampler, gold):
    acs, _ = sampler.prepare_sampling_loop()
    for st in gold["steps"]:
        i = st["i"]
        m1, m2, m3, m4, mn = sampler.step_scalars(None if i == 0 else acs[i - 1], acs[i], acs[i + 1])
        got = [m1, m2] + ([] if i == 0 else [m3, m4])
        np.testing.assert_allclose(got, st["mult"], rtol=1e-6, atol=1e-7, err_msg=f"step {i}")
        if np.isnan(st["mult_noise"]):
            assert np.isnan(mn)
        else:
            assert abs(mn - st["mult_noise"]) <= 1e-6 * max(1.0, abs(st["mult_noise"]))
        assert abs(float(sampler.quantize(acs[i])) - st["a_quantized"]) < 1e-7


def test_unsupported_sampler_options_fail_loudly():
    with pytest.raises(NotImplementedError):
        S.VPSDEDPMPP2MSampler(num_steps=50, sdedit=True, device="cpu")
    with pytest.raises(NotImplementedError):
        S.ZeroSNRDDPMDiscretization(keep_start=True)
    with pytest.raises(NotImplementedError):
        S.DynamicCFG(6, 5, 50, dyn_thresh_config={"target": "x"})
    with pytest.raises(NotImplementedError):
        S.DiscreteDenoiser(discretization_config=YAML_DISC, quantize_c_noise=True)


def test_guider_prepare_inputs_order():
    g = S.DynamicCFG(6, 5, 50)
    x, s = torch.ones(1, 2, 3), torch.ones(1)
    c, uc = {"crossattn": torch.full((1, 2, 2), 2.0)}, {"crossattn": torch.zeros(1, 2, 2)}
    x2, s2, c2 = g.prepare_inputs(x, s, c, uc)
    assert x2.shape[0] == 2 and s2.shape[0] == 2
    assert torch.equal(c2["crossattn"][0], uc["crossattn"][0]) and torch.equal(c2["crossattn"][1], c["crossattn"][0])  # uncond first
