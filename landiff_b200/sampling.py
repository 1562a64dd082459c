"""Sampler / denoiser / guider plugin surface of the hot path, B200-native.

Same class names and constructor kwargs as the reference's
`sgm/modules/diffusionmodules/{sampling,denoiser,guiders,discretizer,denoiser_scaling}.py`, so the YAML
`sampler_config` / `denoiser_config` can point here.  The host side only computes per-step SCALARS (schedule
tables, CFG scale, DPM-Solver++ coefficients — a few dozen flops per step, done on the CPU in fp32 torch exactly
like the reference expressions so the IEEE limits at zero terminal SNR come out identically); all per-element work
(denoiser scaling + CFG combine + DPM++(2M) SDE update, ~15 eager kernels in the reference) is ONE fused CUDA kernel
(`ld_sampler_update`).  Noise is drawn with `torch.randn_like` in the reference's order so a seeded run consumes the
identical RNG stream.  No per-step host sync (the reference has one at sampling.py:772).

Reference map:  ZeroSNRDDPMDiscretization discretizer.py:80-141 · DiscreteDenoiser denoiser.py:44-77 ·
VideoScaling denoiser_scaling.py:62-70 · DynamicCFG guiders.py:58-79 · VanillaCFG.prepare_inputs guiders.py:46-55 ·
VideoDDIMSampler sampling.py:538-611 · VPSDEDPMPP2MSampler sampling.py:678-837.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import numpy as np
import torch

from . import ops
from .dit import instantiate_from_config


class ZeroSNRDDPMDiscretization:
    def __init__(self, linear_start=0.00085, linear_end=0.0120, num_timesteps=1000, shift_scale=1.0, keep_start=False,
                 post_shift=False):
        if keep_start or post_shift:
            raise NotImplementedError("keep_start / post_shift are not used by the shipped LanDiff config")
        self.num_timesteps = num_timesteps
        betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, num_timesteps, dtype=torch.float64) ** 2).numpy()
        ac = np.cumprod(1.0 - betas, axis=0)
        self.alphas_cumprod = ac / (shift_scale + (1 - shift_scale) * ac)  # SNR shift
        self.shift_scale = shift_scale

    def get_sigmas(self, n, device="cpu", return_idx=False):
        if n < self.num_timesteps:
            timesteps = np.linspace(self.num_timesteps - 1, 0, n, endpoint=False).astype(int)[::-1]
            ac = self.alphas_cumprod[timesteps]
        elif n == self.num_timesteps:
            timesteps = np.arange(self.num_timesteps)
            ac = self.alphas_cumprod
        else:
            raise ValueError
        s = torch.tensor(ac, dtype=torch.float32).sqrt()
        s0, sT = s[0].clone(), s[-1].clone()
        s = (s - sT) * (s0 / (s0 - sT))  # zero terminal SNR rescale
        s = torch.flip(s, (0,))
        return (s, timesteps) if return_idx else s

    def __call__(self, n, do_append_zero=True, device="cpu", flip=False, return_idx=False):
        out = self.get_sigmas(n, device=device, return_idx=return_idx)
        sigmas, idx = out if return_idx else (out, None)
        if do_append_zero:
            sigmas = torch.cat([sigmas, sigmas.new_zeros([1])])
        if flip:
            sigmas = torch.flip(sigmas, (0,))
        return (sigmas, idx) if return_idx else sigmas


class VideoScaling:
    def __call__(self, alphas_cumprod_sqrt, **additional_model_inputs):
        c_skip = alphas_cumprod_sqrt
        c_out = -((1 - alphas_cumprod_sqrt ** 2) ** 0.5)
        c_in = torch.ones_like(alphas_cumprod_sqrt)
        c_noise = additional_model_inputs["idx"].clone()
        return c_skip, c_out, c_in, c_noise


class DynamicCFG:
    """scale(step_index) = 1 + s (1 - cos(pi (step_index / num_steps)^exp)) / 2; the sampler calls it with
    step_index = num_steps - timestep (sampling.py:600-606), far outside [0, 1] — reproduced as is."""

    def __init__(self, scale, exp, num_steps, dyn_thresh_config=None):
        if dyn_thresh_config is not None:
            raise NotImplementedError("only NoDynamicThresholding is implemented")
        self.scale, self.exp, self.num_steps = scale, exp, num_steps

    def scale_schedule(self, sigma, step_index):
        return 1 + self.scale * (1 - math.cos(math.pi * (step_index / self.num_steps) ** self.exp)) / 2

    def prepare_inputs(self, x, s, c, uc):
        c_out = dict()
        for k in c:
            if k in ["vector", "crossattn", "concat"]:
                c_out[k] = torch.cat((uc[k], c[k]), 0)  # uncond first
            else:
                assert c[k] == uc[k]
                c_out[k] = c[k]
        return torch.cat([x] * 2), torch.cat([s] * 2), c_out

    def __call__(self, x, sigma, step_index, scale=None):
        x_u, x_c = x.chunk(2)
        return x_u + self.scale_schedule(sigma, float(step_index)) * (x_c - x_u)


class DiscreteDenoiser(torch.nn.Module):
    """network(input*c_in, c_noise, cond) * c_out + input * c_skip with sigma snapped to the 1000-entry table."""

    def __init__(self, weighting_config=None, scaling_config=None, num_idx=1000, discretization_config=None,
                 do_append_zero=False, quantize_c_noise=True, flip=True):
        super().__init__()
        disc = instantiate_from_config(_localise(discretization_config))
        self.sigmas = disc(num_idx, do_append_zero=do_append_zero, flip=flip)  # CPU fp32
        if quantize_c_noise:
            raise NotImplementedError("quantize_c_noise=False in the shipped config")
        self.scaling = VideoScaling()

    def quantize_sigma_host(self, a: float) -> float:
        return float(self.sigmas[(torch.tensor(a, dtype=torch.float32) - self.sigmas).abs().argmin()])

    def forward(self, network, input, sigma, cond, **additional_model_inputs):
        tab = self.sigmas.to(sigma.device)
        sigma = tab[(sigma - tab[:, None]).abs().argmin(dim=0)]
        shape = sigma.shape
        sigma = sigma[(...,) + (None,) * (input.ndim - sigma.ndim)]
        c_skip, c_out, c_in, c_noise = self.scaling(sigma, **additional_model_inputs)
        return network(input * c_in, c_noise.reshape(shape), cond, **additional_model_inputs) * c_out + input * c_skip


_LOCAL = {
    "ZeroSNRDDPMDiscretization": "landiff_b200.sampling.ZeroSNRDDPMDiscretization",
    "DynamicCFG": "landiff_b200.sampling.DynamicCFG",
    "VideoScaling": "landiff_b200.sampling.VideoScaling",
}


def _localise(config):
    """Accept the reference's YAML blocks unchanged: map the reference class paths of the tiny host-side helpers to
    the restatements in this module (the reference package itself is not required at run time)."""
    if config is None:
        return None
    cls = config["target"].rsplit(".", 1)[-1]
    if cls in _LOCAL:
        return {"target": _LOCAL[cls], "params": dict(config.get("params", {}) or {})}
    return config


DEFAULT_DISCRETIZATION = {"target": "landiff_b200.sampling.ZeroSNRDDPMDiscretization", "params": {"shift_scale": 3.0}}
DEFAULT_GUIDER = {"target": "landiff_b200.sampling.DynamicCFG", "params": {"scale": 6, "exp": 5, "num_steps": 50}}


class VPSDEDPMPP2MSampler:
    """DPM-Solver++(2M) SDE sampler for the VP schedule (reference sampling.py:678-837) with the per-element work
    fused into one kernel per step."""

    def __init__(self, discretization_config=None, num_steps=None, guider_config=None, verbose=False, device="cuda",
                 fixed_frames=0, sdedit=False):
        if sdedit:
            raise NotImplementedError("sdedit is not used by the shipped LanDiff config")
        self.num_steps = num_steps
        self.discretization = instantiate_from_config(_localise(discretization_config or DEFAULT_DISCRETIZATION))
        self.guider = instantiate_from_config(_localise(guider_config or DEFAULT_GUIDER))
        if not isinstance(self.guider, DynamicCFG):
            raise NotImplementedError("only DynamicCFG guidance is implemented")
        self.verbose, self.device, self.fixed_frames = verbose, device, fixed_frames
        self._table = self.discretization(1000, do_append_zero=False, flip=True)  # DiscreteDenoiser's table

    # -- host-side scalars ------------------------------------------------------------------------------------
    def prepare_sampling_loop(self, num_steps=None):
        acs, timesteps = self.discretization(self.num_steps if num_steps is None else num_steps, device="cpu",
                                             return_idx=True, do_append_zero=False)
        acs = torch.cat([acs, acs.new_ones([1])])
        timesteps = np.concatenate([[-1], np.asarray(timesteps)])
        return acs, timesteps

    @staticmethod
    def step_scalars(a_prev, a, a_next):
        """get_variables / get_mult / mult_noise (sampling.py:679-720, :766-769) in fp32 torch on the host."""
        lamb = ((a ** 2 / (1 - a ** 2)) ** 0.5).log()
        lamb_next = ((a_next ** 2 / (1 - a_next ** 2)) ** 0.5).log()
        h = lamb_next - lamb
        m1 = ((1 - a_next ** 2) / (1 - a ** 2)) ** 0.5 * (-h).exp()
        m2 = (-2 * h).expm1() * a_next
        mn = (1 - a_next ** 2) ** 0.5 * (1 - (-2 * h).exp()) ** 0.5
        if a_prev is None:
            return float(m1), float(m2), 0.0, 0.0, float(mn)
        lamb_prev = ((a_prev ** 2 / (1 - a_prev ** 2)) ** 0.5).log()
        r = (lamb - lamb_prev) / h
        return float(m1), float(m2), float(1 + 1 / (2 * r)), float(1 / (2 * r)), float(mn)

    def quantize(self, a: torch.Tensor) -> torch.Tensor:
        return self._table[(a - self._table).abs().argmin()]

    # -- the loop -----------------------------------------------------------------------------------------------
    @torch.no_grad()
    def sample(self, network: Callable, x: torch.Tensor, cond: Dict, uc: Dict, num_steps: Optional[int] = None,
               cfg_group=None, step_callback=None, start_step: int = 0, max_steps: Optional[int] = None,
               noise_fn: Optional[Callable] = None, **kwargs) -> torch.Tensor:
        """Fast path: drives `network(x2, t2, cond2, idx=t2)` (batch = [uncond, cond]) directly and applies
        `ld_sampler_update` once per step.  With `cfg_group` (a landiff_b200.parallel.CFGGroup) each rank evaluates
        one batch row and the two bf16 outputs are exchanged over NCCL.  `start_step` / `max_steps` run a slice of
        the schedule (benchmarking); a slice that does not start at 0 begins without DPM++ history.  `noise_fn(x)`
        replaces `torch.randn_like` (same draw order as the reference: one draw per step, a second one on every step
        after the first) so a trajectory can be replayed against a CPU noise stream."""
        randn = torch.randn_like if noise_fn is None else noise_fn
        acs, timesteps = self.prepare_sampling_loop(num_steps)
        n = len(acs) - 1
        total = self.num_steps if num_steps is None else num_steps
        x = x.to(torch.float32).contiguous()
        prefix = x[:, :self.fixed_frames].clone() if self.fixed_frames > 0 else None
        B = x.shape[0]
        if B != 1:
            raise NotImplementedError("the fused sampler processes one video per call (the reference inference path)")
        ctx2 = torch.cat((uc["crossattn"], cond["crossattn"]), 0)
        old = None
        x_next = torch.empty_like(x)
        den_a, den_b = torch.empty_like(x), torch.empty_like(x)
        stop = n if max_steps is None else min(n, start_step + max_steps)
        for i in range(start_step, stop):
            if prefix is not None:
                x[:, :self.fixed_frames] = prefix
            a_prev = None if i == 0 else acs[i - 1]
            a, a_next = acs[i], acs[i + 1]
            timestep = float(timesteps[-(i + 1)])
            aq = self.quantize(a)
            c_skip, c_out = float(aq), float(-((1 - aq ** 2) ** 0.5))
            cfg = self.guider.scale_schedule(None, total - timestep)
            if cfg_group is None:
                if getattr(network, "owned_latent_mask", None) is not None:
                    raise RuntimeError("this network is token-sharded (ring sequence parallel): its output must be assembled "
                                       "across ranks — pass cfg_group=landiff_b200.parallel.CFGGroup(layout)")
                t2 = torch.full((2,), timestep, dtype=torch.float32, device=x.device)
                net = network(torch.cat([x, x]), t2, {"crossattn": ctx2}, idx=t2, **kwargs)
                net_u, net_c = net[0:1], net[1:2]
            else:
                net_u, net_c = cfg_group.evaluate(network, x, timestep, ctx2, **kwargs)
            last = (total - i) == 1
            den_out = den_a if (i % 2 == 0) else den_b
            if last:
                ops.sampler_update(x, net_u, net_c, None, None, c_skip=c_skip, c_out=c_out, cfg=cfg, mode=2, x_out=x_next,
                                   den_out=den_out)
            else:
                m1, m2, m3, m4, mn = self.step_scalars(a_prev, a, a_next)
                eps = randn(x)  # x_standard's draw (sampling.py:771)
                if old is None or float(a_next) < 1e-14:
                    ops.sampler_update(x, net_u, net_c, None, eps, c_skip=c_skip, c_out=c_out, cfg=cfg, m1=m1, m2=m2, mn=mn,
                                       mode=0, x_out=x_next, den_out=den_out)
                else:
                    eps = randn(x)  # the reference draws a second tensor for x_advanced (:776-781)
                    ops.sampler_update(x, net_u, net_c, old, eps, c_skip=c_skip, c_out=c_out, cfg=cfg, m1=m1, m2=m2, m3=m3,
                                       m4=m4, mn=mn, mode=1, x_out=x_next, den_out=den_out)
            old = den_out
            x, x_next = x_next, x
            if step_callback is not None:
                step_callback(i, x)
        if prefix is not None:
            x[:, :self.fixed_frames] = prefix
        return x

    def __call__(self, denoiser, x, cond, uc=None, num_steps=None, scale=None, scale_emb=None, **kwargs):
        """Reference-compatible entry (sampling.py:785-837): `denoiser(input, sigma, c, **kw)` is the engine's lambda
        around DiscreteDenoiser+network.  CFG combine + DPM++ update run in the fused kernel on the fp32 denoised
        rows (c_skip = 0, c_out = 1 there)."""
        acs, timesteps = self.prepare_sampling_loop(num_steps)
        n = len(acs) - 1
        total = self.num_steps if num_steps is None else num_steps
        uc = cond if uc is None else uc
        x = x.to(torch.float32).contiguous()
        prefix = x[:, :self.fixed_frames].clone() if self.fixed_frames > 0 else None
        s_in = x.new_ones([x.shape[0]])
        old = None
        for i in range(n):
            if prefix is not None:
                x = torch.cat([prefix, x[:, self.fixed_frames:]], dim=1)
            a_prev = None if i == 0 else acs[i - 1]
            a, a_next = acs[i], acs[i + 1]
            timestep = float(timesteps[-(i + 1)])
            extra = dict(kwargs)
            extra["idx"] = torch.cat([x.new_ones([x.shape[0]]) * timestep] * 2)
            den2 = denoiser(*self.guider.prepare_inputs(x, s_in * float(a), cond, uc), **extra).to(torch.float32)
            den_u, den_c = [d.contiguous() for d in den2.chunk(2)]
            cfg = self.guider.scale_schedule(None, total - timestep)
            if (total - i) == 1:
                x, old = ops.sampler_update_f32(x, den_u, den_c, None, None, cfg=cfg, mode=2)
                continue
            m1, m2, m3, m4, mn = self.step_scalars(a_prev, a, a_next)
            eps = torch.randn_like(x)
            if old is None or float(a_next) < 1e-14:
                x, old = ops.sampler_update_f32(x, den_u, den_c, None, eps, cfg=cfg, m1=m1, m2=m2, mn=mn, mode=0)
            else:
                eps = torch.randn_like(x)
                x, old = ops.sampler_update_f32(x, den_u, den_c, old, eps, cfg=cfg, m1=m1, m2=m2, m3=m3, m4=m4, mn=mn, mode=1)
        if prefix is not None:
            x = torch.cat([prefix, x[:, self.fixed_frames:]], dim=1)
        return x
