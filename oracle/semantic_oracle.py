"""TEST INFRASTRUCTURE ONLY — never imported by landiff_b200/ (tests/test_abi.py enforces it).

CPU fp32 restatement of the semantic conditioner's upsample path, written with
torch.nn.functional on a plain state dict carrying the reference's parameter names:
  SemanticCond.video_theia_interpolate_forward + forward   landiff/diffusion/semantic_models/condition.py:86-137
  Decoder.forward                                          .../modules/vq_gan_blocks.py:577-606
  ResnetBlock.forward (temb = None, dropout 0)             .../modules/vq_gan_blocks.py:126-147
  Upsample.forward, pixelshuffle flavour                   .../modules/vq_gan_blocks.py:59-66
  Normalize = GroupNorm(32, eps 1e-6), nonlinearity=swish  .../modules/vq_gan_blocks.py:30-38
Pinned against the reference's own modules by tests/golden/semantic_ref.pt (oracle/make_semantic_golden.py) in
tests/test_semantic.py::test_oracle_matches_reference_golden.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List

import torch
import torch.nn.functional as F


@dataclass
class SemanticConfig:
    """Defaults = the shipped YAML (cogvideox_2b_control_theia_interpolate_video_vq.yaml:53-80)."""
    z_channels: int = 768
    ch: int = 512
    ch_mult: List[float] = field(default_factory=lambda: [0.25, 1])
    num_res_blocks: int = 4
    out_ch: int = 64       # = SemanticCond.out_dim
    target_dim: int = 16

    def decoder_params(self) -> dict:
        return dict(z_channels=self.z_channels, resolution=16, in_channels=512, out_ch=self.out_ch, ch=self.ch,
                    ch_mult=list(self.ch_mult), num_res_blocks=self.num_res_blocks, attn_resolutions=[], dropout=0.0,
                    use_mid_attention=False, upsample_type="pixelshuffle")

    def cond_kwargs(self, decoder_target: str, dtype) -> dict:
        return dict(semantic_model_config={"target": "torch.nn.Identity"},
                    upsample_model_config={"target": decoder_target, "params": self.decoder_params()},
                    dtype=dtype, out_dim=self.out_ch, target_dim=self.target_dim, feature_type="video_theia_interpolate",
                    zero_init_conv_out=True)


SHIPPED = SemanticConfig()
SMALL = SemanticConfig(z_channels=128, ch=256, ch_mult=[0.25, 1], num_res_blocks=1)   # widths 256 -> 64: fast CPU cases


def param_shapes(cfg: SemanticConfig) -> Dict[str, tuple]:
    """Parameter names and shapes of SemanticCond (without the semantic_model), in the reference's naming."""
    out: Dict[str, tuple] = {}

    def conv(name, cin, cout, k=3):
        out[name + ".weight"] = (cout, cin, k, k)
        out[name + ".bias"] = (cout,)

    def norm(name, c):
        out[name + ".weight"] = (c,)
        out[name + ".bias"] = (c,)

    def res(name, cin, cout):
        norm(name + ".norm1", cin); conv(name + ".conv1", cin, cout)
        norm(name + ".norm2", cout); conv(name + ".conv2", cout, cout)
        if cin != cout:
            conv(name + ".nin_shortcut", cin, cout, 1)

    u = "upsample_model."
    nres = len(cfg.ch_mult)
    block_in = int(cfg.ch * cfg.ch_mult[-1])
    conv(u + "conv_in", cfg.z_channels, block_in)
    res(u + "mid.block_1", block_in, block_in)
    res(u + "mid.block_2", block_in, block_in)
    for lvl in reversed(range(nres)):
        block_out = int(cfg.ch * cfg.ch_mult[lvl])
        for j in range(cfg.num_res_blocks + 1):
            res(f"{u}up.{lvl}.block.{j}", block_in, block_out)
            block_in = block_out
        if lvl != 0:
            conv(f"{u}up.{lvl}.upsample.conv", block_in // 4, block_in)
    norm(u + "norm_out", block_in)
    conv(u + "conv_out", block_in, cfg.out_ch)
    conv("conv_out", cfg.out_ch, cfg.target_dim)
    return out


def random_state_dict(cfg: SemanticConfig, seed: int) -> Dict[str, torch.Tensor]:
    """Seeded, bf16-representable parameters: conv weights N(0, 1/fan_in), biases N(0, 0.1^2), GroupNorm weights
    1 + N(0, 0.1^2); the zero-initialised conv_out is randomised too so the path carries signal."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in sorted(param_shapes(cfg).items()):
        noise = torch.randn(shape, generator=g)
        if len(shape) == 4:
            v = noise / (shape[1] * shape[2] * shape[3]) ** 0.5
        elif ".norm" in name and name.endswith("weight"):
            v = 1.0 + 0.1 * noise
        else:
            v = 0.1 * noise
        sd[name] = v.bfloat16().float()
    return sd


def _swish(x):
    return x * torch.sigmoid(x)


def _gn(x, sd, name):
    return F.group_norm(x, 32, sd[name + ".weight"], sd[name + ".bias"], eps=1e-6)


def _conv(x, sd, name, pad=1):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=1, padding=pad)


def _res(x, sd, name):
    h = _conv(_swish(_gn(x, sd, name + ".norm1")), sd, name + ".conv1")
    h = _conv(_swish(_gn(h, sd, name + ".norm2")), sd, name + ".conv2")
    if name + ".nin_shortcut.weight" in sd:
        x = _conv(x, sd, name + ".nin_shortcut", pad=0)
    return x + h


def semantic_oracle(sd: Dict[str, torch.Tensor], features: torch.Tensor, cfg: SemanticConfig) -> torch.Tensor:
    """features [B, T, z_channels, h, w] -> [B, T, target_dim, 2h, 2w]  (forward(semantic_feature_before_upsample=...))."""
    B, T = features.shape[:2]
    sd = {k: v.to(features.dtype) for k, v in sd.items()}
    u = "upsample_model."
    with torch.no_grad():
        h = _conv(features.reshape(B * T, *features.shape[2:]), sd, u + "conv_in")
        h = _res(h, sd, u + "mid.block_1")
        h = _res(h, sd, u + "mid.block_2")
        for lvl in reversed(range(len(cfg.ch_mult))):
            for j in range(cfg.num_res_blocks + 1):
                h = _res(h, sd, f"{u}up.{lvl}.block.{j}")
            if lvl != 0:
                h = _conv(F.pixel_shuffle(h, 2), sd, f"{u}up.{lvl}.upsample.conv")
        h = _conv(_swish(_gn(h, sd, u + "norm_out")), sd, u + "conv_out")
        h = _conv(h, sd, "conv_out")
    return h.reshape(B, T, *h.shape[1:])


def features_for(cfg: SemanticConfig, seed: int, B: int, T: int, h: int, w: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, T, cfg.z_channels, h, w, generator=g).bfloat16().float()
