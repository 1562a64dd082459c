"""TEST INFRASTRUCTURE ONLY.  Builds the REFERENCE modules (imported unchanged from /root/reference on top of
oracle/sat_shim.py) at an arbitrary size, with the seeded random init of SURVEY.md §8d.  Only usable in the
build container (the GPU box has no /root/reference); used by oracle/make_golden.py and the CPU tests that
validate oracle/dit_oracle.py against the reference code itself.
"""
from __future__ import annotations

import copy
import os
from dataclasses import dataclass

import torch

from . import sat_shim


@dataclass
class DiTConfig:
    """Shape of the ControlDiffWarp pair.  Defaults = the shipped 2B YAML (…video_vq.yaml:25-150)."""
    hidden_size: int = 1920
    num_heads: int = 30
    main_layers: int = 30
    control_layers: int = 15
    time_embed_dim: int = 512
    text_hidden: int = 4096
    text_length: int = 226
    latent_t: int = 13          # (num_frames-1)//4 + 1
    latent_h: int = 60
    latent_w: int = 90
    in_channels: int = 16
    patch_size: int = 2
    interp: float = 1.875

    @property
    def num_frames(self):
        return (self.latent_t - 1) * 4 + 1

    @property
    def n_img(self):
        return self.latent_t * (self.latent_h // 2) * (self.latent_w // 2)

    @property
    def n_tok(self):
        return self.text_length + self.n_img


TINY = DiTConfig(hidden_size=128, num_heads=2, main_layers=2, control_layers=1, time_embed_dim=64, text_hidden=64,
                 text_length=6, latent_t=2, latent_h=8, latent_w=12)
CONFIG1 = DiTConfig(latent_t=2, latent_h=30, latent_w=44)  # BASELINE config 1 (240x352)
FULL = DiTConfig()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(sat_shim.REFERENCE_ROOT, "landiff"))


def _params(cfg: DiTConfig, layers: int, control: bool):
    mods = {
        "pos_embed_config": {"target": "landiff.diffusion.dit_video_concat.Basic3DPositionEmbeddingMixin",
                             "params": {"text_length": cfg.text_length, "height_interpolation": cfg.interp,
                                        "width_interpolation": cfg.interp}},
        "patch_embed_config": {"target": "landiff.diffusion.dit_video_concat.ImagePatchEmbeddingMixin",
                               "params": {"text_hidden_size": cfg.text_hidden}},
    }
    if control:
        mods["semantic_condition_config"] = {"target": "torch.nn.Identity"}
        mods["adaln_layer_config"] = {"target": "landiff.diffusion.dit_video_concat.ControlOutAdaLNMixin",
                                      "params": {"qk_ln": True, "use_zero_linears": True}}
        mods["final_layer_config"] = {"target": "landiff.diffusion.dit_video_concat.EmptyFinalLayerMixin"}
    else:
        mods["adaln_layer_config"] = {"target": "landiff.diffusion.dit_video_concat.ControlAdaLNMixin",
                                      "params": {"qk_ln": True, "use_semantic_injection_adaln": False,
                                                 "control_layers": cfg.control_layers}}
        mods["final_layer_config"] = {"target": "landiff.diffusion.dit_video_concat.FinalLayerMixin"}
    p = dict(time_embed_dim=cfg.time_embed_dim, elementwise_affine=True, num_frames=cfg.num_frames,
             time_compressed_rate=4, latent_width=cfg.latent_w, latent_height=cfg.latent_h, num_layers=layers,
             patch_size=cfg.patch_size, in_channels=cfg.in_channels, out_channels=cfg.in_channels,
             hidden_size=cfg.hidden_size, adm_in_channels=256, num_attention_heads=cfg.num_heads,
             transformer_args=sat_shim.transformer_args(), modules=mods)
    if control:
        p["use_semantic_injection_adaln"] = False
    return p


def seeded_init_(module: torch.nn.Module, seed: int, strong: bool = False) -> None:
    """SURVEY §8d init: every weight N(0,0.02^2), biases N(0,0.02^2), LayerNorm weight 1+N(0,0.02^2); zero-init tensors
    (zero_linears) are re-randomised too so the control path carries signal; pos_embedding is left as built.
    `strong=True` is the structural-test init: matrices N(0, 1/fan_in), biases / norm offsets N(0, 0.1^2), so that the
    time embedding, every adaLN shift/scale/gate and the control branch are O(1) and a mis-ordered chunk or a wrong
    segment shows up as an O(1) error instead of hiding under the tolerance."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(module.named_parameters()):
            if name.endswith("pos_embedding"):
                continue
            noise = torch.randn(p.shape, generator=g, dtype=torch.float32)
            if strong:
                std = (1.0 / (p[0].numel() ** 0.5)) if p.dim() >= 2 else 0.1
            else:
                std = 0.02
            is_ln_weight = name.endswith("weight") and p.dim() == 1
            p.copy_((1.0 + noise * std) if is_ln_weight else noise * std)


def build_reference(cfg: DiTConfig, seed: int = 0, dtype: str = "fp32", strong: bool = False):
    """Returns (control_model, main_model): reference ControlDiffusionTransformer / DiffusionTransformer."""
    sat_shim.install()
    from landiff.diffusion import dit_video_concat as ref  # the unchanged reference module

    ctrl = ref.ControlDiffusionTransformer(**copy.deepcopy(_params(cfg, cfg.control_layers, True)), dtype=dtype)
    main = ref.DiffusionTransformer(**copy.deepcopy(_params(cfg, cfg.main_layers, False)), dtype=dtype)
    seeded_init_(ctrl, seed, strong)
    seeded_init_(main, seed + 1, strong)
    ctrl.eval()
    main.eval()
    return ctrl, main


def reference_forward(ctrl, main, x, t, context, semantic_feature):
    """ControlDiffWarp.forward (dit_video_concat.py:1196-1200) without the checkpoint-reading ctor."""
    from landiff.diffusion.sgm.util import InferValueRegistry

    InferValueRegistry.clear()
    InferValueRegistry.register("semantic_feature", semantic_feature)
    with torch.no_grad():
        control_layers_output = ctrl(x, timesteps=t, context=context, idx=t)
        out = main(x, timesteps=t, context=context, idx=t, control_layers_output=control_layers_output)
    InferValueRegistry.clear()
    return out, control_layers_output
