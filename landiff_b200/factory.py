"""Builders for the drop-in network at the shipped 2B shape (or any smaller shape), the way the reference engine
builds it from YAML (`diffusion_video.py:449-480`): ControlDiffWarp(OpenAIWrapper(main), OpenAIWrapper(control))."""
from __future__ import annotations

import argparse
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import dit


@dataclass
class DiTShape:
    """Defaults = configs/cogvideox_2b_control_theia_interpolate_video_vq.yaml:25-150 (49 frames 480x720)."""
    hidden_size: int = 1920
    num_heads: int = 30
    main_layers: int = 30
    control_layers: int = 15
    time_embed_dim: int = 512
    text_hidden: int = 4096
    text_length: int = 226
    latent_t: int = 13
    latent_h: int = 60
    latent_w: int = 90
    in_channels: int = 16
    interp: float = 1.875

    @property
    def n_img(self):
        return self.latent_t * (self.latent_h // 2) * (self.latent_w // 2)

    @property
    def n_tok(self):
        return self.text_length + self.n_img


FULL = DiTShape()
CONFIG1 = DiTShape(latent_t=2, latent_h=30, latent_w=44)   # BASELINE config 1: 5 frames, 240x352 (patch-aligned)
TINY = DiTShape(hidden_size=128, num_heads=2, main_layers=2, control_layers=1, time_embed_dim=64, text_hidden=64,
                text_length=6, latent_t=2, latent_h=8, latent_w=12)


def transformer_args():
    return argparse.Namespace(checkpoint_activations=False, vocab_size=1, max_sequence_length=64, layernorm_order="pre",
                              skip_init=False, model_parallel_size=1, is_decoder=False)


def network_params(cfg, control: bool, target_pkg: str = "landiff.diffusion.dit_video_concat") -> dict:
    """The YAML `params` block of `control_network_config` / `network_config` for shape `cfg`."""
    mods = {
        "pos_embed_config": {"target": f"{target_pkg}.Basic3DPositionEmbeddingMixin",
                             "params": {"text_length": cfg.text_length, "height_interpolation": cfg.interp,
                                        "width_interpolation": cfg.interp}},
        "patch_embed_config": {"target": f"{target_pkg}.ImagePatchEmbeddingMixin",
                               "params": {"text_hidden_size": cfg.text_hidden}},
    }
    if control:
        mods["semantic_condition_config"] = {"target": "torch.nn.Identity"}
        mods["adaln_layer_config"] = {"target": f"{target_pkg}.ControlOutAdaLNMixin",
                                      "params": {"qk_ln": True, "use_zero_linears": True}}
        mods["final_layer_config"] = {"target": f"{target_pkg}.EmptyFinalLayerMixin"}
    else:
        mods["adaln_layer_config"] = {"target": f"{target_pkg}.ControlAdaLNMixin",
                                      "params": {"qk_ln": True, "use_semantic_injection_adaln": False,
                                                 "control_layers": cfg.control_layers}}
        mods["final_layer_config"] = {"target": f"{target_pkg}.FinalLayerMixin"}
    p = dict(time_embed_dim=cfg.time_embed_dim, elementwise_affine=True, num_frames=(cfg.latent_t - 1) * 4 + 1,
             time_compressed_rate=4, latent_width=cfg.latent_w, latent_height=cfg.latent_h,
             num_layers=cfg.control_layers if control else cfg.main_layers, patch_size=2, in_channels=cfg.in_channels,
             out_channels=cfg.in_channels, hidden_size=cfg.hidden_size, adm_in_channels=256,
             num_attention_heads=cfg.num_heads, transformer_args=transformer_args(), modules=mods)
    if control:
        p["use_semantic_injection_adaln"] = False
    return p


def build_warp(cfg=FULL, device: Optional[str] = None, sd_ctrl: Optional[Dict[str, torch.Tensor]] = None,
               sd_main: Optional[Dict[str, torch.Tensor]] = None, meta_init: bool = False) -> dit.ControlDiffWarp:
    """ControlDiffWarp over OpenAIWrapper-wrapped control/main networks (bf16), optionally loaded from state dicts
    with the reference's key names and moved to `device`."""
    ctrl = dit.ControlDiffusionTransformer(**network_params(cfg, True), dtype="bf16")
    main = dit.DiffusionTransformer(**network_params(cfg, False), dtype="bf16")
    if sd_ctrl is not None:
        ctrl.load_state_dict(sd_ctrl, strict=True)
    if sd_main is not None:
        main.load_state_dict(sd_main, strict=True)
    warp = dit.ControlDiffWarp(dit.OpenAIWrapper(main, False, torch.bfloat16), dit.OpenAIWrapper(ctrl, False, torch.bfloat16),
                               None, True)
    if device is not None:
        warp = warp.to(torch.bfloat16).to(device)
    return warp


def random_init_(warp: dit.ControlDiffWarp, seed: int = 0, std: float = 0.02) -> None:
    """On-device seeded random init for benchmarking (no checkpoint exists offline): N(0, std^2) everywhere,
    LayerNorm weights 1 + N(0, std^2), zero-linears randomised so the control path is not vacuous."""
    dev = next(warp.parameters()).device
    g = torch.Generator(device=dev).manual_seed(seed)
    with torch.no_grad():
        for name, p in warp.named_parameters():
            if name.endswith("pos_embedding"):
                continue
            noise = torch.randn(p.shape, generator=g, device=dev, dtype=torch.float32) * std
            if name.endswith("weight") and p.dim() == 1:
                noise += 1.0
            p.copy_(noise.to(p.dtype))
