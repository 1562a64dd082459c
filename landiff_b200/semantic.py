"""Drop-in for the reference's semantic conditioner.

Reference: `landiff.diffusion.semantic_models.condition.SemanticCond` (condition.py:30-137) built from
`modules.semantic_condition_config` of the control network (cogvideox_2b_control_theia_interpolate_video_vq.yaml:53-80):

    semantic tokens --semantic_model (VideoVQWrap: the tokenizer's decoder, stays reference code)--> features
    [B, T, 768, H/16, W/16] --upsample_model (VQGAN-style conv Decoder, vq_gan_blocks.py:480-606)--> [B*T, 64, H/8, W/8]
    --conv_out (zero-initialised 3x3, 64 -> 16)--> semantic_feature [B, T, 16, H/8, W/8], added to the control net's latent.

Everything after the tokenizer's features is built here on the CUDA kernels of csrc/conv_kernels.cu + the tcgen05 GEMM:
same constructor kwargs, same `forward(visual, indexs, vq_origin_features, semantic_feature_before_upsample)` signature,
same state-dict names (`upsample_model.conv_in.weight`, `upsample_model.mid.block_1.norm1.weight`,
`upsample_model.up.1.upsample.conv.weight`, `conv_out.weight`, ...), so a checkpoint's `semantic_conditioner.*` tensors load
unchanged.  `semantic_model` is instantiated from its config exactly like the reference does (condition.py:47) and is only
called, never re-implemented: it is the LLM-side tokenizer, outside the diffusion hot path.

Integration: in the model YAML set
    modules.semantic_condition_config.target: landiff_b200.semantic.SemanticCond
(the nested `upsample_model_config` may keep pointing at the reference Decoder class: only its `params` are read).

Supported decoder configuration = the shipped one and its relatives: no attention blocks (`attn_resolutions: []`,
`use_mid_attention: False`), `upsample_type: pixelshuffle`, `resamp_with_conv: True`; anything else raises.
There is no CPU path: tensors must live on an sm_100 GPU.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops
from .dit import instantiate_from_config

BF16 = torch.bfloat16
GN_GROUPS = 32      # vq_gan_blocks.py:35-38
GN_EPS = 1e-6


class _Conv(nn.Module):
    """Parameter holder with nn.Conv2d's state-dict layout; caches the GEMM operand [Cout, (ky, kx, cin)]."""

    def __init__(self, cin: int, cout: int, k: int = 3):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k))
        self.bias = nn.Parameter(torch.zeros(cout))
        bound = 1.0 / (cin * k * k) ** 0.5          # nn.Conv2d's default init range
        nn.init.uniform_(self.weight, -bound, bound)
        nn.init.uniform_(self.bias, -bound, bound)
        self.k = k
        self._taps = None
        self._taps_key = None

    def taps(self) -> torch.Tensor:
        w = self.weight
        key = (w.data_ptr(), w._version, w.dtype, str(w.device))
        if self._taps_key != key:
            self._taps = ops.conv_weight_taps(w) if self.k == 3 else w.detach().reshape(w.shape[0], -1).contiguous()
            self._taps_key = key
        return self._taps


class _GroupNorm(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))


def _gn(x: torch.Tensor, norm: _GroupNorm):
    return (ops.groupnorm_stats(x, GN_GROUPS, GN_EPS), norm.weight, norm.bias, GN_GROUPS)


class ResnetBlock(nn.Module):
    """vq_gan_blocks.py:90-147 with temb_channels = 0 (the Decoder passes temb = None) and dropout 0."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.norm1 = _GroupNorm(cin)
        self.conv1 = _Conv(cin, cout)
        self.norm2 = _GroupNorm(cout)
        self.conv2 = _Conv(cout, cout)
        if cin != cout:
            self.nin_shortcut = _Conv(cin, cout, k=1)

    def forward(self, x: torch.Tensor, col=None) -> torch.Tensor:   # x channels-last [F, H, W, cin]
        h = ops.conv3x3(x, self.conv1.taps(), self.conv1.bias, gn=_gn(x, self.norm1), col=col)
        if hasattr(self, "nin_shortcut"):
            sc = ops.gemm(x.view(-1, x.shape[-1]), self.nin_shortcut.taps(), epilogue=ops.EPI_BIAS,
                          bias=self.nin_shortcut.bias).view(*x.shape[:3], -1)
        else:
            sc = x
        return ops.conv3x3(h, self.conv2.taps(), self.conv2.bias, gn=_gn(h, self.norm2), add=sc, col=col)


class _Upsample(nn.Module):
    """vq_gan_blocks.py:41-66, `pixelshuffle` flavour: PixelShuffle(2) then a 3x3 conv C/4 -> C."""

    def __init__(self, c: int):
        super().__init__()
        self.conv = _Conv(c // 4, c)

    def forward(self, x, col=None):
        return ops.conv3x3(ops.pixel_shuffle2(x), self.conv.taps(), self.conv.bias, col=col)


class Decoder(nn.Module):
    """The reference `Decoder` (vq_gan_blocks.py:480-606) for the attention-free pixelshuffle configuration; forward takes
    and returns channels-last bf16 frames."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, give_pre_end=False, use_mid_attention=True,
                 upsample_type="interpolate"):
        super().__init__()
        if use_mid_attention or len(attn_resolutions) > 0:
            raise NotImplementedError("landiff_b200.semantic.Decoder: attention blocks are not part of the shipped semantic "
                                      "conditioner (use_mid_attention: False, attn_resolutions: [])")
        if upsample_type != "pixelshuffle" or not resamp_with_conv:
            raise NotImplementedError("landiff_b200.semantic.Decoder: only upsample_type='pixelshuffle' with a conv is built")
        if give_pre_end or dropout != 0.0:
            raise NotImplementedError("landiff_b200.semantic.Decoder: give_pre_end / dropout are not supported")
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        block_in = int(ch * ch_mult[-1])
        for c in [block_in] + [int(ch * m) for m in ch_mult]:
            if c % 64:
                raise ValueError(f"landiff_b200.semantic.Decoder: channel width {c} must be a multiple of 64")
        self.conv_in = _Conv(z_channels, block_in)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(block_in, block_in)
        self.mid.block_2 = ResnetBlock(block_in, block_in)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block_out = int(ch * ch_mult[i_level])
            up = nn.Module()
            up.block = nn.ModuleList()
            up.attn = nn.ModuleList()
            for _ in range(num_res_blocks + 1):
                up.block.append(ResnetBlock(block_in, block_out))
                block_in = block_out
            if i_level != 0:
                up.upsample = _Upsample(block_in)
            self.up.insert(0, up)
        self.norm_out = _GroupNorm(block_in)
        self.conv_out = _Conv(block_in, out_ch)

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        col = None   # every width here is a multiple of 64: implicit convolution, no im2col buffer
        h = ops.conv3x3(z, self.conv_in.taps(), self.conv_in.bias, col=col)
        h = self.mid.block_1(h, col)
        h = self.mid.block_2(h, col)
        for i_level in reversed(range(self.num_resolutions)):
            for blk in self.up[i_level].block:
                h = blk(h, col)
            if i_level != 0:
                h = self.up[i_level].upsample(h, col)
        return ops.conv3x3(h, self.conv_out.taps(), self.conv_out.bias, gn=_gn(h, self.norm_out), col=col)


def pad_to_square(x: torch.Tensor, pad_values):
    """condition.py:15-27 (uint8 frames [..., C, H, W]; pads right / bottom with the per-channel fill)."""
    h, w = x.shape[-2:]
    if h == w:
        return x
    s = max(h, w)
    out = torch.empty(*x.shape[:-2], s, s, dtype=x.dtype, device=x.device)
    fill = torch.tensor(pad_values, dtype=x.dtype, device=x.device).view(*([1] * (x.dim() - 3)), -1, 1, 1)
    out[...] = fill
    out[..., :h, :w] = x
    return out


def prepare_visual(visual: torch.Tensor, pad_values) -> torch.Tensor:
    """[-1, 1] float frames [..., C, H, W] -> uint8 frames padded to a square, as the tokenizer expects them
    (condition.py:118-123: (v + 1) / 2, clamp, torchvision `to_dtype(uint8, scale=True)` = multiply by 256 - 1e-3 and
    truncate; condition.py:95-99: pad right / bottom with grey 127)."""
    v = ((visual + 1.0) / 2.0).clamp(0, 1)
    v = v.mul(255.0 + 1.0 - 1e-3).to(torch.uint8)
    return pad_to_square(v, pad_values)


class SemanticCond(nn.Module):
    """condition.py:30-137."""

    def __init__(self, *, semantic_model_config, upsample_model_config, dtype, out_dim, target_dim, dowsample_factor=16,
                 feature_type: str = "video_theia_interpolate", zero_init_conv_out: bool = True,
                 augmenter_params: Optional[dict] = None, **kwargs):
        super().__init__()
        if feature_type != "video_theia_interpolate":
            raise ValueError(f"Unknown feature type: {feature_type}")
        if target_dim != 16:
            raise NotImplementedError("landiff_b200.semantic.SemanticCond: target_dim must be 16 (the latent channel count)")
        self.semantic_model = instantiate_from_config(semantic_model_config)
        if upsample_model_config is None:
            raise NotImplementedError("landiff_b200.semantic.SemanticCond needs an upsample_model_config")
        self.upsample_model = Decoder(**upsample_model_config.get("params", {}))
        self.conv_out = _Conv(out_dim, target_dim)
        if zero_init_conv_out:                      # zero_module (landiff/utils.py), condition.py:49-52
            nn.init.zeros_(self.conv_out.weight)
            nn.init.zeros_(self.conv_out.bias)
        self.feature_type = feature_type
        self.dowsample_factor = dowsample_factor
        self.dtype = dtype
        self.pad_values = [127, 127, 127]
        for m in (self.upsample_model, self.conv_out):
            m.to(BF16)

    @property
    def device(self):
        return next(self.upsample_model.parameters()).device

    def upsample_features(self, features: torch.Tensor) -> torch.Tensor:
        """features [B, T, C, h, w] (any float dtype) -> semantic_feature [B, T, 16, 2h, 2w] bf16: `upsample_model` +
        `conv_out` (condition.py:104-110, 131-136)."""
        if features.dim() != 5:
            raise ValueError(f"semantic features must be [B, T, C, h, w], got {tuple(features.shape)}")
        if not features.is_cuda:
            raise RuntimeError("landiff_b200.semantic has no CPU path: move the features to an sm_100 GPU")
        B, T = features.shape[:2]
        f = features.reshape(B * T, *features.shape[2:])
        if f.dtype not in (BF16, torch.float32):
            f = f.float()
        z = ops.nchw_to_nhwc(f.contiguous())
        h = self.upsample_model(z)                                               # [B*T, 2h, 2w, out_dim]
        out = ops.conv3x3_to_nchw16(h, self.conv_out.weight, self.conv_out.bias)     # [B*T, 16, 2h, 2w]
        return out.view(B, T, *out.shape[1:])

    def forward(self, visual: torch.Tensor = None, indexs: torch.Tensor = None, vq_origin_features: torch.Tensor = None,
                semantic_feature_before_upsample: torch.Tensor = None) -> torch.Tensor:
        target = None
        if visual is not None:
            oh, ow = visual.shape[-2:]
            target = (oh // self.dowsample_factor, ow // self.dowsample_factor)
            visual = prepare_visual(visual, self.pad_values)
        if semantic_feature_before_upsample is None:
            features = self.semantic_model(visual, indexs)
        else:
            features = semantic_feature_before_upsample
        if target is not None:
            features = features[..., : target[0], : target[1]]
        return self.upsample_features(features)
