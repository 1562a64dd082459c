"""B200-native drop-in for LanDiff's diffusion-stage network (reference: landiff/diffusion/dit_video_concat.py).

Same class names, constructor kwargs, `forward(x, timesteps=None, context=None, y=None, **kwargs)` signature and
state-dict keys as the reference, so switching implementation is a change of the YAML `target:` module path
(`landiff.diffusion.dit_video_concat.X` -> `landiff_b200.dit.X`), see INTEGRATION.md.  The nn.Modules here only
HOLD parameters under the reference's names; all arithmetic is done by the hand-written sm_100a kernels behind
the C-ABI (landiff_b200.ops).  bf16 only, inference only, no CPU path.

Reference map (file:line in /root/reference/landiff/diffusion/dit_video_concat.py):
  DiffusionTransformer          :670-909      ControlDiffusionTransformer :912-1027
  ControlDiffWarp               :1164-1200    AdaLNMixin.layer_forward    :540-629
  ControlOutAdaLNMixin          :1203-1238    ControlAdaLNMixin           :1241-1372
  ImagePatchEmbeddingMixin      :25-68        Basic3DPositionEmbeddingMixin :200-246
  FinalLayerMixin               :413-460      get_3d_sincos_pos_embed     :71-171
SAT-owned pieces (SelfAttention/MLP/LayerNorm/final_layernorm) follow SURVEY.md Appendix A.
"""
from __future__ import annotations

import importlib
import sys
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops
from ._C import EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_POS, EPI_GATED_RESID, EPI_NONE, EPI_QKV, EPI_UNPATCHIFY

BF16 = torch.bfloat16
F32 = torch.float32

str_to_dtype = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}

BLOCK_LAYERNORM_EPS = 1e-5  # SAT layernorm_epsilon default, passed explicitly to the layernorm factory
QK_LAYERNORM_EPS = 1e-6     # dit_video_concat.py:519-538
FINAL_NORM_EPS = 1e-6       # dit_video_concat.py:428-430


# ------------------------------------------------------------------------------------------------ plugin glue
def get_obj_from_str(string: str):
    module, cls = string.rsplit(".", 1)
    return getattr(importlib.import_module(module), cls)


def instantiate_from_config(config, **extra_kwargs):
    """Same contract as the reference's sgm/util.py:273-292."""
    if "target" not in config:
        raise KeyError("Expected key `target` to instantiate.")
    return get_obj_from_str(config["target"])(**config.get("params", dict()), **extra_kwargs)


class InferValueRegistry:
    """Process-global side channel carrying `semantic_feature` / `semantic_token` into the control network
    (reference: sgm/util.py:409-427).  When the reference package is loaded, ITS registry is used so that values
    registered by the reference's inference wrapper (dif_infer.py:161-165) are seen here."""
    _values: Dict[str, object] = {}

    @classmethod
    def register(cls, key, value):
        cls._values[key] = value

    @classmethod
    def get_value(cls, key):
        return cls._values.get(key, None)

    @classmethod
    def clear(cls):
        cls._values = {}


def _registry():
    ref = sys.modules.get("landiff.diffusion.sgm.util")
    if ref is not None and hasattr(ref, "InferValueRegistry"):
        return ref.InferValueRegistry
    return InferValueRegistry


# ------------------------------------------------------------------------------------------------ pos-emb table
def _sincos_1d(dim: int, pos: torch.Tensor) -> torch.Tensor:
    """[sin | cos] of pos * 10000^(-i/(dim/2)), float64 (dit_video_concat.py:150-171)."""
    omega = 1.0 / (10000.0 ** (torch.arange(dim // 2, dtype=torch.float64) / (dim / 2.0)))
    ang = pos.reshape(-1).to(torch.float64)[:, None] * omega[None, :]
    return torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)


def sincos_pos_embed_3d(dim: int, grid_h: int, grid_w: int, t_size: int, h_interp=1.0, w_interp=1.0, t_interp=1.0):
    """[t_size*grid_h*grid_w, dim] table: temporal sincos in the first dim/4 channels, spatial sincos (first half from
    the W coordinate, second half from the H coordinate — the reference meshgrid puts w first) in the other 3/4."""
    assert dim % 4 == 0
    d_sp, d_t = dim // 4 * 3, dim // 4
    hh = (torch.arange(grid_h, dtype=torch.float32) / h_interp)
    ww = (torch.arange(grid_w, dtype=torch.float32) / w_interp)
    w_coord = ww[None, :].expand(grid_h, grid_w)  # varies along width
    h_coord = hh[:, None].expand(grid_h, grid_w)
    spatial = torch.cat([_sincos_1d(d_sp // 2, w_coord), _sincos_1d(d_sp // 2, h_coord)], dim=1)  # [H*W, d_sp]
    tt = torch.arange(t_size, dtype=torch.float32) / t_interp
    temporal = _sincos_1d(d_t, tt)  # [T, d_t]
    full = torch.cat([temporal[:, None, :].expand(t_size, grid_h * grid_w, d_t),
                      spatial[None, :, :].expand(t_size, grid_h * grid_w, d_sp)], dim=-1)
    return full.reshape(t_size * grid_h * grid_w, dim).float()


# ------------------------------------------------------------------------------------------------ param holders
class _Holder(nn.Module):
    """Parameter container: never called."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("landiff_b200 parameter holders are not callable; use the owning model's forward")


class _Attention(_Holder):
    def __init__(self, d):
        super().__init__()
        self.query_key_value = nn.Linear(d, 3 * d, bias=True)
        self.dense = nn.Linear(d, d, bias=True)


class _MLP(_Holder):
    def __init__(self, d):
        super().__init__()
        self.dense_h_to_4h = nn.Linear(d, 4 * d, bias=True)
        self.dense_4h_to_h = nn.Linear(4 * d, d, bias=True)


class _Layer(_Holder):
    def __init__(self, d):
        super().__init__()
        self.input_layernorm = nn.LayerNorm(d, eps=BLOCK_LAYERNORM_EPS)
        self.attention = _Attention(d)
        self.post_attention_layernorm = nn.LayerNorm(d, eps=BLOCK_LAYERNORM_EPS)
        self.mlp = _MLP(d)


class _Transformer(_Holder):
    def __init__(self, d, num_layers):
        super().__init__()
        self.layers = nn.ModuleList([_Layer(d) for _ in range(num_layers)])
        self.final_layernorm = nn.LayerNorm(d, eps=BLOCK_LAYERNORM_EPS)
        self.layernorm_order = "pre"


class _PosEmbed(_Holder):
    def __init__(self, height, width, compressed_num_frames, hidden_size, text_length=0, height_interpolation=1.0,
                 width_interpolation=1.0, time_interpolation=1.0):
        super().__init__()
        self.text_length = text_length
        self.num_patches = height * width * compressed_num_frames
        self.pos_embedding = nn.Parameter(torch.zeros(1, int(text_length + self.num_patches), int(hidden_size)),
                                          requires_grad=False)
        table = sincos_pos_embed_3d(hidden_size, height, width, compressed_num_frames, height_interpolation,
                                    width_interpolation, time_interpolation)
        self.pos_embedding.data[:, -self.num_patches:].copy_(table)


class _PatchEmbed(_Holder):
    def __init__(self, in_channels, hidden_size, patch_size, bias=True, text_hidden_size=None):
        super().__init__()
        if not bias or text_hidden_size is None:
            raise NotImplementedError("landiff_b200 implements the shipped config: conv bias and text_proj present")
        self.proj = nn.Conv2d(in_channels, hidden_size, kernel_size=patch_size, stride=patch_size, bias=True)
        self.text_proj = nn.Linear(text_hidden_size, hidden_size)


class _AdaLN(_Holder):
    def __init__(self, hidden_size, num_layers, time_embed_dim, hidden_size_head, zero_linears: bool):
        super().__init__()
        self.adaLN_modulations = nn.ModuleList(
            [nn.Sequential(nn.SiLU(), nn.Linear(time_embed_dim, 12 * hidden_size)) for _ in range(num_layers)])
        self.query_layernorm_list = nn.ModuleList(
            [nn.LayerNorm(hidden_size_head, eps=QK_LAYERNORM_EPS) for _ in range(num_layers)])
        self.key_layernorm_list = nn.ModuleList(
            [nn.LayerNorm(hidden_size_head, eps=QK_LAYERNORM_EPS) for _ in range(num_layers)])
        if zero_linears:
            self.zero_linears = nn.ModuleList([nn.Linear(hidden_size, hidden_size, bias=False) for _ in range(num_layers)])
            for p in self.zero_linears.parameters():  # zero_module(), dit_video_concat.py:1209-1217
                p.detach().zero_()


class _FinalLayer(_Holder):
    def __init__(self, hidden_size, time_embed_dim, patch_size, out_channels):
        super().__init__()
        self.norm_final = nn.LayerNorm(hidden_size, eps=FINAL_NORM_EPS)
        self.linear = nn.Linear(hidden_size, patch_size * patch_size * out_channels, bias=True)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(time_embed_dim, 2 * hidden_size, bias=True))


def _cfg_params(modules: dict, key: str) -> dict:
    if key not in modules:
        raise KeyError(f"modules['{key}'] missing")
    return dict(modules[key].get("params", {}) or {})


def _cfg_target(modules: dict, key: str) -> str:
    return modules[key]["target"].rsplit(".", 1)[-1]


class SequenceShard:
    """Token-axis shard [start, start+count) of the text+image sequence owned by this rank (ring sequence
    parallelism).  The default is the whole sequence."""

    def __init__(self, start: int, count: int, total: int):
        self.start, self.count, self.total = start, count, total

    @property
    def whole(self):
        return self.start == 0 and self.count == self.total


# ------------------------------------------------------------------------------------------------ the network
class DiffusionTransformer(nn.Module):
    """Main DiT (reference :670-909).  Constructor kwargs as in the YAML `network_config.params`."""

    _is_control = False

    def __init__(self, transformer_args=None, num_frames=49, time_compressed_rate=4, latent_width=90, latent_height=60,
                 patch_size=2, in_channels=16, out_channels=16, hidden_size=1920, num_layers=30, num_attention_heads=30,
                 elementwise_affine=True, time_embed_dim=None, num_classes=None, modules={}, input_time="adaln",
                 adm_in_channels=None, parallel_output=True, height_interpolation=1.0, width_interpolation=1.0,
                 time_interpolation=1.0, use_SwiGLU=False, use_RMSNorm=False, zero_init_y_embed=False, **kwargs):
        super().__init__()
        if use_SwiGLU or use_RMSNorm or num_classes is not None or input_time != "adaln" or not elementwise_affine:
            raise NotImplementedError("landiff_b200 implements the shipped 2B configuration only "
                                      "(GELU-tanh MLP, LayerNorm, adaLN time input, no class conditioning)")
        if patch_size != 2:
            raise NotImplementedError("patch_size must be 2")
        order = getattr(transformer_args, "layernorm_order", None) if not isinstance(transformer_args, dict) \
            else transformer_args.get("layernorm_order")
        if order not in (None, "pre"):
            raise NotImplementedError("only layernorm_order='pre' is implemented")
        if hidden_size % num_attention_heads or hidden_size // num_attention_heads != 64:
            raise NotImplementedError("the attention kernel is specialised for head_dim 64")
        dt = kwargs.pop("dtype", "bf16")
        self.dtype = str_to_dtype[dt] if isinstance(dt, str) else dt
        self.latent_width, self.latent_height, self.patch_size = latent_width, latent_height, patch_size
        self.num_frames, self.time_compressed_rate = num_frames, time_compressed_rate
        self.in_channels, self.out_channels = in_channels, out_channels
        self.hidden_size = self.model_channels = hidden_size
        self.time_embed_dim = time_embed_dim if time_embed_dim is not None else hidden_size
        self.num_classes = None
        self.num_layers, self.num_attention_heads = num_layers, num_attention_heads
        self.spatial_length = latent_width * latent_height // patch_size ** 2
        self.compressed_num_frames = (num_frames - 1) // time_compressed_rate + 1
        self._build_modules(modules)
        self._ws = {}
        self.attn_variant = 0
        self.shard: Optional[SequenceShard] = None   # explicit token shard (tests); normally derived from sp_layout
        self.sp_layout = None                        # set by landiff_b200.parallel.attach for ring SP
        self.ring = None                             # set by landiff_b200.parallel.attach

    # -- construction -------------------------------------------------------------------------------------
    def _build_modules(self, modules):
        d, te = self.hidden_size, self.time_embed_dim
        self.time_embed = nn.Sequential(nn.Linear(d, te), nn.SiLU(), nn.Linear(te, te))
        self.transformer = _Transformer(d, self.num_layers)
        pe = _cfg_params(modules, "pos_embed_config")
        if _cfg_target(modules, "pos_embed_config") != "Basic3DPositionEmbeddingMixin":
            raise NotImplementedError("only Basic3DPositionEmbeddingMixin (additive 3D sincos) is implemented")
        self.text_length = int(pe.get("text_length", 0))
        pa = _cfg_params(modules, "patch_embed_config")
        ad = _cfg_params(modules, "adaln_layer_config")
        if not ad.get("qk_ln", True):
            raise NotImplementedError("qk_ln=False is not implemented")
        if ad.get("use_semantic_injection_adaln", False):
            raise NotImplementedError("use_semantic_injection_adaln is not implemented (undefined in the reference too)")
        self.control_layers = int(ad.get("control_layers", 15))
        zero_linears = self._is_control and ad.get("use_zero_linears", True)
        self.use_zero_linears = bool(zero_linears)
        self.mixins = nn.ModuleDict()
        self.mixins["pos_embed"] = _PosEmbed(self.latent_height // 2, self.latent_width // 2, self.compressed_num_frames, d,
                                             **pe)
        self.mixins["patch_embed"] = _PatchEmbed(self.in_channels, d, self.patch_size, **pa)
        self.mixins["adaln_layer"] = _AdaLN(d, self.num_layers, te, d // self.num_attention_heads, zero_linears)
        if _cfg_target(modules, "final_layer_config") == "FinalLayerMixin":
            self.mixins["final_layer"] = _FinalLayer(d, te, self.patch_size, self.out_channels)
            self.has_final = True
        else:  # EmptyFinalLayerMixin (control net, :1375-1387)
            self.mixins["final_layer"] = _Holder()
            self.has_final = False

    # -- workspace ------------------------------------------------------------------------------------------
    def _workspace(self, B, T, H, W, device):
        n_img_total = T * (H // 2) * (W // 2)
        n_total = self.text_length + n_img_total
        if self.shard is not None:
            shard = self.shard
        elif self.sp_layout is not None and self.sp_layout.sp_size > 1:
            from .parallel import shard_bounds

            start, count = shard_bounds(n_total, self.sp_layout.sp_size, self.sp_layout.sp_rank)
            shard = SequenceShard(start, count, n_total)
        else:
            shard = SequenceShard(0, n_total, n_total)
        if shard.total != n_total:
            raise ValueError(f"sequence shard built for {shard.total} tokens, input has {n_total}")
        key = (B, T, H, W, shard.start, shard.count, str(device))
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        d, L, R = self.hidden_size, self.num_layers, shard.count
        nh = self.num_attention_heads
        e = lambda *s, dt=BF16: torch.empty(*s, dtype=dt, device=device)
        # the main net keeps its 30-layer residual stream in fp32 (bf16 rounding of the stream after every residual add
        # is the dominant error of the eager reference, ~1e-2 rel-L2 after 45 layers); the control net restarts its
        # stream from a bf16 GEMM output (the zero-linear) every layer, so it stays bf16 like the reference.
        stream_dt = BF16 if self._is_control else F32
        ws = dict(shard=shard, n_total=n_total, n_img_total=n_img_total,
                  hidden=e(B, R, d, dt=stream_dt), ln=e(B, R, d), attn=e(B, R, d), h4=e(B, R, 4 * d),
                  q=e(B, nh, R, 64), kv=e(2, B, nh, R, 64),
                  temb=e(B, d, dt=F32), e1=e(B, self.time_embed_dim, dt=F32), emb=e(B, self.time_embed_dim, dt=F32),
                  mod=e(L, B, 12 * d, dt=F32), tvec=e(B, dt=F32))
        ws["k"], ws["v"] = ws["kv"][0], ws["kv"][1]  # one contiguous K|V buffer: a ring hop is a single send
        txt_lo, txt_hi = max(shard.start, 0), min(shard.start + R, self.text_length)
        ws["n_txt"] = max(txt_hi - txt_lo, 0)                      # text rows in this shard (always a prefix)
        ws["n_img"] = R - ws["n_txt"]
        ws["g0"] = max(shard.start - self.text_length, 0)           # first image token of the shard
        ws["cols"] = e(B * max(ws["n_img"], 1), self.in_channels * 4)
        if self._is_control:
            ws["ctrl"] = e(L, B, R, d)
        if self.has_final:
            ws["fmod"] = e(B, 2 * d, dt=F32)
            ws["fin"] = e(B * max(ws["n_img"], 1), d)
            ws["out"] = e(B, T, self.out_channels, H, W)
            # token-major output for the peer-copy exchange (landiff_b200/parallel.py OutputGather): every rank's block has
            # the size of the largest shard so that one fixed-size copy moves it
            ws["tok_rows"] = (n_total + max(self.sp_layout.sp_size, 1) - 1) // max(self.sp_layout.sp_size, 1) if self.sp_layout else R
            ws["tok_out"] = e(B, ws["tok_rows"], self.out_channels * 4)
        self._ws[key] = ws
        return ws

    def _check_ready(self, x):
        if not x.is_cuda:
            raise RuntimeError("landiff_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        p = self.transformer.layers[0].attention.dense.weight
        if p.dtype != BF16 or not p.is_cuda:
            raise RuntimeError("landiff_b200 modules must be moved to CUDA in bf16 (`model.to(torch.bfloat16).cuda()`)")

    # -- pieces ---------------------------------------------------------------------------------------------
    def _time_and_modulation(self, ws, timesteps):
        B = ws["tvec"].numel()
        ws["tvec"].copy_(timesteps.reshape(-1).to(F32)[:B])
        # timestep_embedding computed in fp32 then cast to the model dtype (util.py:207-233, :888-891)
        ops.timestep_embedding(ws["tvec"], self.model_channels, round_bf16=True, out=ws["temb"])
        te = self.time_embed
        ops.small_linear(ws["temb"], te[0].weight, te[0].bias, act_out=1, out=ws["e1"])
        ops.small_linear(ws["e1"], te[2].weight, te[2].bias, out=ws["emb"])
        # adaLN_modulation(emb) = Linear(SiLU(emb)) for every layer (:555-568): ONE batched GEMV over the layers' weight
        # matrices (a device table of their addresses, rebuilt only if a parameter's storage moves)
        ada = self.mixins["adaln_layer"].adaLN_modulations
        key = tuple(ada[i][1].weight.data_ptr() for i in range(self.num_layers))
        if ws.get("ada_key") != key:
            ws["ada_w"] = ops.pointer_table([ada[i][1].weight for i in range(self.num_layers)], ws["mod"].device)
            ws["ada_b"] = ops.pointer_table([ada[i][1].bias for i in range(self.num_layers)], ws["mod"].device)
            ws["ada_key"] = key
        ops.small_linear_batched(ws["emb"], ws["ada_w"], ws["ada_b"], self.num_layers, 12 * self.hidden_size, act_in=1,
                                 out=ws["mod"])

    def _embed(self, ws, x, context, sem):
        B, T, C, H, W = x.shape
        d, TL = self.hidden_size, self.text_length
        shard, R = ws["shard"], ws["shard"].count
        hidden2d = ws["hidden"].view(B * R, d)
        pe = self.mixins["patch_embed"]
        pos = self.mixins["pos_embed"].pos_embedding[0]
        if pos.shape[0] < ws["n_total"]:
            raise ValueError("input has more tokens than the position-embedding table was built for")
        if ws["n_img"] > 0:
            ops.patchify(x, sem, g0=ws["g0"], n=ws["n_img"], out=ws["cols"])
            ops.gemm(ws["cols"], pe.proj.weight.view(d, -1), epilogue=EPI_BIAS_POS, bias=pe.proj.bias, out=hidden2d,
                     rows_per_batch=ws["n_img"], out_rows_per_batch=R, out_row_offset=ws["n_txt"],
                     tok_offset=TL + ws["g0"], text_len=TL, pos=pos)
        if ws["n_txt"] > 0:
            ctx = context.to(BF16)
            if shard.start != 0 or ws["n_txt"] != TL:
                ctx = ctx[:, shard.start:shard.start + ws["n_txt"]]
            ctx = ctx.contiguous().view(B * ws["n_txt"], -1)
            ops.gemm(ctx, pe.text_proj.weight, epilogue=EPI_BIAS_POS, bias=pe.text_proj.bias, out=hidden2d,
                     rows_per_batch=ws["n_txt"], out_rows_per_batch=R, out_row_offset=0, tok_offset=shard.start,
                     text_len=TL, pos=pos)

    def _attention(self, ws, B, R):
        if self.ring is not None:
            self.ring.attention(ws, self.attn_variant)
        else:
            ops.attention(ws["q"], ws["k"], ws["v"], out=ws["attn"], variant=self.attn_variant)

    def _block(self, ws, i, cur, out_hidden, add2=None):
        """One AdaLN transformer block (:540-629).  Reads the residual stream from `cur`, leaves the result in
        `out_hidden` (may alias `cur`)."""
        B, R, d = cur.shape
        TL, nh = self.text_length, self.num_attention_heads
        layer = self.transformer.layers[i]
        ada = self.mixins["adaln_layer"]
        mod = ws["mod"][i].view(B, 12, d)
        mstride = 12 * d
        tok0 = ws["shard"].start
        cur2d, out2d = cur.view(B * R, d), out_hidden.view(B * R, d)
        ln2d = ws["ln"].view(B * R, d)
        # chunks: 0 shift_msa 1 scale_msa 2 gate_msa 3 shift_mlp 4 scale_mlp 5 gate_mlp, 6..11 same for text (:555-568)
        ops.layernorm_modulate(cur2d, layer.input_layernorm.weight, layer.input_layernorm.bias, BLOCK_LAYERNORM_EPS,
                               mod[:, 0], mod[:, 1], mod[:, 6], mod[:, 7], mstride, B, R, tok0, TL, out=ln2d)
        qln, kln = ada.query_layernorm_list[i], ada.key_layernorm_list[i]
        att = layer.attention
        ops.gemm(ln2d, att.query_key_value.weight, epilogue=EPI_QKV, bias=att.query_key_value.bias, rows_per_batch=R,
                 qkv=(ws["q"], ws["k"], ws["v"]), qk_ln=(qln.weight, qln.bias, kln.weight, kln.bias),
                 ln_eps=QK_LAYERNORM_EPS, heads=nh, qkv_row_offset=0)
        self._attention(ws, B, R)
        ops.gemm(ws["attn"].view(B * R, d), att.dense.weight, epilogue=EPI_GATED_RESID, bias=att.dense.bias, out=out2d,
                 rows_per_batch=R, tok_offset=tok0, text_len=TL, resid=cur2d, gate_img=mod[:, 2], gate_txt=mod[:, 8],
                 mod_batch_stride=mstride)
        ops.layernorm_modulate(out2d, layer.post_attention_layernorm.weight, layer.post_attention_layernorm.bias,
                               BLOCK_LAYERNORM_EPS, mod[:, 3], mod[:, 4], mod[:, 9], mod[:, 10], mstride, B, R, tok0, TL,
                               out=ln2d)
        mlp = layer.mlp
        ops.gemm(ln2d, mlp.dense_h_to_4h.weight, epilogue=EPI_BIAS_GELU, bias=mlp.dense_h_to_4h.bias,
                 out=ws["h4"].view(B * R, 4 * d))
        ops.gemm(ws["h4"].view(B * R, 4 * d), mlp.dense_4h_to_h.weight, epilogue=EPI_GATED_RESID,
                 bias=mlp.dense_4h_to_h.bias, out=out2d, rows_per_batch=R, tok_offset=tok0, text_len=TL, resid=out2d,
                 add2=None if add2 is None else add2.view(B * R, d), gate_img=mod[:, 5], gate_txt=mod[:, 11],
                 mod_batch_stride=mstride)

    def _final(self, ws, B, T, H, W):
        d, TL = self.hidden_size, self.text_length
        fl = self.mixins["final_layer"]
        R = ws["shard"].count
        ops.small_linear(ws["emb"], fl.adaLN_modulation[1].weight, fl.adaLN_modulation[1].bias, act_in=1, out=ws["fmod"])
        fln = self.transformer.final_layernorm
        ops.final_norm_modulate(ws["hidden"].view(B * R, d), fln.weight, fln.bias, BLOCK_LAYERNORM_EPS,
                                fl.norm_final.weight, fl.norm_final.bias, FINAL_NORM_EPS, ws["fmod"][:, :d],
                                ws["fmod"][:, d:], 2 * d, B, R, ws["shard"].start, TL, out=ws["fin"])
        if getattr(self, "token_major_out", False):
            # sequence / CFG parallel: the (row, token shard) block stays token-major [B, tok_rows, 64] (rows [0, n_img) valid);
            # parallel.OutputGather ships it to every rank and unpatchifies there
            ops.gemm(ws["fin"], fl.linear.weight, epilogue=EPI_BIAS, bias=fl.linear.bias,
                     out=ws["tok_out"].view(B * ws["tok_rows"], -1), rows_per_batch=ws["n_img"], out_rows_per_batch=ws["tok_rows"])
            return ws["tok_out"]
        ops.gemm(ws["fin"], fl.linear.weight, epilogue=EPI_UNPATCHIFY, bias=fl.linear.bias, out=ws["out"],
                 rows_per_batch=ws["n_img"], tok_offset=TL + ws["g0"], text_len=TL,
                 patch_grid=(T, H // 2, W // 2, self.out_channels))
        return ws["out"]

    def _prepare(self, x, timesteps, context):
        self._check_ready(x)
        if x.dim() != 5 or x.shape[2] != self.in_channels:
            raise ValueError(f"x must be [B,T,{self.in_channels},H,W], got {tuple(x.shape)}")
        if x.dtype not in (F32, BF16):
            raise TypeError("x must be fp32 or bf16")
        if context is None or context.dim() != 3 or context.shape[1] != self.text_length:
            raise ValueError(f"context must be [B,{self.text_length},text_hidden]")
        if x.shape[3] % 2 or x.shape[4] % 2:
            raise ValueError("latent height/width must be divisible by the patch size 2")
        B, T, C, H, W = x.shape
        ws = self._workspace(B, T, H, W, x.device)
        return ws, x.contiguous()

    # -- public forward ---------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, timesteps=None, context=None, y=None, **kwargs):
        assert y is None, "must specify y if and only if the model is class-conditional"  # :885-887
        ws, x = self._prepare(x, timesteps, context)
        B, T, C, H, W = x.shape
        ctrl = kwargs.get("control_layers_output", None)
        if ctrl is not None:
            assert len(ctrl) == self.control_layers, f"{len(ctrl)} != {self.control_layers}"  # :1354-1356
        self._time_and_modulation(ws, timesteps)
        self._embed(ws, x, context, None)
        hidden = ws["hidden"]
        for i in range(self.num_layers):
            add2 = None
            if ctrl is not None and i < len(ctrl):
                c = ctrl[i]
                add2 = c["hidden_states"] if isinstance(c, dict) else c
                if tuple(add2.shape) != tuple(hidden.shape) or add2.dtype != BF16:
                    raise ValueError("control_layers_output[i] must be a bf16 [B, N, d] tensor (dict branch, :1364-1370)")
            self._block(ws, i, hidden, hidden, add2=add2)
        return self._final(ws, B, T, H, W)


class ControlDiffusionTransformer(DiffusionTransformer):
    """Control branch (reference :912-1027): x + semantic feature, N blocks each followed by a d x d zero-linear;
    returns the per-layer projected hidden states as a list of {"hidden_states": [B, N, d]} dicts."""

    _is_control = True

    def __init__(self, *args, **kwargs):
        if kwargs.pop("use_semantic_injection_adaln", False):
            raise NotImplementedError("use_semantic_injection_adaln is not implemented")
        kwargs.pop("uncertainty_sampling_mode", None)
        kwargs.pop("semantic_video_frames", None)
        modules = kwargs.get("modules", {})
        super().__init__(*args, **kwargs)
        # the semantic conditioner is whatever `semantic_condition_config` names — the reference SemanticCond, or the
        # drop-in `landiff_b200.semantic.SemanticCond` (CUDA upsample path); it is instantiated
        # only if given, with the keyword-only `dtype` the reference passes (dit_video_concat.py:926-928, condition.py:32-45)
        sc = modules.get("semantic_condition_config")
        self.semantic_conditioner = instantiate_from_config(sc, dtype=self.dtype) \
            if sc and sc.get("target") != "torch.nn.Identity" else nn.Identity()

    def _semantic_feature(self, x):
        reg = _registry()
        sem = reg.get_value("semantic_feature")
        if sem is None:
            tok = reg.get_value("semantic_token")
            if tok is None or isinstance(self.semantic_conditioner, nn.Identity):
                raise RuntimeError("no `semantic_feature` registered in InferValueRegistry and no semantic conditioner "
                                   "available to compute it (dit_video_concat.py:939-982)")
            sem = self.semantic_conditioner(indexs=tok)
            reg.register("semantic_feature", sem)
        if sem.dtype not in (BF16, F32):
            sem = sem.to(BF16)
        if sem.dtype == F32:
            sem = sem.to(BF16)  # `semantic_feature.to(self.dtype)` (:972-973, :980-982)
        if sem.shape[1:] != x.shape[1:]:
            raise ValueError(f"semantic_feature shape {tuple(sem.shape)} does not match x {tuple(x.shape)}")
        return sem.contiguous()

    @torch.no_grad()
    def forward(self, x, timesteps=None, context=None, y=None, **kwargs):
        assert y is None, "must specify y if and only if the model is class-conditional"
        ws, x = self._prepare(x, timesteps, context)
        sem = self._semantic_feature(x)
        self._time_and_modulation(ws, timesteps)
        self._embed(ws, x, context, sem)
        B, R, d = ws["hidden"].shape
        zl = self.mixins["adaln_layer"].zero_linears if self.use_zero_linears else None
        cur = ws["hidden"]
        outs: List[dict] = []
        for i in range(self.num_layers):
            self._block(ws, i, cur, ws["hidden"])
            if zl is not None:  # hidden = zero_linears[i](hidden): next layer's input AND the exported signal (:1231-1238)
                ops.gemm(ws["hidden"].view(B * R, d), zl[i].weight, epilogue=EPI_NONE, out=ws["ctrl"][i].view(B * R, d))
            else:
                ws["ctrl"][i].copy_(ws["hidden"])
            cur = ws["ctrl"][i]
            outs.append({"hidden_states": cur})
        return outs


class OpenAIWrapper(nn.Module):
    """Same contract as the reference's sgm/modules/diffusionmodules/wrappers.py:11-51."""

    def __init__(self, diffusion_model, compile_model: bool = False, dtype: torch.dtype = torch.float32):
        super().__init__()
        if compile_model:
            raise NotImplementedError("torch.compile is not used by landiff_b200")
        self.diffusion_model = diffusion_model
        self.dtype = dtype

    def forward(self, x, t, c: dict, **kwargs):
        ctx = c.get("crossattn", None)
        if "concat" in c and c["concat"] is not None and c["concat"].numel() > 0:
            raise NotImplementedError("`concat` conditioning (i2v) is not part of the shipped LanDiff config")
        return self.diffusion_model(x, timesteps=t, context=ctx, y=c.get("vector", None), **kwargs)


class ControlDiffWarp(nn.Module):
    """Control net then main net (reference :1164-1200).  `pretrain_diffusion_model_ckpt_path` may be None / missing
    for random-init benchmarking; when given it is loaded exactly like the reference does (:1176-1189)."""

    def __init__(self, main_model, control_model, pretrain_diffusion_model_ckpt_path: Optional[str] = None,
                 freeze_dit: bool = True):
        super().__init__()
        self.main_model = main_model
        self.control_model = control_model
        self.freeze_dit = freeze_dit
        if pretrain_diffusion_model_ckpt_path:
            import os

            if os.path.exists(pretrain_diffusion_model_ckpt_path):
                static = torch.load(pretrain_diffusion_model_ckpt_path, map_location="cpu")["module"]
                new_static = {k[6:]: v for k, v in static.items() if k.startswith("model.")}
                missing, unexpected = self.main_model.load_state_dict(new_static, strict=False)
                assert len(unexpected) == 0, f"unexpected_keys: {unexpected}"
                self.control_model.load_state_dict(new_static, strict=False)
            else:
                raise FileNotFoundError(pretrain_diffusion_model_ckpt_path)
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, *args, **kwargs):
        control_layers_output = self.control_model(*args, **kwargs)
        kwargs["control_layers_output"] = control_layers_output
        return self.main_model(*args, **kwargs)
