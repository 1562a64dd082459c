"""TEST INFRASTRUCTURE ONLY — never imported by the product path (landiff_b200/).

A minimal stand-in for the un-vendored third-party package `SwissArmyTransformer==0.4.12` (`sat`), pinned by the
reference at requirements.txt:7 / pyproject.toml:22 / uv.lock:2733-2735, plus empty stubs for `omegaconf`,
`pytorch_lightning` and `imageio`, so that the reference's OWN modules
(`landiff/diffusion/dit_video_concat.py`, `sgm/modules/diffusionmodules/{sampling,guiders,denoiser,
discretizer}.py`) import UNCHANGED from /root/reference in the build container and can generate golden vectors.

PARITY UNPINNED: the arithmetic SAT owns (QKV split order, SDPA scale, block-LayerNorm eps, bias presence,
final_layernorm) is restated here from the package's published behaviour; no test or
golden vector inside /root/reference pins it.  The uncertain choices are module-level parameters below.

What is restated (SAT 0.4.12 semantics):
  * BaseMixin / BaseModel: mixins in an nn.ModuleDict, hook collection in insertion order, `@non_conflict`
    chaining with `old_impl`, `add_mixin(name, mixin, reinit)`, forward -> transformer(**kw).
  * BaseTransformer.forward: word_embedding_forward hook, `+ position_embedding_forward`, per-layer
    `layer_forward` hook with `layer_id` as a 0-d tensor, optional per-layer hidden states, final_layernorm,
    `final_forward` hook; returns [final] + per-layer dicts.
  * BaseTransformerLayer: input_layernorm / attention / post_attention_layernorm / mlp.
  * SelfAttention: fused query_key_value Linear (bias) split into 3 contiguous chunks, heads-major transpose,
    attention_fn hook, dense Linear (bias).  attention_fn_default: full SDPA, scale 1/sqrt(hd).
  * MLP: dense_h_to_4h -> activation -> dense_4h_to_h (biases).
"""
from __future__ import annotations

import argparse
import math
import sys
import types
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

# ---- the SAT-owned choices this shim parameterises -------------------------------------------
BLOCK_LAYERNORM_EPS = 1e-5   # SAT passes eps=layernorm_epsilon (default 1e-5) to the `layernorm` factory
USE_FINAL_LAYERNORM = True   # BaseTransformer(use_final_layernorm=True) default

REFERENCE_ROOT = "/root/reference"

HOOK_NAMES = [
    "attention_fn", "attention_forward", "cross_attention_forward", "mlp_forward", "word_embedding_forward",
    "position_embedding_forward", "final_forward", "layer_forward", "cross_layer_embedding_forward",
    "attention_forward_default", "branch_embedding_forward", "branch_final_forward",
]


def attention_fn_default(query_layer, key_layer, value_layer, attention_mask, attention_dropout=None,
                         log_attention_weights=None, scaling_attention_score=True, **kwargs):
    """SAT transformer_defaults.attention_fn_default, SDPA branch (mask all ones => full attention)."""
    assert log_attention_weights is None
    assert scaling_attention_score
    assert bool((attention_mask > 0).all()), "only the full-attention branch is restated"
    return F.scaled_dot_product_attention(query_layer, key_layer, value_layer, attn_mask=None, dropout_p=0.0,
                                          is_causal=False)


HOOKS_DEFAULT = {name: None for name in HOOK_NAMES}
HOOKS_DEFAULT["attention_fn"] = attention_fn_default


def non_conflict(func):
    func.non_conflict = True
    return func


class BaseMixin(nn.Module):
    def __init__(self):
        super().__init__()

    def reinit(self, parent_model=None):
        pass


class LayerNorm(nn.LayerNorm):
    def __init__(self, *args, pb_relax=False, **kwargs):
        super().__init__(*args, **kwargs)
        self.pb_relax = pb_relax


class RMSNorm(nn.Module):
    def __init__(self, hidden_size, eps=1e-6, **kwargs):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.eps = eps

    def forward(self, x):
        v = x.float().pow(2).mean(-1, keepdim=True)
        return (self.weight * (x.float() * torch.rsqrt(v + self.eps))).to(x.dtype)


class ColumnParallelLinear(nn.Linear):
    def __init__(self, input_size, output_size, bias=True, gather_output=True, **kwargs):
        super().__init__(input_size, output_size, bias=bias)


class RowParallelLinear(nn.Linear):
    def __init__(self, input_size, output_size, bias=True, **kwargs):
        super().__init__(input_size, output_size, bias=bias)


class SelfAttention(nn.Module):
    def __init__(self, hidden_size, num_attention_heads, layer_id, hooks, params_dtype):
        super().__init__()
        self.hooks = hooks
        self.layer_id = layer_id
        self.num_attention_heads_per_partition = num_attention_heads
        self.hidden_size_per_attention_head = hidden_size // num_attention_heads
        self.query_key_value = ColumnParallelLinear(hidden_size, 3 * hidden_size, bias=True)
        self.dense = RowParallelLinear(hidden_size, hidden_size, bias=True)

    def _transpose_for_scores(self, t):
        new_shape = t.size()[:-1] + (self.num_attention_heads_per_partition, self.hidden_size_per_attention_head)
        return t.view(*new_shape).permute(0, 2, 1, 3)

    def forward(self, hidden_states, mask, *args, **kw_args):
        attention_fn = self.hooks.get("attention_fn") or attention_fn_default
        mixed = self.query_key_value(hidden_states)
        q, k, v = mixed.chunk(3, dim=-1)  # stride=3: three equal contiguous chunks
        q, k, v = self._transpose_for_scores(q), self._transpose_for_scores(k), self._transpose_for_scores(v)
        ctx = attention_fn(q, k, v, mask, None, **kw_args)
        ctx = ctx.permute(0, 2, 1, 3).contiguous()
        ctx = ctx.view(*ctx.size()[:-2], -1)
        return self.dense(ctx)


class MLP(nn.Module):
    def __init__(self, hidden_size, inner_hidden_size, activation_func):
        super().__init__()
        self.activation_func = activation_func
        self.dense_h_to_4h = ColumnParallelLinear(hidden_size, inner_hidden_size, bias=True)
        self.dense_4h_to_h = RowParallelLinear(inner_hidden_size, hidden_size, bias=True)

    def forward(self, hidden_states, **kw_args):
        return self.dense_4h_to_h(self.activation_func(self.dense_h_to_4h(hidden_states)))


class BaseTransformerLayer(nn.Module):
    def __init__(self, hidden_size, num_attention_heads, layer_id, layernorm, activation_func, hooks, params_dtype,
                 layernorm_order):
        super().__init__()
        self.layer_id = layer_id
        self.hooks = hooks
        self.layernorm_order = layernorm_order
        self.input_layernorm = layernorm(hidden_size, eps=BLOCK_LAYERNORM_EPS)
        self.attention = SelfAttention(hidden_size, num_attention_heads, layer_id, hooks, params_dtype)
        self.post_attention_layernorm = layernorm(hidden_size, eps=BLOCK_LAYERNORM_EPS)
        self.mlp = MLP(hidden_size, 4 * hidden_size, activation_func)


class BaseTransformer(nn.Module):
    def __init__(self, num_layers, vocab_size, hidden_size, num_attention_heads, max_sequence_length,
                 layernorm_order="pre", layernorm=LayerNorm, activation_func=None, hooks=None, params_dtype=torch.float,
                 **kwargs):
        super().__init__()
        self.hooks = dict(hooks or {})
        self.num_layers = num_layers
        self.hidden_size = hidden_size
        self.layernorm_order = layernorm_order
        self.word_embeddings = nn.Embedding(vocab_size, hidden_size)
        self.position_embeddings = nn.Embedding(max_sequence_length, hidden_size)
        act = activation_func if activation_func is not None else partial(F.gelu, approximate="tanh")
        self.layers = nn.ModuleList([
            BaseTransformerLayer(hidden_size, num_attention_heads, i, layernorm, act, self.hooks, params_dtype,
                                 layernorm_order) for i in range(num_layers)])
        self.use_final_layernorm = USE_FINAL_LAYERNORM
        if self.use_final_layernorm:
            self.final_layernorm = layernorm(hidden_size, eps=BLOCK_LAYERNORM_EPS)

    def forward(self, input_ids, position_ids, attention_mask, *, output_hidden_states=False, **kw_args):
        hooks = self.hooks
        hidden = hooks["word_embedding_forward"](input_ids, output_cross_layer={}, **kw_args)
        pe = hooks["position_embedding_forward"](position_ids, output_cross_layer={}, **kw_args)
        if pe is not None:
            hidden = hidden + pe
        outputs_per_layer = []
        for i in range(self.num_layers):
            args = dict(kw_args, layer_id=torch.tensor(i), position_ids=position_ids, output_this_layer={},
                        output_cross_layer={})
            hidden = hooks["layer_forward"](hidden, attention_mask, **args)
            out_this = {}
            if output_hidden_states:
                out_this["hidden_states"] = hidden
            outputs_per_layer.append(out_this)
        logits = self.final_layernorm(hidden) if self.use_final_layernorm else hidden
        final = hooks["final_forward"](logits, **kw_args)
        return [final] + outputs_per_layer


class BaseModel(nn.Module):
    def __init__(self, args, transformer=None, params_dtype=torch.float, **kwargs):
        super().__init__()
        self.mixins = nn.ModuleDict()
        self.collect_hooks_()
        self.transformer = BaseTransformer(
            num_layers=args.num_layers, vocab_size=args.vocab_size, hidden_size=args.hidden_size,
            num_attention_heads=args.num_attention_heads, max_sequence_length=args.max_sequence_length,
            layernorm_order=args.layernorm_order, hooks=self.hooks, params_dtype=params_dtype, **kwargs)

    def add_mixin(self, name, new_mixin, reinit=False):
        assert name not in self.mixins
        assert isinstance(new_mixin, BaseMixin)
        self.mixins[name] = new_mixin
        object.__setattr__(new_mixin, "transformer", self.transformer)  # reference to, not child of
        self.collect_hooks_()
        if reinit:
            new_mixin.reinit(self)

    def collect_hooks_(self):
        hooks, origins = {}, {}
        for name in HOOK_NAMES:
            if hasattr(self, name):  # model-level override
                hooks[name] = getattr(self, name)
                origins[name] = "model"
            for mixin_name, m in self.mixins.items():
                if hasattr(m, name):
                    fn = getattr(m, name)
                    if hasattr(fn, "non_conflict"):
                        old = hooks.get(name, HOOKS_DEFAULT.get(name))
                        hooks[name] = partial(fn, old_impl=old)
                    elif name in hooks and origins[name] != "model":
                        raise ValueError(f"hook {name} conflicts between {mixin_name} and {origins[name]}")
                    else:
                        hooks[name] = fn
                    origins[name] = mixin_name
        self.hooks = hooks
        if hasattr(self, "transformer"):
            self.transformer.hooks.clear()
            self.transformer.hooks.update(hooks)
            for layer in self.transformer.layers:
                layer.attention.hooks = self.transformer.hooks
        return hooks

    def forward(self, *args, **kwargs):
        self.transformer.hooks.clear()
        self.transformer.hooks.update(self.hooks)
        return self.transformer(*args, **kwargs)


def _print_rank0(msg, level=None, flush=False):
    pass


class _Dummy:
    """attribute sink for stub modules: any attribute is a dummy class usable as a base class / annotation"""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return _Dummy()


def _stub_module(name):
    m = types.ModuleType(name)

    def _getattr(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return type(attr, (_Dummy,), {})

    m.__getattr__ = _getattr
    m.__path__ = []
    return m


_installed = False


def install():
    """Put the shim + stubs into sys.modules and /root/reference on sys.path.  Idempotent."""
    global _installed
    if _installed:
        return
    import os

    os.environ.setdefault("LANDIFF_SKIP_INIT", "1")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    sat = mod("sat")
    sat.helpers = mod("sat.helpers", print_rank0=_print_rank0)
    sat.model = mod("sat.model", BaseModel=BaseModel)
    sat.model.base_model = mod("sat.model.base_model", BaseModel=BaseModel, non_conflict=non_conflict,
                               get_model=lambda args, cls, **kw: cls(args, **kw))
    sat.model.mixins = mod("sat.model.mixins", BaseMixin=BaseMixin)
    sat.mpu = mod("sat.mpu", get_model_parallel_world_size=lambda: 1, get_model_parallel_rank=lambda: 0,
                  get_model_parallel_group=lambda: None, get_data_parallel_world_size=lambda: 1,
                  get_model_parallel_src_rank=lambda: 0)
    sat.mpu.layers = mod("sat.mpu.layers", ColumnParallelLinear=ColumnParallelLinear, RowParallelLinear=RowParallelLinear)
    sat.ops = mod("sat.ops")
    sat.ops.layernorm = mod("sat.ops.layernorm", LayerNorm=LayerNorm, RMSNorm=RMSNorm)
    sat.transformer_defaults = mod("sat.transformer_defaults", HOOKS_DEFAULT=HOOKS_DEFAULT,
                                   attention_fn_default=attention_fn_default)
    for name in ("imageio", "pytorch_lightning", "omegaconf", "kornia", "fiddle", "vector_quantize_pytorch", "xformers",
                 "xformers.ops", "deepspeed", "wandb"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _stub_module(name)
    _installed = True


def transformer_args():
    """YAML transformer_args (configs/cogvideox_2b_control_theia_interpolate_video_vq.yaml:43-50, :119-126)."""
    return argparse.Namespace(checkpoint_activations=False, vocab_size=1, max_sequence_length=64, layernorm_order="pre",
                              skip_init=False, model_parallel_size=1, is_decoder=False)
