"""CPU: the drop-in modules mirror the reference's plugin surface — constructor kwargs, state-dict keys/shapes,
position-embedding table, registry side channel — without running any compute (that needs the GPU)."""
import copy

import pytest
import torch

from conftest import needs_reference
from landiff_b200 import dit
from landiff_b200.factory import TINY, build_warp, network_params
from oracle import dit_oracle as O


def test_shapes_agree_between_factory_and_oracle():
    for a, b in ((TINY, O.TINY),):
        for k in ("hidden_size", "num_heads", "main_layers", "control_layers", "time_embed_dim", "text_hidden",
                  "text_length", "latent_t", "latent_h", "latent_w", "in_channels", "interp"):
            assert getattr(a, k) == getattr(b, k), k


def test_state_dict_keys_and_shapes_match_contract():
    warp = build_warp(TINY)
    for model, control in ((warp.control_model.diffusion_model, True), (warp.main_model.diffusion_model, False)):
        sd = model.state_dict()
        want = O.param_shapes(O.TINY, control)
        assert set(sd.keys()) == set(want.keys())
        for k, shp in want.items():
            assert tuple(sd[k].shape) == tuple(shp), k


def test_pos_embedding_table_matches_oracle():
    cfg = O.TINY
    warp = build_warp(TINY)
    pe = warp.main_model.diffusion_model.mixins["pos_embed"].pos_embedding
    tab = O.pos_embed_3d(cfg.hidden_size, cfg.latent_h // 2, cfg.latent_w // 2, cfg.latent_t, cfg.interp, cfg.interp)
    assert torch.equal(pe[0, cfg.text_length:], tab)
    assert torch.count_nonzero(pe[0, :cfg.text_length]) == 0


def test_zero_linears_start_at_zero_like_reference():
    warp = build_warp(TINY)
    for p in warp.control_model.diffusion_model.mixins["adaln_layer"].zero_linears.parameters():
        assert torch.count_nonzero(p) == 0


def test_unsupported_options_fail_loudly():
    p = network_params(TINY, control=False)
    with pytest.raises(NotImplementedError):
        dit.DiffusionTransformer(**{**p, "use_SwiGLU": True})
    with pytest.raises(NotImplementedError):
        dit.DiffusionTransformer(**{**p, "num_attention_heads": 4})  # head_dim 32
    q = copy.deepcopy(p)
    q["modules"]["pos_embed_config"]["target"] = "landiff.diffusion.dit_video_concat.Rotary3DPositionEmbeddingMixin"
    with pytest.raises(NotImplementedError):
        dit.DiffusionTransformer(**q)


def test_cpu_forward_refuses():
    warp = build_warp(TINY)
    x = torch.zeros(2, 2, 16, 8, 12)
    with pytest.raises(RuntimeError, match="no CPU path"):
        warp(x, torch.zeros(2), {"crossattn": torch.zeros(2, 6, 64)})


def test_registry_side_channel():
    dit.InferValueRegistry.clear()
    assert dit.InferValueRegistry.get_value("semantic_feature") is None
    dit.InferValueRegistry.register("semantic_feature", 3)
    assert dit.InferValueRegistry.get_value("semantic_feature") == 3
    dit.InferValueRegistry.clear()
    assert dit.InferValueRegistry.get_value("semantic_feature") is None


@needs_reference
def test_state_dict_interchangeable_with_reference_modules():
    """Keys AND shapes equal the reference modules' own state_dict (built on the SAT shim); both directions load."""
    from oracle import ref_build as rb

    ctrl_ref, main_ref = rb.build_reference(rb.TINY, seed=3)
    warp = build_warp(TINY)
    ours_c, ours_m = warp.control_model.diffusion_model, warp.main_model.diffusion_model
    for ours, ref in ((ours_c, ctrl_ref), (ours_m, main_ref)):
        a, b = ours.state_dict(), ref.state_dict()
        assert set(a.keys()) == set(b.keys())
        for k in a:
            assert a[k].shape == b[k].shape, k
        ours.load_state_dict(b, strict=True)
        ref.load_state_dict(ours.state_dict(), strict=True)
    mine = build_warp(TINY).main_model.diffusion_model.mixins["pos_embed"].pos_embedding
    assert torch.equal(mine, main_ref.mixins.pos_embed.pos_embedding)  # bit-exact fp32 sincos table


@needs_reference
def test_registry_uses_reference_registry_when_loaded():
    from oracle import sat_shim

    sat_shim.install()
    from landiff.diffusion.sgm.util import InferValueRegistry as RefReg

    assert dit._registry() is RefReg
