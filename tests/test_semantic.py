"""Semantic conditioner, upsample path: oracle pinned to the reference's own modules (CPU),
drop-in module contract (CPU), CUDA path against the reference-generated golden and against the oracle at the shipped
widths and a full-resolution frame (GPU)."""
from pathlib import Path

import pytest
import torch

from oracle import semantic_oracle as S

GOLDEN = Path(__file__).parent / "golden" / "semantic_ref.pt"
CASES = {"shipped": S.SHIPPED, "small": S.SMALL}


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


@pytest.mark.parametrize("tag", ["small", "shipped"])
def test_oracle_matches_reference_golden(tag):
    blob = torch.load(GOLDEN)[tag]
    cfg = CASES[tag]
    x = S.features_for(cfg, blob["xseed"], blob["B"], blob["T"], blob["h"], blob["w"])
    y = S.semantic_oracle(S.random_state_dict(cfg, blob["wseed"]), x, cfg)
    assert y.shape == blob["out"].shape
    assert rel(y, blob["out"]) < 1e-5


def test_drop_in_state_dict_contract():
    """Same parameter names and shapes as the reference SemanticCond (oracle.param_shapes is checked against the reference
    by the golden generator's strict load), required keyword-only `dtype`, unsupported decoder options refused."""
    from landiff_b200.semantic import SemanticCond

    for cfg in (S.SMALL, S.SHIPPED):
        m = SemanticCond(**cfg.cond_kwargs("landiff.diffusion.semantic_models.modules.vq_gan_blocks.Decoder", torch.bfloat16))
        own = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert own == S.param_shapes(cfg)
        assert all(v.dtype == torch.bfloat16 for v in m.state_dict().values())
        assert float(m.conv_out.weight.detach().abs().sum()) == 0.0 and float(m.conv_out.bias.detach().abs().sum()) == 0.0   # zero_module
        m.load_state_dict({k: v.bfloat16() for k, v in S.random_state_dict(cfg, 5).items()}, strict=True)
    with pytest.raises(TypeError):
        kw = S.SMALL.cond_kwargs("x", torch.bfloat16)
        kw.pop("dtype")
        SemanticCond(**kw)
    bad = S.SMALL.cond_kwargs("x", torch.bfloat16)
    bad["upsample_model_config"]["params"]["use_mid_attention"] = True
    with pytest.raises(NotImplementedError):
        SemanticCond(**bad)
    m = SemanticCond(**S.SMALL.cond_kwargs("x", torch.bfloat16))
    with pytest.raises(RuntimeError):   # no CPU path
        m(semantic_feature_before_upsample=torch.zeros(1, 1, S.SMALL.z_channels, 4, 4))


def test_visual_preparation_matches_torchvision():
    """The uint8 conversion and the grey square padding in front of the tokenizer == the torchvision calls the reference
    makes (condition.py:15-27, 118-123)."""
    tv = pytest.importorskip("torchvision.transforms.v2")
    from landiff_b200.semantic import prepare_visual

    g = torch.Generator().manual_seed(3)
    for shape in [(1, 3, 3, 48, 80), (2, 2, 3, 64, 40), (1, 1, 3, 32, 32)]:
        visual = torch.rand(shape, generator=g) * 2.4 - 1.2          # values beyond [-1, 1] exercise the clamp
        want = ((visual + 1.0) / 2.0).clamp(0, 1)
        want = tv.functional.to_dtype(want, dtype=torch.uint8, scale=True)
        h, w = want.shape[-2:]
        if h != w:
            pad = (0, 0, h - w, 0) if h > w else (0, 0, 0, w - h)   # torchvision order: left, top, right, bottom
            want = tv.functional.pad(want, list(pad), fill=[127, 127, 127])
        got = prepare_visual(visual, [127, 127, 127])
        assert got.dtype == torch.uint8 and torch.equal(got, want)


def test_control_net_builds_the_drop_in_conditioner():
    """`modules.semantic_condition_config.target: landiff_b200.semantic.SemanticCond` inside the control network config."""
    from landiff_b200 import dit
    from landiff_b200.factory import DiTShape, network_params
    from landiff_b200.semantic import SemanticCond

    shape = DiTShape(hidden_size=128, num_heads=2, main_layers=2, control_layers=1, time_embed_dim=64, text_hidden=64,
                     text_length=6, latent_t=2, latent_h=8, latent_w=12)
    kw = network_params(shape, control=True)
    kw["modules"]["semantic_condition_config"] = {
        "target": "landiff_b200.semantic.SemanticCond",
        "params": {k: v for k, v in S.SMALL.cond_kwargs("x", None).items() if k != "dtype"}}
    net = dit.ControlDiffusionTransformer(**kw)
    assert isinstance(net.semantic_conditioner, SemanticCond)
    assert "semantic_conditioner.upsample_model.up.1.upsample.conv.weight" in net.state_dict()


def _cuda_module(cfg, wseed):
    from landiff_b200.semantic import SemanticCond

    m = SemanticCond(**cfg.cond_kwargs("landiff.diffusion.semantic_models.modules.vq_gan_blocks.Decoder", torch.bfloat16))
    m.load_state_dict({k: v.bfloat16() for k, v in S.random_state_dict(cfg, wseed).items()}, strict=True)
    return m.cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["small", "shipped"])
def test_cuda_path_matches_reference_golden(tag):
    blob = torch.load(GOLDEN)[tag]
    cfg = CASES[tag]
    m = _cuda_module(cfg, blob["wseed"])
    x = S.features_for(cfg, blob["xseed"], blob["B"], blob["T"], blob["h"], blob["w"]).cuda()
    y = m(semantic_feature_before_upsample=x)
    assert y.dtype == torch.bfloat16 and tuple(y.shape) == tuple(blob["out"].shape)
    r = rel(y.float().cpu(), blob["out"])
    cos = torch.nn.functional.cosine_similarity(y.float().cpu().flatten(), blob["out"].flatten(), dim=0).item()
    print(f"semantic conditioner {tag}: CUDA bf16 vs reference fp32 golden rel-L2 {r:.3e} cos {cos:.6f}")
    # bf16 activations through 26 convolutions / 27 GroupNorms against an fp32 reference: same budget as the DiT step
    assert r < 1e-2 and cos > 0.9999
    # bf16 input == fp32 input (the reference casts features to self.dtype, condition.py:106-107)
    y2 = m(semantic_feature_before_upsample=x.bfloat16())
    assert torch.equal(y, y2)


@pytest.mark.gpu
def test_cuda_kernels_against_torch():
    """Each conv-stack kernel against the torch op it replaces, ragged sizes included."""
    from landiff_b200 import ops

    torch.manual_seed(0)
    dev = "cuda"
    F_, C_, H, W = 3, 128, 7, 11
    x = torch.randn(F_, C_, H, W, device=dev)
    xl = ops.nchw_to_nhwc(x)
    assert torch.equal(xl, x.bfloat16().permute(0, 2, 3, 1).contiguous())
    assert torch.equal(ops.nchw_to_nhwc(x.bfloat16()), xl)
    # GroupNorm statistics
    st = ops.groupnorm_stats(xl, 32, 1e-6)
    xg = xl.float().permute(0, 3, 1, 2).reshape(F_, 32, -1)
    assert torch.allclose(st[..., 0], xg.mean(-1), rtol=1e-4, atol=1e-5)
    assert torch.allclose(st[..., 1], (xg.var(-1, unbiased=False) + 1e-6).rsqrt(), rtol=1e-4)
    assert torch.equal(st, ops.groupnorm_stats(xl, 32, 1e-6))      # fixed reduction order: bit-reproducible
    # conv = im2col (+ GroupNorm + swish) + GEMM (+ residual)
    w = (torch.randn(64, C_, 3, 3, device=dev) / (9 * C_) ** 0.5).bfloat16()
    b = (torch.randn(64, device=dev) * 0.1).bfloat16()
    gm = (1 + 0.1 * torch.randn(C_, device=dev)).bfloat16()
    bt = (0.1 * torch.randn(C_, device=dev)).bfloat16()
    xn = xl.float().permute(0, 3, 1, 2)
    ref_plain = torch.nn.functional.conv2d(xn, w.float(), b.float(), padding=1)
    got = ops.conv3x3(xl, ops.conv_weight_taps(w), b)                       # implicit: TMA im2col-mode gather
    assert rel(got.float().permute(0, 3, 1, 2), ref_plain) < 4e-3
    got_x = ops.conv3x3(xl, ops.conv_weight_taps(w), b, implicit=False)     # explicit im2col buffer
    assert torch.equal(got, got_x)                                          # same MMA order, same operands: bit-equal
    act = torch.nn.functional.group_norm(xn, 32, gm.float(), bt.float(), eps=1e-6)
    act = act * torch.sigmoid(act)
    add = torch.randn(F_, H, W, 64, device=dev).bfloat16()
    ref_gn = torch.nn.functional.conv2d(act, w.float(), b.float(), padding=1) + add.float().permute(0, 3, 1, 2)
    got = ops.conv3x3(xl, ops.conv_weight_taps(w), b, gn=(st, gm, bt, 32), add=add)
    assert rel(got.float().permute(0, 3, 1, 2), ref_gn) < 6e-3     # bf16 rounding of the activated input + output
    got_x = ops.conv3x3(xl, ops.conv_weight_taps(w), b, gn=(st, gm, bt, 32), add=add, implicit=False,
                        max_col_bytes=H * W * 9 * C_ * 2)           # one frame per chunk
    assert torch.equal(got, got_x)
    assert rel(ops.groupnorm_apply(xl, st, gm, bt, 32, True).float(), act.permute(0, 2, 3, 1)) < 4e-3
    # a multi-tile, multi-frame implicit convolution whose 128-position tiles straddle rows and frames
    xb = torch.randn(5, 13, 17, 192, device=dev).bfloat16()
    wb = (torch.randn(128, 192, 3, 3, device=dev) / (9 * 192) ** 0.5).bfloat16()
    ref_b = torch.nn.functional.conv2d(xb.float().permute(0, 3, 1, 2), wb.float(), None, padding=1)
    got_b = ops.conv3x3(xb, ops.conv_weight_taps(wb), None)
    assert rel(got_b.float().permute(0, 3, 1, 2), ref_b) < 4e-3
    assert torch.equal(got_b, ops.conv3x3(xb, ops.conv_weight_taps(wb), None, implicit=False))
    # pixel shuffle
    ps = ops.pixel_shuffle2(xl)
    assert torch.equal(ps.permute(0, 3, 1, 2), torch.nn.functional.pixel_shuffle(xl.permute(0, 3, 1, 2), 2))
    # direct 16-channel conv to NCHW
    w16 = (torch.randn(16, 64, 3, 3, device=dev) / 24.0).bfloat16()
    b16 = (torch.randn(16, device=dev) * 0.1).bfloat16()
    x64 = torch.randn(F_, H, W, 64, device=dev).bfloat16()
    ref16 = torch.nn.functional.conv2d(x64.float().permute(0, 3, 1, 2), w16.float(), b16.float(), padding=1)
    got16 = ops.conv3x3_to_nchw16(x64, w16, b16)
    assert tuple(got16.shape) == (F_, 16, H, W) and rel(got16.float(), ref16) < 4e-3
    assert float(ops.conv3x3_to_nchw16(x64, torch.zeros_like(w16), None).float().abs().sum()) == 0.0   # zero_module init


@pytest.mark.gpu
def test_cuda_path_full_resolution_frame_against_oracle():
    """Shipped widths on 2 frames of the real 30 x 45 feature grid (480 x 720 / 16) -> 60 x 90, vs the fp32 oracle on CPU;
    also the registered-token route of the control net: forward(indexs=...) calls the (stub) semantic_model."""
    cfg = S.SHIPPED
    m = _cuda_module(cfg, 31)
    x = S.features_for(cfg, 32, 1, 2, 30, 45)
    want = S.semantic_oracle(S.random_state_dict(cfg, 31), x, cfg)
    y = m(semantic_feature_before_upsample=x.cuda())
    assert tuple(y.shape) == (1, 2, 16, 60, 90)
    r = rel(y.float().cpu(), want)
    print(f"semantic conditioner 2 x 30x45 -> 60x90: CUDA vs fp32 oracle rel-L2 {r:.3e}")
    assert r < 1e-2

    class Tokenizer(torch.nn.Module):           # stands in for VideoVQWrap: (visual, indexs) -> features
        def forward(self, visual, indexs):
            assert visual is None
            return x.cuda()[:, : indexs.shape[1]]

    m.semantic_model = Tokenizer()
    y2 = m(indexs=torch.zeros(1, 2, dtype=torch.long, device="cuda"))
    assert torch.equal(y, y2)
