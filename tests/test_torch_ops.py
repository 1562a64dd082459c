"""`torch.ops.landiff_b200.*`: every compute entry point of the C-ABI is registered as a PyTorch custom op at import
(CPU check) and each op returns exactly what the direct ctypes path returns (GPU check, bit-equal: same kernel)."""
import pytest
import torch

import landiff_b200  # noqa: F401  (registers the ops)
from landiff_b200 import ops
from landiff_b200._C import EPI_BIAS_GELU, EPI_BIAS_POS, EPI_GATED_RESID, EPI_QKV, EPI_UNPATCHIFY

NS = torch.ops.landiff_b200


def test_every_compute_entry_point_is_a_registered_op():
    assert len(ops.TORCH_OPS) == 24
    for name in ops.TORCH_OPS:
        op = getattr(NS, name)
        assert "landiff_b200::" + name in str(op.default._schema)
    # CPU tensors are refused by dispatch (no CPU kernel is registered: there is no CPU fallback)
    with pytest.raises((NotImplementedError, RuntimeError)):
        NS.timestep_embedding(torch.zeros(2), 64)


@pytest.mark.gpu
def test_ops_match_the_ctypes_path():
    dev = "cuda"
    torch.manual_seed(0)
    B, R, TL, D, H = 2, 200, 30, 384, 6
    M = B * R
    a = (torch.randn(M, D, device=dev) * 0.5).bfloat16()
    w = (torch.randn(D, D, device=dev) * 0.05).bfloat16()
    bias = (torch.randn(D, device=dev) * 0.1).bfloat16()
    mod = torch.randn(B, 12, D, device=dev) * 0.5
    eq = lambda x, y: torch.equal(x.float(), y.float())
    assert eq(NS.linear(a, w, bias, EPI_BIAS_GELU), ops.gemm(a, w, epilogue=EPI_BIAS_GELU, bias=bias))
    resid = torch.randn(M, D, device=dev)
    add2 = torch.randn(M, D, device=dev).bfloat16()
    got = NS.linear_gated_residual(a, w, bias, resid, mod[:, 2], mod[:, 8], add2, R, 0, TL, 12 * D)
    ref = ops.gemm(a, w, epilogue=EPI_GATED_RESID, bias=bias, out=torch.empty_like(resid), rows_per_batch=R, text_len=TL,
                   resid=resid, add2=add2, gate_img=mod[:, 2], gate_txt=mod[:, 8], mod_batch_stride=12 * D)
    assert got.dtype == torch.float32 and eq(got, ref)
    wq = (torch.randn(3 * D, D, device=dev) * 0.05).bfloat16()
    bq = (torch.randn(3 * D, device=dev) * 0.1).bfloat16()
    ln = [(1 + 0.1 * torch.randn(64, device=dev)).bfloat16(), (0.1 * torch.randn(64, device=dev)).bfloat16()] * 2
    q, k, v = NS.linear_qkv(a, wq, bq, *ln, H, R, 1e-6)
    q2, k2, v2 = [torch.empty_like(q) for _ in range(3)]
    ops.gemm(a, wq, epilogue=EPI_QKV, bias=bq, rows_per_batch=R, qkv=(q2, k2, v2), qk_ln=ln, ln_eps=1e-6, heads=H)
    assert eq(q, q2) and eq(k, k2) and eq(v, v2)
    o = NS.attention(q, k, v)
    assert eq(o, ops.attention(q, k, v))
    o2, lse, of = NS.attention_lse(q, k, v)
    assert eq(o2, o) and lse.shape == (B * H, R) and of.shape == (B * H, R, 64)
    oa, la = of.clone(), lse.clone()
    out_b = torch.empty_like(o)
    NS.attention_merge(oa, la, of, lse, out_b, B, H, R)     # merging a result with itself leaves O, adds 1 to log2-LSE
    torch.cuda.synchronize()
    assert torch.allclose(oa, of, atol=1e-6) and torch.allclose(la, lse + 1.0, atol=1e-5)
    pos = torch.randn(TL + R, D, device=dev).bfloat16()
    h1 = torch.zeros(B, TL + R, D, device=dev, dtype=torch.bfloat16)
    h2 = torch.zeros_like(h1)
    NS.linear_bias_pos(a, w, bias, pos, h1.view(-1, D), R, TL + R, TL, TL, TL)
    ops.gemm(a, w, epilogue=EPI_BIAS_POS, bias=bias, out=h2.view(-1, D), rows_per_batch=R, out_rows_per_batch=TL + R,
             out_row_offset=TL, tok_offset=TL, text_len=TL, pos=pos)
    assert eq(h1, h2)
    T, Hp, Wp, Cc = 2, 4, 5, 16
    a2 = (torch.randn(B * T * Hp * Wp, D, device=dev) * 0.5).bfloat16()
    w2 = (torch.randn(64, D, device=dev) * 0.05).bfloat16()
    b2 = (torch.randn(64, device=dev) * 0.1).bfloat16()
    u1 = torch.zeros(B, T, Cc, 2 * Hp, 2 * Wp, device=dev, dtype=torch.bfloat16)
    u2 = torch.zeros_like(u1)
    NS.linear_unpatchify(a2, w2, b2, u1, T * Hp * Wp, TL, TL, T, Hp, Wp, Cc)
    ops.gemm(a2, w2, epilogue=EPI_UNPATCHIFY, bias=b2, out=u2, rows_per_batch=T * Hp * Wp, tok_offset=TL, text_len=TL,
             patch_grid=(T, Hp, Wp, Cc))
    assert eq(u1, u2) and float(u1.float().abs().sum()) > 0
    lw = (1 + 0.1 * torch.randn(D, device=dev)).bfloat16()
    lb = (0.1 * torch.randn(D, device=dev)).bfloat16()
    args = (a, lw, lb, 1e-5, mod[:, 0], mod[:, 1], mod[:, 6], mod[:, 7], 12 * D, B, R, 0, TL)
    assert eq(NS.layernorm_modulate(*args), ops.layernorm_modulate(*args))
    fm = torch.randn(B, 2, D, device=dev)
    fargs = (a, lw, lb, 1e-5, lw, lb, 1e-6, fm[:, 0], fm[:, 1], 2 * D, B, R, 0, TL)
    assert eq(NS.final_norm_modulate(*fargs), ops.final_norm_modulate(*fargs))
    x = torch.randn(B, T, Cc, 2 * Hp, 2 * Wp, device=dev)
    sem = (torch.randn(1, T, Cc, 2 * Hp, 2 * Wp, device=dev) * 0.1).bfloat16()
    assert eq(NS.patchify(x, sem), ops.patchify(x, sem))
    assert eq(NS.patchify(x, None, 3, 17), ops.patchify(x, None, g0=3, n=17))
    xe = torch.randn(B, 64, device=dev)
    wl = (torch.randn(96, 64, device=dev) * 0.05).bfloat16()
    assert eq(NS.small_linear(xe, wl, None, 1, 0, True), ops.small_linear(xe, wl, None, act_in=1))
    # batched GEMV (all adaLN projections of a network in one launch) == the per-layer launches, bit for bit
    wls = [(torch.randn(96, 64, device=dev) * 0.05).bfloat16() for _ in range(5)]
    bls = [(torch.randn(96, device=dev) * 0.1).bfloat16() for _ in range(5)]
    yb = NS.small_linear_batched(xe, wls, bls, 1, True)
    assert yb.shape == (5, B, 96)
    for l in range(5):
        assert eq(yb[l], ops.small_linear(xe, wls[l], bls[l], act_in=1)), l
    t = torch.tensor([999.0, 19.0], device=dev)
    assert eq(NS.timestep_embedding(t, 128), ops.timestep_embedding(t, 128))
    xl, old, eps = [torch.randn(1, 2, 4, 6, 8, device=dev) for _ in range(3)]
    nu, nc = torch.randn_like(xl).bfloat16(), torch.randn_like(xl).bfloat16()
    kw = dict(c_skip=0.3, c_out=-0.95, cfg=4.5, m1=0.9, m2=-0.2, m3=1.7, m4=0.7, mn=0.1)
    g1 = NS.sampler_update(xl, nu, nc, old, eps, *kw.values(), 1)
    g2 = ops.sampler_update(xl, nu, nc, old, eps, mode=1, **kw)
    assert eq(g1[0], g2[0]) and eq(g1[1], g2[1])
    # the fp32-row flavour used by the reference-compatible sampler entry
    g3 = NS.sampler_update(xl, nu.float(), nc.float(), old, eps, *kw.values(), 1)
    assert eq(g3[0], g2[0])
    # conv-stack ops of the semantic conditioner (section 8 row f2)
    xc = torch.randn(2, 64, 5, 6, device=dev)
    xl = NS.nchw_to_nhwc(xc)
    assert eq(xl, ops.nchw_to_nhwc(xc))
    st = NS.groupnorm_stats(xl, 32, 1e-6)
    assert eq(st, ops.groupnorm_stats(xl, 32, 1e-6))
    gmm, btt = (1 + 0.1 * torch.randn(64, device=dev)).bfloat16(), (0.1 * torch.randn(64, device=dev)).bfloat16()
    assert eq(NS.im2col3x3(xl, st, gmm, btt, 32, True), ops.im2col3x3(xl, gn=(st, gmm, btt, 32)))
    assert eq(NS.im2col3x3(xl, None, None, None), ops.im2col3x3(xl))
    assert eq(NS.groupnorm_apply(xl, st, gmm, btt, 32, True), ops.groupnorm_apply(xl, st, gmm, btt, 32, True))
    wt = ops.conv_weight_taps((torch.randn(64, 64, 3, 3, device=dev) / 24).bfloat16())
    assert eq(NS.conv3x3(xl, wt, None, None), ops.conv3x3(xl, wt, None))
    assert eq(NS.pixel_shuffle2(xl), ops.pixel_shuffle2(xl))
    w16 = (torch.randn(16, 64, 3, 3, device=dev) / 24).bfloat16()
    assert eq(NS.conv3x3_to_nchw16(xl, w16, None), ops.conv3x3_to_nchw16(xl, w16, None))
    addt = torch.randn(M, D, device=dev).bfloat16()
    assert eq(NS.linear_bias_add(a, w, bias, addt), ops.gemm(a, w, epilogue=ops.EPI_BIAS_ADD, bias=bias, add2=addt))
    assert torch.allclose(NS.linear_bias_add(a, w, bias, addt).float(), NS.linear(a, w, bias, 1).float() + addt.float(), atol=0.05)
