"""GPU parity of every hand-written kernel, called through the C-ABI (ctypes), against a plain torch fp32
restatement of the same op on the same seeded inputs.  Tolerance: bf16 outputs -> rel-L2 <= 4e-3; fp32 outputs
(sampler update, GEMV, timestep embedding) -> 1e-5.  Shapes cover tile tails (M not a multiple of 128, N-tile 64/128/192),
text/image segment boundaries inside a tile, batch-row modulation, row remaps, K/V tails and nq != nkv."""
import pytest
import torch

from landiff_b200 import ops
from landiff_b200._C import (EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_POS, EPI_GATED_RESID, EPI_NONE, EPI_QKV, EPI_UNPATCHIFY)

pytestmark = pytest.mark.gpu
dev = "cuda"
BF16_TOL = 4e-3


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def report(name, got, ref, tol=BF16_TOL):
    assert torch.isfinite(got.float()).all(), f"{name}: non-finite output"
    r = rel(got, ref)
    assert r <= tol, f"{name}: rel-L2 {r:.3e} > {tol}"

def test_gemm_all_epilogues():
    torch.manual_seed(0)
    for (M, N, K) in [(128, 192, 64), (128, 192, 128), (300, 192, 128), (1000, 1920, 1920), (515, 256, 192), (130, 64, 1920),
                      (4096, 7680, 1920), (2000, 1920, 7680)]:
        a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        out = ops.gemm(a, w, epilogue=EPI_NONE)
        torch.cuda.synchronize()
        report(f"gemm NONE {M}x{N}x{K}", out, a.float() @ w.float().T)
    M, N, K = 2 * 443, 1920, 1920
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = (torch.randn(N, device=dev) * 0.1).bfloat16()
    ref = a.float() @ w.float().T + bias.float()
    report("gemm BIAS", ops.gemm(a, w, epilogue=EPI_BIAS, bias=bias), ref)
    report("gemm BIAS_GELU", ops.gemm(a, w, epilogue=EPI_BIAS_GELU, bias=bias),
           torch.nn.functional.gelu(ref, approximate="tanh"))
    # gated residual, 2 samples x 443 tokens, 226... use text_len 100
    B, R, TL = 2, 443, 100
    mod = torch.randn(B, 12, N, device=dev) * 0.5
    resid = torch.randn(M, N, device=dev).bfloat16()
    add2 = torch.randn(M, N, device=dev).bfloat16()
    gate_img, gate_txt = mod[:, 2], mod[:, 8]
    tok = torch.arange(M, device=dev) % R
    bidx = torch.arange(M, device=dev) // R
    gsel = torch.where((tok < TL)[:, None], gate_txt[bidx], gate_img[bidx])
    for use_add2 in (False, True):
        out = ops.gemm(a, w, epilogue=EPI_GATED_RESID, bias=bias, rows_per_batch=R, text_len=TL, resid=resid,
                       add2=add2 if use_add2 else None, gate_img=gate_img, gate_txt=gate_txt, mod_batch_stride=12 * N)
        r = resid.float() + gsel * ref + (add2.float() if use_add2 else 0)
        report(f"gemm GATED_RESID add2={use_add2}", out, r)
    # in-place variant (out aliases resid)
    res2 = resid.clone()
    ops.gemm(a, w, epilogue=EPI_GATED_RESID, bias=bias, rows_per_batch=R, text_len=TL, resid=res2, out=res2,
             gate_img=gate_img, gate_txt=gate_txt, mod_batch_stride=12 * N)
    report("gemm GATED_RESID in-place", res2, resid.float() + gsel * ref)
    # QKV
    H = 6
    N3 = 3 * H * 64
    wq = (torch.randn(N3, K, device=dev) * 0.05).bfloat16()
    bq = (torch.randn(N3, device=dev) * 0.1).bfloat16()
    lnp = [(1 + 0.1 * torch.randn(64, device=dev)).bfloat16(), (0.1 * torch.randn(64, device=dev)).bfloat16(),
           (1 + 0.1 * torch.randn(64, device=dev)).bfloat16(), (0.1 * torch.randn(64, device=dev)).bfloat16()]
    q = torch.zeros(B, H, R + 5, 64, device=dev, dtype=torch.bfloat16)
    k = torch.zeros_like(q)
    v = torch.zeros_like(q)
    ops.gemm(a, wq, epilogue=EPI_QKV, bias=bq, rows_per_batch=R, qkv=(q, k, v), qk_ln=lnp, ln_eps=1e-6, heads=H,
             qkv_row_offset=5)
    qkv_ref = (a.float() @ wq.float().T + bq.float()).bfloat16().float().view(B, R, 3, H, 64).permute(2, 0, 3, 1, 4)
    ln = torch.nn.functional.layer_norm
    report("gemm QKV q", q[:, :, 5:], ln(qkv_ref[0], (64,), lnp[0].float(), lnp[1].float(), 1e-6))
    report("gemm QKV k", k[:, :, 5:], ln(qkv_ref[1], (64,), lnp[2].float(), lnp[3].float(), 1e-6))
    report("gemm QKV v", v[:, :, 5:], qkv_ref[2])
    # BIAS_POS with row remap: M rows of image tokens written after TL text rows
    pos = torch.randn(TL + R, N, device=dev).bfloat16()
    hidden = torch.zeros(B, TL + R, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(a, w, epilogue=EPI_BIAS_POS, bias=bias, out=hidden.view(-1, N), rows_per_batch=R, out_rows_per_batch=TL + R,
             out_row_offset=TL, tok_offset=TL, text_len=TL, pos=pos)
    report("gemm BIAS_POS remap", hidden[:, TL:], ref.view(B, R, N) + pos[TL:].float())
    assert bool((hidden[:, :TL] == 0).all()), "BIAS_POS remap touched text rows"
    # UNPATCHIFY: T=2,Hp=3,Wp=5 -> 30 image tokens per sample
    T, Hp, Wp, Cc = 2, 3, 5, 16
    n_img = T * Hp * Wp
    a2 = (torch.randn(B * n_img, K, device=dev) * 0.5).bfloat16()
    w2 = (torch.randn(64, K, device=dev) * 0.05).bfloat16()
    b2 = (torch.randn(64, device=dev) * 0.1).bfloat16()
    out = torch.zeros(B, T, Cc, 2 * Hp, 2 * Wp, device=dev, dtype=torch.bfloat16)
    ops.gemm(a2, w2, epilogue=EPI_UNPATCHIFY, bias=b2, out=out, rows_per_batch=n_img, tok_offset=TL, text_len=TL,
             patch_grid=(T, Hp, Wp, Cc))
    y = (a2.float() @ w2.float().T + b2.float()).view(B, T, Hp, Wp, Cc, 2, 2)
    report("gemm UNPATCHIFY", out, y.permute(0, 1, 4, 2, 5, 3, 6).reshape(B, T, Cc, 2 * Hp, 2 * Wp))


# lse is log2(sum of the bf16-rounded P the PV product uses) + reference: accumulated by the tensor core from the same
# values as the numerator (fast path), so it carries their zero-mean rounding: up to ~1e-3 absolute on |lse| ~ 6 for short rows
LSE_TOL = 2e-4


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5])
def test_attention(variant):
    torch.manual_seed(1)
    for (B, H, nq, nkv) in [(1, 1, 128, 128), (1, 1, 256, 128), (1, 2, 300, 300), (2, 3, 886, 886), (1, 2, 500, 1000),
                            (1, 4, 4444, 17776), (1, 1, 200, 50), (1, 2, 256, 192), (1, 1, 130, 8888)]:
        q = torch.randn(B, H, nq, 64, device=dev).bfloat16()
        k = torch.randn(B, H, nkv, 64, device=dev).bfloat16()
        v = torch.randn(B, H, nkv, 64, device=dev).bfloat16()
        lse = torch.zeros(B * H, nq, device=dev)
        of = torch.zeros(B * H, nq, 64, device=dev)
        out = ops.attention(q, k, v, variant=variant, lse=lse, out_f32=of)
        torch.cuda.synchronize()
        ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float())
        report(f"attn v{variant} B{B} H{H} nq{nq} nkv{nkv}", out, ref.permute(0, 2, 1, 3).reshape(B, nq, H * 64))
        s = (q.float() @ k.float().transpose(-1, -2)) * 0.125
        lse_ref = torch.logsumexp(s, -1) * 1.4426950408889634
        report("   lse", lse, lse_ref.view(B * H, nq), LSE_TOL)
        report("   out_f32", of, ref.reshape(B * H, nq, 64))


@pytest.mark.parametrize("variant", [0, 1, 2, 4, 5])
def test_attention_overflow_fixup(variant):
    """One late key whose score is hundreds of log2-units above every row's first-block maximum: the fixed reference
    maximum of the fast path overflows (on MUFU columns: +inf; on polynomial columns: the clamped exponent field 255) and
    the CTA re-runs its rows through the exact path inside the same launch.  Rows with a negative projection on that key
    never overflow (mixed flagged / unflagged CTAs).  The key sits on column 2500 % 64 = 4 (pair 2: MUFU for KP = 4,
    polynomial for none); a second run moves it to a polynomial column."""
    torch.manual_seed(5)
    B, H, nq, nkv = 1, 2, 1000, 3000
    q = torch.randn(B, H, nq, 64, device=dev)
    k = torch.randn(B, H, nkv, 64, device=dev)
    v = torch.randn(B, H, nkv, 64, device=dev)
    q[:, 0, :, 0] = q[:, 0, :, 0].abs() + 3.0          # head 0: every row overflows
    q[:, 1, :300, 0] = -(q[:, 1, :300, 0].abs() + 3.0)  # head 1: first 300 rows never do, the rest do
    q[:, 1, 300:, 0] = q[:, 1, 300:, 0].abs() + 3.0
    q, v = q.bfloat16(), v.bfloat16()
    for key in (2500, 2496, 2561):     # columns 4 (MUFU pair for KP = 4), 0 (polynomial pair), 1 (second lane of a polynomial pair)
        kk = k.clone()
        kk[:, :, key, :] = 0
        kk[:, :, key, 0] = 400.0
        kk = kk.bfloat16()
        lse = torch.zeros(B * H, nq, device=dev)
        out = ops.attention(q, kk, v, variant=variant, lse=lse)
        torch.cuda.synchronize()
        ref = torch.nn.functional.scaled_dot_product_attention(q.float(), kk.float(), v.float())
        report(f"attn overflow v{variant} key {key}", out, ref.permute(0, 2, 1, 3).reshape(B, nq, H * 64))
        s = (q.float() @ kk.float().transpose(-1, -2)) * 0.125
        report("   lse", lse, (torch.logsumexp(s, -1) * 1.4426950408889634).view(B * H, nq), LSE_TOL)


@pytest.mark.parametrize("k0", [40.0, 90.0, 150.0])
def test_attention_late_dominant_key_without_overflow(k0):
    """A late key 25-110 log2-units above every row's first-block maximum: below the overflow limit, so the default
    kernel keeps its fixed reference maximum (P up to 2^110) and must still match — floating point is scale-invariant."""
    torch.manual_seed(6)
    B, H, nq, nkv = 1, 2, 520, 2000
    q = torch.randn(B, H, nq, 64, device=dev)
    k = torch.randn(B, H, nkv, 64, device=dev)
    v = torch.randn(B, H, nkv, 64, device=dev)
    q[..., 0] = q[..., 0].abs() + 3.0
    k[:, :, 1500, :] = 0
    k[:, :, 1500, 0] = k0
    k[:, :, 1700, :] = 0
    k[:, :, 1700, 0] = k0 - 1.0       # a second, slightly weaker key so the result is not a single-row copy
    q, k, v = q.bfloat16(), k.bfloat16(), v.bfloat16()
    lse = torch.zeros(B * H, nq, device=dev)
    out = ops.attention(q, k, v, lse=lse)
    torch.cuda.synchronize()
    ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float())
    report(f"attn dominant key {k0}", out, ref.permute(0, 2, 1, 3).reshape(B, nq, H * 64))
    s = (q.float() @ k.float().transpose(-1, -2)) * 0.125
    report("   lse", lse, (torch.logsumexp(s, -1) * 1.4426950408889634).view(B * H, nq), LSE_TOL)


@pytest.mark.parametrize("variant", [0, 1])
def test_attention_over_kv_shards_equals_monolithic(variant):
    """ONE launch over 2-4 K/V shards (ragged shard lengths: box tails of 1 and 2 sub-blocks, a shard shorter than one
    sub-block, the sequence-parallel shapes 4444 / 8888) == attention over the concatenated keys.  Shards live in
    separate buffers with more rows than used (kv_rows > nkv)."""
    torch.manual_seed(7)
    for (B, H, nq, lens) in [(1, 2, 300, (128, 128)), (1, 2, 200, (64, 200, 30)), (2, 3, 443, (443, 443)),
                             (1, 2, 384, (1, 129, 65, 191)), (1, 2, 500, (4444, 4444, 4444, 4444)), (1, 1, 260, (8888, 8888))]:
        nkv = sum(lens)
        q = torch.randn(B, H, nq, 64, device=dev).bfloat16()
        k = torch.randn(B, H, nkv, 64, device=dev).bfloat16()
        v = torch.randn(B, H, nkv, 64, device=dev).bfloat16()
        shards, o = [], 0
        for n in lens:
            kb = torch.full((B, H, n + 37, 64), float("nan"), device=dev, dtype=torch.bfloat16)   # rows beyond nkv: poison
            vb = torch.full((B, H, n + 37, 64), float("nan"), device=dev, dtype=torch.bfloat16)
            kb[:, :, :n] = k[:, :, o:o + n]
            vb[:, :, :n] = v[:, :, o:o + n]
            shards.append((kb, vb, n))
            o += n
        lse = torch.zeros(B * H, nq, device=dev)
        out = ops.attention_shards(q, shards, variant=variant, lse=lse)
        torch.cuda.synchronize()
        ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float())
        report(f"attn shards v{variant} {lens}", out, ref.permute(0, 2, 1, 3).reshape(B, nq, H * 64))
        s = (q.float() @ k.float().transpose(-1, -2)) * 0.125
        report("   lse", lse, (torch.logsumexp(s, -1) * 1.4426950408889634).view(B * H, nq), LSE_TOL)
        if variant == 0:   # bit-identical to the single-buffer launch when the shard boundaries fall on 128-key boxes
            if all(n % 128 == 0 for n in lens[:-1]):
                mono = ops.attention(q, k, v)
                assert torch.equal(mono, out), f"shards {lens}: differs from the monolithic launch"


def test_attention_tail_split():
    """Tail split (wave quantisation): without an lse request, the query blocks of the last, partly empty wave are served by
    several CTAs that each take a share of the key boxes and write fp32 partial results, merged by a second kernel.  Shapes:
    a grid with full waves + a tail (160 blocks), grids smaller than one wave (every block is split, up to 16 ways), key
    ranges that cross ragged shard boundaries, and an overflowing key inside ONE split (exact-path redo of a split CTA).
    Variant 6 is the same kernel with one CTA per query block."""
    torch.manual_seed(11)
    for (B, H, nq, lens) in [(1, 10, 4000, (4000,)), (1, 4, 4444, (17776,)), (1, 2, 500, (4444, 4444, 4444, 4444)),
                             (1, 3, 700, (1000, 129, 65, 2000)), (2, 2, 300, (1537,))]:
        nkv = sum(lens)
        q = torch.randn(B, H, nq, 64, device=dev).bfloat16()
        k = torch.randn(B, H, nkv, 64, device=dev).bfloat16()
        v = torch.randn(B, H, nkv, 64, device=dev).bfloat16()
        shards, o = [], 0
        for n in lens:
            shards.append((k[:, :, o:o + n].contiguous(), v[:, :, o:o + n].contiguous(), n))
            o += n
        out = ops.attention_shards(q, shards)                 # variant 0: tail split where the plan finds one
        plain = ops.attention_shards(q, shards, variant=6)
        torch.cuda.synchronize()
        ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float())
        ref = ref.permute(0, 2, 1, 3).reshape(B, nq, H * 64)
        report(f"attn tail split B{B} H{H} nq{nq} {lens}", out, ref)
        report("   no split", plain, ref)
        assert (out.float() - plain.float()).abs().max().item() < 2e-2
        assert torch.equal(out, ops.attention_shards(q, shards)), "tail split is not reproducible"
    # an overflowing key inside one split: that split's CTAs re-run through the exact path and still write partials
    B, H, nq, nkv = 1, 2, 1000, 3000
    q = torch.randn(B, H, nq, 64, device=dev)
    k = torch.randn(B, H, nkv, 64, device=dev)
    v = torch.randn(B, H, nkv, 64, device=dev).bfloat16()
    q[..., 0] = q[..., 0].abs() + 3.0
    k[:, :, 2500, :] = 0
    k[:, :, 2500, 0] = 400.0
    q, k = q.bfloat16(), k.bfloat16()
    out = ops.attention(q, k, v)
    torch.cuda.synchronize()
    ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float())
    report("attn tail split + overflow", out, ref.permute(0, 2, 1, 3).reshape(B, nq, H * 64))


def test_attention_shard_arrival_flags():
    """A shard guarded by an arrival flag is not read before the flag reaches the expected value: the flag is raised by
    a stream memory operation on a SECOND stream after a delay and a late fill of the buffer; without the in-kernel wait
    the kernel would read the poison the buffer holds until then.  A flag that never arrives times out after ~2 s,
    raises the status word and does not hang."""
    import ctypes as C

    from landiff_b200 import _C

    torch.manual_seed(8)
    B, H, nq, n0, n1 = 1, 2, 256, 192, 320
    q = torch.randn(B, H, nq, 64, device=dev).bfloat16()
    k = torch.randn(B, H, n0 + n1, 64, device=dev).bfloat16()
    v = torch.randn(B, H, n0 + n1, 64, device=dev).bfloat16()
    k0, v0 = k[:, :, :n0].contiguous(), v[:, :, :n0].contiguous()
    k1 = torch.full((B, H, n1, 64), float("nan"), device=dev, dtype=torch.bfloat16)
    v1 = torch.full((B, H, n1, 64), float("nan"), device=dev, dtype=torch.bfloat16)
    flag = torch.zeros(64, device=dev, dtype=torch.int32)
    side = torch.cuda.Stream()
    # first use of a kernel triggers a lazy module load, which cannot proceed while another kernel of this context is
    # spinning: run everything the side stream will launch once before (in production the spinning kernel waits for a
    # PEER's copy engine, which needs nothing from this context)
    with torch.cuda.stream(side):
        torch.cuda._sleep(1000)
        scratch = torch.empty_like(k1)
        scratch.copy_(k[:, :, n0:])
        _C.check(_C.load().ld_stream_write_u32(flag.data_ptr() + 128, 1, side.cuda_stream), "ld_stream_write_u32")
    torch.cuda.synchronize()
    out = ops.attention_shards(q, [(k0, v0, None), (k1, v1, None, flag.data_ptr(), 7)])   # starts, then spins on the flag
    with torch.cuda.stream(side):
        torch.cuda._sleep(int(2e7))                      # ~10 ms: the kernel is certainly waiting by now
        k1.copy_(k[:, :, n0:])
        v1.copy_(v[:, :, n0:])
        _C.check(_C.load().ld_stream_write_u32(flag.data_ptr(), 7, side.cuda_stream), "ld_stream_write_u32")
    torch.cuda.synchronize()
    ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float())
    report("attn shard behind an arrival flag", out, ref.permute(0, 2, 1, 3).reshape(B, nq, H * 64))
    assert ops.attention_status() == 0
    # transfer ids wrap: (int32)(flag - value) >= 0
    out = ops.attention_shards(q, [(k0, v0, None), (k1, v1, None, flag.data_ptr(), 5)])
    torch.cuda.synchronize()
    report("attn shard, flag already past the id", out, ref.permute(0, 2, 1, 3).reshape(B, nq, H * 64))
    # a flag that never arrives: bounded wait, status bit, no hang
    ops.attention_shards(q, [(k0, v0, None), (k1, v1, None, flag.data_ptr(), 9)])
    torch.cuda.synchronize()
    assert ops.attention_status() == 1 and ops.attention_status() == 0


def test_row_kernels():
    torch.manual_seed(2)
    B, R, TL, D = 2, 443, 100, 1920
    x = torch.randn(B * R, D, device=dev).bfloat16()
    w = (1 + 0.1 * torch.randn(D, device=dev)).bfloat16()
    b = (0.1 * torch.randn(D, device=dev)).bfloat16()
    mod = torch.randn(B, 12, D, device=dev) * 0.5
    out = ops.layernorm_modulate(x, w, b, 1e-5, mod[:, 0], mod[:, 1], mod[:, 6], mod[:, 7], 12 * D, B, R, 0, TL)
    ln = torch.nn.functional.layer_norm(x.float(), (D,), w.float(), b.float(), 1e-5).view(B, R, D)
    tok = torch.arange(R, device=dev)
    shift = torch.where((tok < TL)[None, :, None], mod[:, 6][:, None], mod[:, 0][:, None])
    scale = torch.where((tok < TL)[None, :, None], mod[:, 7][:, None], mod[:, 1][:, None])
    report("layernorm_modulate", out.view(B, R, D), ln * (1 + scale) + shift)
    w2 = (1 + 0.1 * torch.randn(D, device=dev)).bfloat16()
    b2 = (0.1 * torch.randn(D, device=dev)).bfloat16()
    fm = torch.randn(B, 2, D, device=dev) * 0.5
    out = ops.final_norm_modulate(x, w, b, 1e-5, w2, b2, 1e-6, fm[:, 0], fm[:, 1], 2 * D, B, R, 0, TL)
    l1 = torch.nn.functional.layer_norm(x.float(), (D,), w.float(), b.float(), 1e-5).bfloat16().float().view(B, R, D)[:, TL:]
    l2 = torch.nn.functional.layer_norm(l1, (D,), w2.float(), b2.float(), 1e-6)
    report("final_norm_modulate", out.view(B, R - TL, D), l2 * (1 + fm[:, 1][:, None]) + fm[:, 0][:, None])
    T, Cc, Hp, Wp = 2, 16, 15, 22
    xin = torch.randn(B, T, Cc, 2 * Hp, 2 * Wp, device=dev)
    sem = (torch.randn(1, T, Cc, 2 * Hp, 2 * Wp, device=dev) * 0.1).bfloat16()
    cols = ops.patchify(xin, sem)
    xs = (xin.bfloat16().float() + sem.float())
    ref = xs.view(B, T, Cc, Hp, 2, Wp, 2).permute(0, 1, 3, 5, 2, 4, 6).reshape(B * T * Hp * Wp, Cc * 4)
    report("patchify (+sem)", cols, ref)
    cols = ops.patchify(xin, None, g0=100, n=300)
    ref = xin.view(B, T, Cc, Hp, 2, Wp, 2).permute(0, 1, 3, 5, 2, 4, 6).reshape(B, T * Hp * Wp, Cc * 4)[:, 100:400]
    report("patchify shard", cols.view(B, 300, 64), ref)
    xe = torch.randn(B, 512, device=dev)
    wl = (torch.randn(23040, 512, device=dev) * 0.05).bfloat16()
    bl = (torch.randn(23040, device=dev) * 0.1).bfloat16()
    y = ops.small_linear(xe, wl, bl, act_in=1, round_bf16=False)
    report("small_linear silu-in", y, torch.nn.functional.silu(xe) @ wl.float().T + bl.float(), 1e-5)
    t = torch.tensor([999.0, 19.0], device=dev)
    te = ops.timestep_embedding(t, 1920, round_bf16=False)
    half = 960
    freqs = torch.exp(-torch.log(torch.tensor(10000.0)) * torch.arange(half, dtype=torch.float32) / half).to(dev)
    args = t[:, None] * freqs[None]
    report("timestep_embedding", te, torch.cat([torch.cos(args), torch.sin(args)], -1), 1e-4)
    n = 13 * 16 * 60 * 90
    xx, old, eps = torch.randn(n, device=dev), torch.randn(n, device=dev), torch.randn(n, device=dev)
    nu, nc = torch.randn(n, device=dev).bfloat16(), torch.randn(n, device=dev).bfloat16()
    kw = dict(c_skip=0.3, c_out=-0.95, cfg=4.5, m1=0.9, m2=-0.2, m3=1.7, m4=0.7, mn=0.1)
    du = nu.float() * kw["c_out"] + xx * kw["c_skip"]
    dc = nc.float() * kw["c_out"] + xx * kw["c_skip"]
    den = du + kw["cfg"] * (dc - du)
    xo, do = ops.sampler_update(xx, nu, nc, old, eps, mode=1, **kw)
    report("sampler_update mode1 x", xo, kw["m1"] * xx - kw["m2"] * (kw["m3"] * den - kw["m4"] * old) + kw["mn"] * eps, 1e-5)
    report("sampler_update den", do, den, 1e-5)
    xo, _ = ops.sampler_update(xx, nu, nc, None, eps, mode=0, **kw)
    report("sampler_update mode0 x", xo, kw["m1"] * xx - kw["m2"] * den + kw["mn"] * eps, 1e-5)
    xo, _ = ops.sampler_update(xx, nu, nc, None, None, mode=2, **kw)
    report("sampler_update mode2 x", xo, den, 1e-5)


def test_unpatchify_blocks_equals_unpatchify_epilogue():
    """Token-major blocks scattered by `ld_unpatchify_blocks` (the copy-engine output exchange of the parallel layouts) ==
    the UNPATCHIFY GEMM epilogue writing the latent layout directly, bit for bit; blocks in arbitrary order, two rows."""
    torch.manual_seed(3)
    T, Hp, Wp, D = 3, 5, 7, 128
    n_img = T * Hp * Wp
    a = (torch.randn(2 * n_img, D, device=dev) * 0.5).bfloat16()
    w = (torch.randn(64, D, device=dev) * 0.05).bfloat16()
    b = (torch.randn(64, device=dev) * 0.1).bfloat16()
    want = torch.zeros(2, T, 16, 2 * Hp, 2 * Wp, device=dev, dtype=torch.bfloat16)
    ops.gemm(a, w, epilogue=EPI_UNPATCHIFY, bias=b, out=want, rows_per_batch=n_img, tok_offset=0, text_len=0,
             patch_grid=(T, Hp, Wp, 16))
    tok = ops.gemm(a, w, epilogue=EPI_BIAS, bias=b).view(2, n_img, 64)
    cuts = [0, 17, 60, n_img]
    blocks = []
    for r in (1, 0):
        for i in (2, 0, 1):
            g0, g1 = cuts[i], cuts[i + 1]
            blocks.append((tok[r, g0:g1].contiguous(), r, g0, g1 - g0))
    got = torch.full_like(want, float("nan"))
    ops.unpatchify_blocks(blocks, got)
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    with pytest.raises(ValueError):
        ops.unpatchify_blocks([(tok[0], 0, n_img - 3, 10)], got)


def test_attention_merge_equals_monolithic():
    """ring hop merge: attention over two K/V halves merged by (O, LSE) == attention over the whole K/V."""
    torch.manual_seed(3)
    B, H, nq, nkv = 1, 3, 300, 700
    q = torch.randn(B, H, nq, 64, device=dev).bfloat16()
    k = torch.randn(B, H, nkv, 64, device=dev).bfloat16()
    v = torch.randn(B, H, nkv, 64, device=dev).bfloat16()
    ref = ops.attention(q, k, v).clone()
    halves = []
    for lo, hi in ((0, 300), (300, 700)):
        lse = torch.zeros(B * H, nq, device=dev)
        of = torch.zeros(B * H, nq, 64, device=dev)
        ops.attention(q, k[:, :, lo:hi].contiguous(), v[:, :, lo:hi].contiguous(), lse=lse, out_f32=of)
        halves.append((of, lse))
    out = torch.zeros(B, nq, H * 64, device=dev, dtype=torch.bfloat16)
    ops.attention_merge(halves[0][0], halves[0][1], halves[1][0], halves[1][1], out, B, H, nq)
    torch.cuda.synchronize()
    report("merge vs monolithic", out, ref, 5e-3)


def test_argument_errors_are_reported():
    from landiff_b200._C import LanDiffB200Error

    a = torch.zeros(128, 100, device=dev, dtype=torch.bfloat16)  # K not a multiple of 64
    w = torch.zeros(64, 100, device=dev, dtype=torch.bfloat16)
    with pytest.raises(LanDiffB200Error, match="multiple of 64"):
        ops.gemm(a, w, epilogue=EPI_NONE)
    with pytest.raises(TypeError):
        ops.gemm(a.float(), w, epilogue=EPI_NONE)
