"""Tensor-level wrappers over the C-ABI (shape / dtype / contiguity / device checks live here, the C side only
sees pointers and sizes).  All kernels are enqueued on torch's current CUDA stream.

The same functions are also registered as `torch.ops.landiff_b200.*` custom ops (see `register_torch_ops`).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _C
from ._C import (EPI_BIAS, EPI_BIAS_ADD, EPI_BIAS_GELU, EPI_BIAS_POS, EPI_GATED_RESID, EPI_NONE, EPI_QKV, EPI_UNPATCHIFY,
                 GemmArgs, check)

BF16 = torch.bfloat16
F32 = torch.float32


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, dtype, name: str, contiguous: bool = True) -> None:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (landiff_b200 has no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if contiguous and not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


def device_check() -> int:
    sms = C.c_int(0)
    check(_C.load().ld_device_check(C.byref(sms)), "ld_device_check")
    return sms.value


def gemm(a: torch.Tensor, w: torch.Tensor, *, epilogue: int, bias: Optional[torch.Tensor] = None,
         out: Optional[torch.Tensor] = None, rows_per_batch: Optional[int] = None, out_rows_per_batch: Optional[int] = None,
         out_row_offset: int = 0, tok_offset: int = 0, text_len: int = 0, resid: Optional[torch.Tensor] = None,
         add2: Optional[torch.Tensor] = None, gate_img: Optional[torch.Tensor] = None,
         gate_txt: Optional[torch.Tensor] = None, mod_batch_stride: int = 0, qkv=None, qk_ln=None, ln_eps: float = 1e-6,
         q_scale: float = 1.0, heads: int = 0, qkv_row_offset: int = 0, pos: Optional[torch.Tensor] = None,
         patch_grid=None, conv: bool = False) -> Optional[torch.Tensor]:
    """out = epilogue(a @ w.T).  a: [M,K] bf16, w: [N,K] bf16.  See include/landiff_b200.h for the epilogues.
    conv=True: implicit 3x3 / stride 1 / padding 1 convolution — a is a channels-last activation tensor [F, H, W, C] and
    w is [N, 9*C] in (ky, kx, cin) order; the A tiles are gathered by TMA in im2col mode, M = F*H*W."""
    _chk(a, BF16, "a")
    _chk(w, BF16, "w")
    g = GemmArgs()
    if conv:
        if a.dim() != 4 or a.shape[3] % 64:
            raise ValueError("gemm(conv=True): a must be channels-last [F, H, W, C] with C a multiple of 64")
        F_, H_, W_, C_ = a.shape
        M, K = F_ * H_ * W_, 9 * C_
        g.conv_F, g.conv_H, g.conv_W, g.conv_C = F_, H_, W_, C_
    else:
        M, K = a.shape
    N, K2 = w.shape
    if K != K2:
        raise ValueError(f"gemm: inner dims differ ({K} vs {K2})")
    g.M, g.N, g.K, g.epilogue = M, N, K, epilogue
    g.A, g.W = a.data_ptr(), w.data_ptr()
    if bias is not None:
        _chk(bias, BF16, "bias")
        if bias.numel() != N:
            raise ValueError("gemm: bias size")
        g.bias = bias.data_ptr()
    g.rows_per_batch = rows_per_batch or M
    g.out_rows_per_batch = out_rows_per_batch if out_rows_per_batch is not None else g.rows_per_batch
    g.out_row_offset, g.tok_offset, g.text_len = out_row_offset, tok_offset, text_len
    ret = None
    if epilogue == EPI_QKV:
        q, k, v = qkv
        for n_, t_ in (("q", q), ("k", k), ("v", v)):
            _chk(t_, BF16, n_)
            if t_.dim() != 4 or t_.shape[-1] != 64 or t_.shape[1] != heads:
                raise ValueError(f"gemm: {n_} must be [B, heads, rows, 64]")
        if not (q.shape[2] == k.shape[2] == v.shape[2]):
            raise ValueError("gemm: q/k/v row counts differ")
        qw, qb, kw, kb = qk_ln
        for t_ in (qw, qb, kw, kb):
            _chk(t_, BF16, "qk_ln")
        g.q, g.k, g.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
        g.q_ln_w, g.q_ln_b, g.k_ln_w, g.k_ln_b = qw.data_ptr(), qb.data_ptr(), kw.data_ptr(), kb.data_ptr()
        g.ln_eps, g.q_scale = ln_eps, q_scale
        g.heads, g.qkv_rows, g.qkv_row_offset = heads, q.shape[2], qkv_row_offset
    else:
        if epilogue == EPI_UNPATCHIFY:
            if out is None:
                raise ValueError("gemm: UNPATCHIFY needs out")
            T, Hp, Wp, Cc = patch_grid
            g.T, g.Hp, g.Wp, g.C = T, Hp, Wp, Cc
            _chk(out, BF16, "out")
            g.ld_out = 0
        else:
            if out is None:
                out = torch.empty((M, N), dtype=BF16, device=a.device)
            if out.dtype == F32 and epilogue in (EPI_GATED_RESID, EPI_BIAS_POS):
                g.out_f32 = 1  # fp32 residual stream
            else:
                _chk(out, BF16, "out", contiguous=False)
            if not out.is_cuda:
                raise ValueError("out must be a CUDA tensor")
            if out.stride(-1) != 1:
                raise ValueError("gemm: out must have unit inner stride")
            g.ld_out = out.stride(0) if out.dim() == 2 else N
        g.out = out.data_ptr()
        ret = out
        if epilogue == EPI_GATED_RESID:
            if resid.dtype == F32:
                g.resid_f32 = 1
            _chk(resid, resid.dtype if resid.dtype in (BF16, F32) else BF16, "resid", contiguous=False)
            _chk(gate_img, F32, "gate_img", contiguous=False)
            _chk(gate_txt, F32, "gate_txt", contiguous=False)
            g.resid, g.gate_img, g.gate_txt = resid.data_ptr(), gate_img.data_ptr(), gate_txt.data_ptr()
            g.mod_batch_stride = mod_batch_stride
            if add2 is not None:
                _chk(add2, BF16, "add2", contiguous=False)
                g.add2 = add2.data_ptr()
        if epilogue == EPI_BIAS_POS:
            _chk(pos, BF16, "pos")
            g.pos = pos.data_ptr()
        if epilogue == EPI_BIAS_ADD:
            _chk(add2, BF16, "add2")
            if tuple(add2.shape) != tuple(out.shape) or not out.is_contiguous():
                raise ValueError("gemm: BIAS_ADD needs a contiguous out and an add2 of the same shape")
            g.add2 = add2.data_ptr()
    check(_C.load().ld_gemm_bf16(C.byref(g), _stream()), "ld_gemm_bf16")
    return ret


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, nq: Optional[int] = None, nkv: Optional[int] = None,
              out: Optional[torch.Tensor] = None, lse: Optional[torch.Tensor] = None,
              out_f32: Optional[torch.Tensor] = None, variant: int = 0) -> torch.Tensor:
    """softmax(q k^T / 8) v for q,k,v [B, H, rows, 64] bf16 -> [B, nq, H*64] bf16 (full attention)."""
    for n_, t_ in (("q", q), ("k", k), ("v", v)):
        _chk(t_, BF16, n_)
        if t_.dim() != 4 or t_.shape[-1] != 64:
            raise ValueError(f"attention: {n_} must be [B, H, rows, 64]")
    B, H, q_rows, _ = q.shape
    kv_rows = k.shape[2]
    nq = q_rows if nq is None else nq
    nkv = kv_rows if nkv is None else nkv
    if out is None:
        out = torch.empty((B, nq, H * 64), dtype=BF16, device=q.device)
    _chk(out, BF16, "out")
    if k.shape != v.shape or k.shape[:2] != q.shape[:2]:
        raise ValueError(f"attention: q {tuple(q.shape)} / k {tuple(k.shape)} / v {tuple(v.shape)} do not agree")
    if not (0 < nq <= q_rows and 0 < nkv <= kv_rows):
        raise ValueError(f"attention: nq={nq} / nkv={nkv} outside the buffers ({q_rows} / {kv_rows} rows)")
    if out.numel() != B * nq * H * 64:
        raise ValueError(f"attention: out must hold [B, nq, H*64] = {B * nq * H * 64} elements, got {out.numel()}")
    if lse is not None:
        _chk(lse, F32, "lse")
        if lse.numel() != B * H * nq:
            raise ValueError("attention: lse must hold [B*H, nq] elements")
    if out_f32 is not None:
        _chk(out_f32, F32, "out_f32")
        if lse is None or out_f32.numel() != B * H * nq * 64:
            raise ValueError("attention: out_f32 must hold [B*H, nq, 64] elements and needs lse")
    one = (_C.KvShard * 1)()
    one[0].k, one[0].v, one[0].nkv, one[0].kv_rows = k.data_ptr(), v.data_ptr(), nkv, kv_rows
    _attention_launch(q, one, 1, out, lse, out_f32, B, H, nq, q_rows, variant)
    return out


def _attention_launch(q, arr, n, out, lse, out_f32, B, H, nq, q_rows, variant) -> None:
    """One launch through the workspace entry point: the tail split's partial results live in a torch allocation (stream-
    ordered caching allocator: safe across streams and under CUDA-graph capture)."""
    lib = _C.load()
    ws, ws_bytes = None, 0
    if lse is None and out_f32 is None and variant not in (1, 6):
        ws_bytes = int(lib.ld_attention_workspace_bytes(arr, n, B, H, nq))
        if ws_bytes:
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=q.device)
    check(lib.ld_attention_shards_ws_bf16(q.data_ptr(), arr, n, out.data_ptr(), _ptr(lse), _ptr(out_f32), B, H, nq, q_rows,
                                          variant, _ptr(ws), ws_bytes, _stream()), "ld_attention_shards_ws_bf16")


def attention_shards(q: torch.Tensor, shards, *, nq: Optional[int] = None, out: Optional[torch.Tensor] = None,
                     lse: Optional[torch.Tensor] = None, out_f32: Optional[torch.Tensor] = None, variant: int = 0) -> torch.Tensor:
    """softmax(q [k_0|k_1|...]^T / 8) [v_0|v_1|...] with K/V given as 1..4 shards of the key sequence, ONE launch.
    `shards`: sequence of (k, v) or (k, v, nkv) or (k, v, nkv, ready_flag_ptr, ready_value) with k, v [B, H, rows, 64]
    bf16; a shard with a ready flag (device address of a u32) may still be in flight: the kernel waits for
    (int32)(*flag - value) >= 0 before reading it (ring sequence parallelism, landiff_b200/parallel.py)."""
    _chk(q, BF16, "q")
    if q.dim() != 4 or q.shape[-1] != 64:
        raise ValueError("attention: q must be [B, H, rows, 64]")
    B, H, q_rows, _ = q.shape
    nq = q_rows if nq is None else nq
    if not 0 < nq <= q_rows:
        raise ValueError(f"attention: nq={nq} outside the q buffer ({q_rows} rows)")
    if not 1 <= len(shards) <= 4:
        raise ValueError(f"attention: 1..4 K/V shards, got {len(shards)}")
    arr = (_C.KvShard * len(shards))()
    for i, sh in enumerate(shards):
        k, v = sh[0], sh[1]
        for n_, t_ in (("k", k), ("v", v)):
            _chk(t_, BF16, n_)
            if t_.dim() != 4 or t_.shape[-1] != 64 or t_.shape[:2] != q.shape[:2]:
                raise ValueError(f"attention: shard {i} {n_} {tuple(t_.shape)} does not match q {tuple(q.shape)}")
        if k.shape != v.shape:
            raise ValueError(f"attention: shard {i} k / v shapes differ")
        nkv = k.shape[2] if len(sh) < 3 or sh[2] is None else int(sh[2])
        if not 0 < nkv <= k.shape[2]:
            raise ValueError(f"attention: shard {i} nkv={nkv} outside its buffer ({k.shape[2]} rows)")
        arr[i].k, arr[i].v, arr[i].nkv, arr[i].kv_rows = k.data_ptr(), v.data_ptr(), nkv, k.shape[2]
        if len(sh) >= 5 and sh[3]:
            arr[i].ready_flag, arr[i].ready_value = int(sh[3]), int(sh[4]) & 0xFFFFFFFF
    if out is None:
        out = torch.empty((B, nq, H * 64), dtype=BF16, device=q.device)
    _chk(out, BF16, "out")
    if out.numel() != B * nq * H * 64:
        raise ValueError(f"attention: out must hold [B, nq, H*64] = {B * nq * H * 64} elements, got {out.numel()}")
    if lse is not None:
        _chk(lse, F32, "lse")
        if lse.numel() != B * H * nq:
            raise ValueError("attention: lse must hold [B*H, nq] elements")
    if out_f32 is not None:
        _chk(out_f32, F32, "out_f32")
        if lse is None or out_f32.numel() != B * H * nq * 64:
            raise ValueError("attention: out_f32 must hold [B*H, nq, 64] elements and needs lse")
    _attention_launch(q, arr, len(shards), out, lse, out_f32, B, H, nq, q_rows, variant)
    return out


def attention_status(reset: bool = True) -> int:
    """Status word of the in-kernel shard waits on the current device (bit 0: a wait timed out).  Synchronises."""
    w = C.c_uint(0)
    rc = _C.load().ld_attention_status(C.byref(w), int(reset))
    if rc != 0:
        check(rc, "ld_attention_status")
    return w.value


def attention_merge(o_acc, lse_acc, o_new, lse_new, out_bf16, batch, heads, nq) -> None:
    for t_ in (o_acc, lse_acc, o_new, lse_new):
        _chk(t_, F32, "merge operand")
    check(_C.load().ld_attention_merge(o_acc.data_ptr(), lse_acc.data_ptr(), o_new.data_ptr(), lse_new.data_ptr(),
                                       _ptr(out_bf16), batch, heads, nq, _stream()), "ld_attention_merge")


def layernorm_modulate(x, w, b, eps, shift_img, scale_img, shift_txt, scale_txt, mod_batch_stride, batch,
                       rows_per_batch, tok_offset, text_len, out=None):
    _chk(x, x.dtype if x.dtype in (BF16, F32) else BF16, "x")
    D = x.shape[-1]
    if out is None:
        out = torch.empty(x.shape, dtype=BF16, device=x.device)
    _chk(out, BF16, "out")
    for t_ in (shift_img, scale_img, shift_txt, scale_txt):
        _chk(t_, F32, "modulation", contiguous=False)
    check(_C.load().ld_layernorm_modulate(x.data_ptr(), int(x.dtype == F32), out.data_ptr(), w.data_ptr(), b.data_ptr(), eps,
                                          shift_img.data_ptr(), scale_img.data_ptr(), shift_txt.data_ptr(),
                                          scale_txt.data_ptr(), mod_batch_stride, batch, rows_per_batch, tok_offset,
                                          text_len, D, _stream()), "ld_layernorm_modulate")
    return out


def final_norm_modulate(x, w1, b1, eps1, w2, b2, eps2, shift, scale, mod_batch_stride, batch, rows_per_batch,
                        tok_offset, text_len, out=None):
    _chk(x, x.dtype if x.dtype in (BF16, F32) else BF16, "x")
    D = x.shape[-1]
    first_img = max(text_len - tok_offset, 0)
    n_img = rows_per_batch - first_img
    if out is None:
        out = torch.empty((batch * n_img, D), dtype=BF16, device=x.device)
    check(_C.load().ld_final_norm_modulate(x.data_ptr(), int(x.dtype == F32), out.data_ptr(), w1.data_ptr(), b1.data_ptr(), eps1,
                                           w2.data_ptr(), b2.data_ptr(), eps2, shift.data_ptr(), scale.data_ptr(),
                                           mod_batch_stride, batch, rows_per_batch, tok_offset, text_len, D, _stream()),
          "ld_final_norm_modulate")
    return out


def patchify(x: torch.Tensor, sem: Optional[torch.Tensor], g0: int = 0, n: Optional[int] = None, out=None):
    """x: [B,T,C,H,W] fp32/bf16; sem: [1,T,C,H,W] or None -> cols [B*n, C*4] bf16 for image tokens [g0, g0+n)."""
    if x.dtype not in (F32, BF16):
        raise TypeError("patchify: x must be fp32 or bf16")
    _chk(x, x.dtype, "x")
    B, T, Cc, H, W = x.shape
    Hp, Wp = H // 2, W // 2
    n = T * Hp * Wp - g0 if n is None else n
    if sem is not None:
        if sem.dtype not in (F32, BF16):
            raise TypeError("patchify: sem must be fp32 or bf16")
        _chk(sem, sem.dtype, "sem")
        if sem.numel() != T * Cc * H * W:
            raise ValueError("patchify: sem must be [1,T,C,H,W]")
    if out is None:
        out = torch.empty((B * n, Cc * 4), dtype=BF16, device=x.device)
    check(_C.load().ld_patchify(x.data_ptr(), int(x.dtype == F32), _ptr(sem), int(sem is not None and sem.dtype == F32),
                                out.data_ptr(), B, T, Cc, Hp, Wp, g0, n, _stream()), "ld_patchify")
    return out


def unpatchify_blocks(blocks, out: torch.Tensor) -> torch.Tensor:
    """Scatter token-major bf16 [count, 64] blocks into the latent layout out [rows, T, 16, H, W].  `blocks`: sequence of
    (tensor or device address, row, g0, count) — block covers image tokens [g0, g0 + count) of output row `row`."""
    _chk(out, BF16, "out")
    if out.dim() != 5 or out.shape[2] != 16 or out.shape[3] % 2 or out.shape[4] % 2:
        raise ValueError("unpatchify_blocks: out must be [rows, T, 16, H, W] with even H, W")
    if not 1 <= len(blocks) <= 16:
        raise ValueError(f"unpatchify_blocks: 1..16 blocks, got {len(blocks)}")
    rows, T, Cc, H, W = out.shape
    tb = _C.TokenBlocks()
    for i, (src, row, g0, count) in enumerate(blocks):
        if isinstance(src, torch.Tensor):
            _chk(src, BF16, "block")
            if src.numel() < count * 64:
                raise ValueError("unpatchify_blocks: block tensor smaller than count x 64")
            src = src.data_ptr()
        if not (0 <= row < rows and g0 >= 0 and count >= 0 and g0 + count <= T * (H // 2) * (W // 2)):
            raise ValueError(f"unpatchify_blocks: block {i} (row {row}, tokens [{g0}, {g0 + count})) outside the output")
        tb.ptr[i], tb.row[i], tb.g0[i], tb.count[i] = int(src), int(row), int(g0), int(count)
    tb.n = len(blocks)
    check(_C.load().ld_unpatchify_blocks(C.byref(tb), out.data_ptr(), T, H // 2, W // 2, Cc, _stream()), "ld_unpatchify_blocks")
    return out


def small_linear(x, w, bias, act_in=0, act_out=0, round_bf16=True, out=None):
    _chk(x, F32, "x")
    _chk(w, BF16, "w")
    Bn, K = x.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((Bn, N), dtype=F32, device=x.device)
    _chk(out, F32, "out")
    check(_C.load().ld_small_linear(x.data_ptr(), w.data_ptr(), _ptr(bias), out.data_ptr(), Bn, N, K, act_in, act_out,
                                    int(round_bf16), _stream()), "ld_small_linear")
    return out


def small_linear_batched(x, w_ptrs, b_ptrs, layers, N, act_in=0, round_bf16=True, out=None):
    """y[l] = act_in(x) @ W_l.T + b_l for `layers` same-shaped bf16 matrices in one launch.  w_ptrs / b_ptrs: int64 CUDA
    tensors holding the device addresses of the matrices / bias vectors (see `pointer_table`)."""
    _chk(x, F32, "x")
    Bn, K = x.shape
    for nm, t_ in (("w_ptrs", w_ptrs), ("b_ptrs", b_ptrs)):
        _chk(t_, torch.int64, nm)
        if t_.numel() != layers:
            raise ValueError(f"small_linear_batched: {nm} must hold {layers} pointers")
    if out is None:
        out = torch.empty((layers, Bn, N), dtype=F32, device=x.device)
    _chk(out, F32, "out")
    if out.numel() != layers * Bn * N:
        raise ValueError("small_linear_batched: out must hold [layers, B, N]")
    check(_C.load().ld_small_linear_batched(x.data_ptr(), w_ptrs.data_ptr(), b_ptrs.data_ptr(), out.data_ptr(), layers, Bn, N, K,
                                            act_in, int(round_bf16), _stream()), "ld_small_linear_batched")
    return out


def pointer_table(tensors, device) -> torch.Tensor:
    """int64 CUDA tensor of the device addresses of `tensors` (None -> 0)."""
    return torch.tensor([0 if t is None else t.data_ptr() for t in tensors], dtype=torch.int64, device=device)


def timestep_embedding(t, dim, max_period=10000.0, round_bf16=True, out=None):
    _chk(t, F32, "t")
    Bn = t.numel()
    if out is None:
        out = torch.empty((Bn, dim), dtype=F32, device=t.device)
    check(_C.load().ld_timestep_embedding(t.data_ptr(), out.data_ptr(), Bn, dim, max_period, int(round_bf16), _stream()),
          "ld_timestep_embedding")
    return out


def sampler_update(x, net_u, net_c, old_den, eps, *, c_skip, c_out, cfg, m1=0.0, m2=0.0, m3=0.0, m4=0.0, mn=0.0, mode=0,
                   x_out=None, den_out=None, net_dtype=BF16):
    _chk(x, F32, "x")
    _chk(net_u, net_dtype, "net_u")
    _chk(net_c, net_dtype, "net_c")
    if net_u.numel() != x.numel() or net_c.numel() != x.numel():
        raise ValueError("sampler_update: net_u/net_c must have x's element count")
    for nm, t_ in (("old_den", old_den), ("eps", eps)):
        if t_ is not None:
            _chk(t_, F32, nm)
            if t_.numel() != x.numel():
                raise ValueError(f"sampler_update: {nm} size")
    n = x.numel()
    if x_out is None:
        x_out = torch.empty_like(x)
    if den_out is None:
        den_out = torch.empty_like(x)

    def fin(v):  # the reference's scalars may be +-inf-derived limits; pass them through as floats
        return float(v)

    check(_C.load().ld_sampler_update(x.data_ptr(), net_u.data_ptr(), net_c.data_ptr(), _ptr(old_den), _ptr(eps),
                                      x_out.data_ptr(), den_out.data_ptr(), n, fin(c_skip), fin(c_out), fin(cfg), fin(m1),
                                      fin(m2), fin(m3), fin(m4), fin(mn), mode, int(net_dtype == F32), _stream()),
          "ld_sampler_update")
    return x_out, den_out


def sampler_update_f32(x, den_u, den_c, old_den, eps, *, cfg, m1=0.0, m2=0.0, m3=0.0, m4=0.0, mn=0.0, mode=0):
    """CFG combine + DPM++ update on already-denoised fp32 rows (reference DiscreteDenoiser kept in the loop)."""
    return sampler_update(x, den_u, den_c, old_den, eps, c_skip=0.0, c_out=1.0, cfg=cfg, m1=m1, m2=m2, m3=m3, m4=m4, mn=mn,
                          mode=mode, net_dtype=F32)


# ---- semantic conditioner, upsample path: channels-last conv stack ---------------------------
def nchw_to_nhwc(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[F, C, H, W] bf16 / fp32 -> channels-last bf16 [F, H, W, C]."""
    if x.dtype not in (BF16, F32):
        raise TypeError(f"nchw_to_nhwc: x must be bf16 or fp32, got {x.dtype}")
    _chk(x, x.dtype, "x")
    if x.dim() != 4:
        raise ValueError("nchw_to_nhwc: x must be [F, C, H, W]")
    F_, C_, H, W = x.shape
    if out is None:
        out = torch.empty((F_, H, W, C_), dtype=BF16, device=x.device)
    _chk(out, BF16, "out")
    check(_C.load().ld_nchw_to_nhwc(x.data_ptr(), 1 if x.dtype == F32 else 0, out.data_ptr(), F_, C_, H * W, _stream()),
          "ld_nchw_to_nhwc")
    return out


def groupnorm_stats(x: torch.Tensor, groups: int = 32, eps: float = 1e-6) -> torch.Tensor:
    """x channels-last [F, H, W, C] bf16 -> fp32 [F, groups, 2] = (mean, rstd); two passes like torch, fixed reduction order."""
    _chk(x, BF16, "x")
    F_, H, W, C_ = x.shape
    stats = torch.empty((F_, groups, 2), dtype=F32, device=x.device)
    scratch = torch.empty((F_ * ((H * W + 63) // 64) * groups,), dtype=F32, device=x.device)
    check(_C.load().ld_groupnorm_stats(x.data_ptr(), stats.data_ptr(), scratch.data_ptr(), F_, H * W, C_, groups, float(eps),
                                       _stream()), "ld_groupnorm_stats")
    return stats


def groupnorm_apply(x: torch.Tensor, stats: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int = 32,
                    swish: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """swish(GroupNorm(x)) on channels-last frames [F, H, W, C] bf16 with `groupnorm_stats`' (mean, rstd)."""
    _chk(x, BF16, "x"); _chk(stats, F32, "stats"); _chk(gamma, BF16, "gamma"); _chk(beta, BF16, "beta")
    F_, H, W, C_ = x.shape
    if stats.numel() != F_ * groups * 2 or gamma.numel() != C_ or beta.numel() != C_:
        raise ValueError("groupnorm_apply: operand sizes")
    if out is None:
        out = torch.empty_like(x)
    _chk(out, BF16, "out")
    check(_C.load().ld_groupnorm_apply(x.data_ptr(), out.data_ptr(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), F_,
                                       H * W, C_, int(groups), 1 if swish else 0, _stream()), "ld_groupnorm_apply")
    return out


def im2col3x3(x: torch.Tensor, out: Optional[torch.Tensor] = None, *, gn=None, swish: bool = True) -> torch.Tensor:
    """x channels-last [F, H, W, C] bf16 -> [F*H*W, 9*C] bf16 (taps-major).  gn = (stats, gamma, beta, groups) applies
    GroupNorm (+ swish) to the input on the fly."""
    _chk(x, BF16, "x")
    F_, H, W, C_ = x.shape
    if out is None:
        out = torch.empty((F_ * H * W, 9 * C_), dtype=BF16, device=x.device)
    _chk(out, BF16, "out")
    if out.numel() != F_ * H * W * 9 * C_:
        raise ValueError("im2col3x3: out has the wrong size")
    if gn is None:
        st = gm = gb = None
        groups = 0
    else:
        st, gm, gb, groups = gn
        _chk(st, F32, "gn stats"); _chk(gm, BF16, "gamma"); _chk(gb, BF16, "beta")
        if st.numel() != F_ * groups * 2 or gm.numel() != C_ or gb.numel() != C_:
            raise ValueError("im2col3x3: GroupNorm operand sizes")
    check(_C.load().ld_im2col3x3(x.data_ptr(), out.data_ptr(), F_, H, W, C_, _ptr(st), _ptr(gm), _ptr(gb), int(groups),
                                 1 if swish else 0, _stream()), "ld_im2col3x3")
    return out


def pixel_shuffle2(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """torch.nn.PixelShuffle(2) on channels-last frames: [F, H, W, 4*Co] -> [F, 2H, 2W, Co]."""
    _chk(x, BF16, "x")
    F_, H, W, C4 = x.shape
    if C4 % 4:
        raise ValueError("pixel_shuffle2: channels must be a multiple of 4")
    if out is None:
        out = torch.empty((F_, 2 * H, 2 * W, C4 // 4), dtype=BF16, device=x.device)
    _chk(out, BF16, "out")
    check(_C.load().ld_pixel_shuffle2(x.data_ptr(), out.data_ptr(), F_, H, W, C4 // 4, _stream()), "ld_pixel_shuffle2")
    return out


def conv3x3_to_nchw16(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], out: Optional[torch.Tensor] = None):
    """3x3 / padding 1 convolution to 16 channels: x channels-last [F, H, W, Cin], w [16, Cin, 3, 3] (torch layout) ->
    NCHW bf16 [F, 16, H, W]."""
    _chk(x, BF16, "x"); _chk(w, BF16, "w")
    F_, H, W, Cin = x.shape
    if tuple(w.shape) != (16, Cin, 3, 3):
        raise ValueError(f"conv3x3_to_nchw16: w must be [16, {Cin}, 3, 3], got {tuple(w.shape)}")
    if bias is not None:
        _chk(bias, BF16, "bias")
    if out is None:
        out = torch.empty((F_, 16, H, W), dtype=BF16, device=x.device)
    _chk(out, BF16, "out")
    check(_C.load().ld_conv3x3_to_nchw16(x.data_ptr(), w.data_ptr(), _ptr(bias), out.data_ptr(), F_, H, W, Cin, _stream()),
          "ld_conv3x3_to_nchw16")
    return out


def conv3x3(x: torch.Tensor, w_taps: torch.Tensor, bias: Optional[torch.Tensor], *, gn=None, swish: bool = True,
            add: Optional[torch.Tensor] = None, implicit: bool = True, col: Optional[torch.Tensor] = None,
            max_col_bytes: int = 256 << 20):
    """3x3 / stride 1 / padding 1 convolution on channels-last frames on the tcgen05 GEMM.  w_taps: [Cout, 9*Cin] bf16 in
    (ky, kx, cin) order (`conv_weight_taps`); gn = (stats, gamma, beta, groups): GroupNorm (+ swish) of the input first;
    add: residual [F, H, W, Cout] added in the GEMM epilogue.
    implicit=True (Cin % 64 == 0): the activation is applied once and the GEMM gathers its A tiles by TMA in im2col mode —
    no im2col buffer.  implicit=False: explicit im2col (activation fused into the gather) + plain GEMM, in frame chunks
    so that the buffer stays under max_col_bytes."""
    _chk(x, BF16, "x")
    F_, H, W, Cin = x.shape
    Cout = w_taps.shape[0]
    out = torch.empty((F_, H, W, Cout), dtype=BF16, device=x.device)
    epi = EPI_BIAS if add is None else EPI_BIAS_ADD
    if implicit and Cin % 64 == 0:
        a = x if gn is None else groupnorm_apply(x, gn[0], gn[1], gn[2], gn[3], swish)
        gemm(a, w_taps, epilogue=epi, bias=bias, out=out.view(-1, Cout), conv=True,
             add2=None if add is None else add.view(-1, Cout))
        return out
    per_frame = H * W * 9 * Cin * 2
    chunk = max(1, min(F_, max_col_bytes // per_frame))
    if col is None or col.numel() < chunk * H * W * 9 * Cin:
        col = torch.empty((chunk * H * W * 9 * Cin,), dtype=BF16, device=x.device)
    for f0 in range(0, F_, chunk):
        f1 = min(F_, f0 + chunk)
        g = None if gn is None else (gn[0][f0:f1], gn[1], gn[2], gn[3])
        a = im2col3x3(x[f0:f1], col[: (f1 - f0) * H * W * 9 * Cin].view((f1 - f0) * H * W, 9 * Cin), gn=g, swish=swish)
        gemm(a, w_taps, epilogue=epi, bias=bias, out=out[f0:f1].view(-1, Cout),
             add2=None if add is None else add[f0:f1].view(-1, Cout))
    return out


def conv_weight_taps(w: torch.Tensor) -> torch.Tensor:
    """torch conv weight [Cout, Cin, 3, 3] -> the GEMM operand [Cout, (ky, kx, cin)] (one-time layout change at load)."""
    return w.detach().permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


_registered = False


def register_torch_ops() -> None:
    """Expose every compute entry point of the C-ABI as `torch.ops.landiff_b200.*` (called at import, below).  The ops
    are thin CUDA-dispatch-key wrappers over the same ctypes calls the modules in `dit.py` / `sampling.py` make; ops that
    write into caller-owned storage declare it in their schema (`Tensor(a!)`)."""
    global _registered
    if _registered:
        return
    lib = torch.library.Library("landiff_b200", "DEF")
    D = lib.define
    D("attention(Tensor q, Tensor k, Tensor v, int variant=0) -> Tensor")
    D("attention_lse(Tensor q, Tensor k, Tensor v, int variant=0) -> (Tensor, Tensor, Tensor)")
    D("attention_merge(Tensor(a!) o_acc, Tensor(b!) lse_acc, Tensor o_new, Tensor lse_new, Tensor(c!)? out_bf16, "
      "int batch, int heads, int nq) -> ()")
    D("linear(Tensor a, Tensor w, Tensor? bias, int epilogue=1) -> Tensor")
    D("linear_gated_residual(Tensor a, Tensor w, Tensor? bias, Tensor resid, Tensor gate_img, Tensor gate_txt, "
      "Tensor? add2, int rows_per_batch, int tok_offset, int text_len, int mod_batch_stride) -> Tensor")
    D("linear_qkv(Tensor a, Tensor w, Tensor bias, Tensor q_ln_w, Tensor q_ln_b, Tensor k_ln_w, Tensor k_ln_b, int heads, "
      "int rows_per_batch, float ln_eps=1e-6) -> (Tensor, Tensor, Tensor)")
    D("linear_bias_pos(Tensor a, Tensor w, Tensor bias, Tensor pos, Tensor(a!) out, int rows_per_batch, "
      "int out_rows_per_batch, int out_row_offset, int tok_offset, int text_len) -> ()")
    D("linear_unpatchify(Tensor a, Tensor w, Tensor bias, Tensor(a!) out, int rows_per_batch, int tok_offset, "
      "int text_len, int T, int Hp, int Wp, int C) -> ()")
    D("layernorm_modulate(Tensor x, Tensor w, Tensor b, float eps, Tensor shift_img, Tensor scale_img, Tensor shift_txt, "
      "Tensor scale_txt, int mod_batch_stride, int batch, int rows_per_batch, int tok_offset, int text_len) -> Tensor")
    D("final_norm_modulate(Tensor x, Tensor w1, Tensor b1, float eps1, Tensor w2, Tensor b2, float eps2, Tensor shift, "
      "Tensor scale, int mod_batch_stride, int batch, int rows_per_batch, int tok_offset, int text_len) -> Tensor")
    D("patchify(Tensor x, Tensor? sem, int g0=0, int n=-1) -> Tensor")
    D("small_linear(Tensor x, Tensor w, Tensor? bias, int act_in=0, int act_out=0, bool round_bf16=True) -> Tensor")
    D("small_linear_batched(Tensor x, Tensor[] ws, Tensor[] biases, int act_in=0, bool round_bf16=True) -> Tensor")
    D("timestep_embedding(Tensor t, int dim, float max_period=10000.0, bool round_bf16=True) -> Tensor")
    D("sampler_update(Tensor x, Tensor net_u, Tensor net_c, Tensor? old_den, Tensor? eps, float c_skip, "
      "float c_out, float cfg, float m1, float m2, float m3, float m4, float mn, int mode) -> (Tensor, Tensor)")

    D("nchw_to_nhwc(Tensor x) -> Tensor")
    D("groupnorm_stats(Tensor x, int groups=32, float eps=1e-6) -> Tensor")
    D("groupnorm_apply(Tensor x, Tensor stats, Tensor gamma, Tensor beta, int groups=32, bool swish=True) -> Tensor")
    D("conv3x3(Tensor x, Tensor w_taps, Tensor? bias, Tensor? add) -> Tensor")
    D("im2col3x3(Tensor x, Tensor? gn_stats, Tensor? gamma, Tensor? beta, int groups=32, bool swish=True) -> Tensor")
    D("pixel_shuffle2(Tensor x) -> Tensor")
    D("conv3x3_to_nchw16(Tensor x, Tensor w, Tensor? bias) -> Tensor")
    D("linear_bias_add(Tensor a, Tensor w, Tensor? bias, Tensor add) -> Tensor")
    D("unpatchify_blocks(Tensor[] blocks, int[] rows, int[] g0, int[] counts, Tensor(a!) out) -> ()")

    def _im2col3x3(x, gn_stats, gamma, beta, groups=32, swish=True):
        gn = None if gn_stats is None else (gn_stats, gamma, beta, groups)
        return im2col3x3(x, gn=gn, swish=swish)

    def _linear_bias_add(a, w, bias, add):
        return gemm(a, w, epilogue=EPI_BIAS_ADD, bias=bias, add2=add)

    def _attention(q, k, v, variant=0):
        return attention(q, k, v, variant=variant)

    def _attention_lse(q, k, v, variant=0):
        B, H, nq, _ = q.shape
        lse = torch.empty((B * H, nq), dtype=F32, device=q.device)
        of = torch.empty((B * H, nq, 64), dtype=F32, device=q.device)
        return attention(q, k, v, variant=variant, lse=lse, out_f32=of), lse, of

    def _attention_merge(o_acc, lse_acc, o_new, lse_new, out_bf16, batch, heads, nq):
        attention_merge(o_acc, lse_acc, o_new, lse_new, out_bf16, batch, heads, nq)

    def _linear(a, w, bias, epilogue=1):
        if epilogue not in (EPI_NONE, EPI_BIAS, EPI_BIAS_GELU):
            raise ValueError("landiff_b200::linear handles epilogues NONE / BIAS / BIAS_GELU; see the other linear_* ops")
        return gemm(a, w, epilogue=epilogue, bias=bias)

    def _linear_gated_residual(a, w, bias, resid, gate_img, gate_txt, add2, rows_per_batch, tok_offset, text_len,
                               mod_batch_stride):
        out = torch.empty_like(resid)
        return gemm(a, w, epilogue=EPI_GATED_RESID, bias=bias, out=out, rows_per_batch=rows_per_batch, tok_offset=tok_offset,
                    text_len=text_len, resid=resid, add2=add2, gate_img=gate_img, gate_txt=gate_txt,
                    mod_batch_stride=mod_batch_stride)

    def _linear_qkv(a, w, bias, q_ln_w, q_ln_b, k_ln_w, k_ln_b, heads, rows_per_batch, ln_eps=1e-6):
        B = a.shape[0] // rows_per_batch
        q = torch.empty((B, heads, rows_per_batch, 64), dtype=BF16, device=a.device)
        k, v = torch.empty_like(q), torch.empty_like(q)
        gemm(a, w, epilogue=EPI_QKV, bias=bias, rows_per_batch=rows_per_batch, qkv=(q, k, v),
             qk_ln=(q_ln_w, q_ln_b, k_ln_w, k_ln_b), ln_eps=ln_eps, heads=heads)
        return q, k, v

    def _linear_bias_pos(a, w, bias, pos, out, rows_per_batch, out_rows_per_batch, out_row_offset, tok_offset, text_len):
        gemm(a, w, epilogue=EPI_BIAS_POS, bias=bias, out=out, rows_per_batch=rows_per_batch,
             out_rows_per_batch=out_rows_per_batch, out_row_offset=out_row_offset, tok_offset=tok_offset, text_len=text_len,
             pos=pos)

    def _linear_unpatchify(a, w, bias, out, rows_per_batch, tok_offset, text_len, T, Hp, Wp, C_):
        gemm(a, w, epilogue=EPI_UNPATCHIFY, bias=bias, out=out, rows_per_batch=rows_per_batch, tok_offset=tok_offset,
             text_len=text_len, patch_grid=(T, Hp, Wp, C_))

    def _small_linear_batched(x, ws, biases, act_in=0, round_bf16=True):
        return small_linear_batched(x, pointer_table(ws, x.device), pointer_table(biases, x.device), len(ws), ws[0].shape[0],
                                    act_in=act_in, round_bf16=round_bf16)

    def _patchify(x, sem, g0=0, n=-1):
        return patchify(x, sem, g0=g0, n=None if n < 0 else n)

    def _sampler_update(x, net_u, net_c, old_den, eps, c_skip, c_out, cfg, m1, m2, m3, m4, mn, mode):
        return sampler_update(x, net_u, net_c, old_den, eps, c_skip=c_skip, c_out=c_out, cfg=cfg, m1=m1, m2=m2, m3=m3,
                              m4=m4, mn=mn, mode=mode, net_dtype=net_u.dtype)

    for name, fn in (("attention", _attention), ("attention_lse", _attention_lse), ("attention_merge", _attention_merge),
                     ("linear", _linear), ("linear_gated_residual", _linear_gated_residual), ("linear_qkv", _linear_qkv),
                     ("linear_bias_pos", _linear_bias_pos), ("linear_unpatchify", _linear_unpatchify),
                     ("layernorm_modulate", layernorm_modulate), ("final_norm_modulate", final_norm_modulate),
                     ("patchify", _patchify), ("small_linear", small_linear), ("small_linear_batched", _small_linear_batched),
                     ("timestep_embedding", timestep_embedding),
                     ("sampler_update", _sampler_update), ("nchw_to_nhwc", lambda x: nchw_to_nhwc(x)),
                     ("groupnorm_stats", groupnorm_stats), ("im2col3x3", _im2col3x3),
                     ("groupnorm_apply", lambda x, st, gm, bt, groups=32, swish=True: groupnorm_apply(x, st, gm, bt, groups, swish)),
                     ("conv3x3", lambda x, w_taps, bias, add: conv3x3(x, w_taps, bias, add=add)),
                     ("pixel_shuffle2", lambda x: pixel_shuffle2(x)),
                     ("conv3x3_to_nchw16", lambda x, w, bias: conv3x3_to_nchw16(x, w, bias)),
                     ("linear_bias_add", _linear_bias_add),
                     ("unpatchify_blocks", lambda blocks, rows, g0, counts, out:
                      (unpatchify_blocks(list(zip(blocks, rows, g0, counts)), out), None)[1])):
        lib.impl(name, fn, "CUDA")
    register_torch_ops._lib = lib  # keep alive
    _registered = True


TORCH_OPS = ("attention", "attention_lse", "attention_merge", "linear", "linear_gated_residual", "linear_qkv",
             "linear_bias_pos", "linear_unpatchify", "layernorm_modulate", "final_norm_modulate", "patchify", "small_linear",
             "small_linear_batched", "timestep_embedding", "sampler_update", "nchw_to_nhwc", "groupnorm_stats", "groupnorm_apply", "conv3x3", "im2col3x3",
             "pixel_shuffle2", "conv3x3_to_nchw16", "linear_bias_add", "unpatchify_blocks")

register_torch_ops()
