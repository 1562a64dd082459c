"""ctypes binding of the C-ABI in include/landiff_b200.h.

The shared library is built in-tree by landiff_b200.build (nvcc, sm_100a).  There is NO fallback: if the
library is missing it is built; if it cannot be built or loaded this module raises.  Every compute entry
point fails with LD_ERR_DEVICE on a machine without an sm_100 GPU.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from . import build as _build

LD_OK = 0

EPI_NONE, EPI_BIAS, EPI_BIAS_GELU, EPI_GATED_RESID, EPI_QKV, EPI_BIAS_POS, EPI_UNPATCHIFY, EPI_BIAS_ADD = range(8)


class GemmArgs(C.Structure):
    """Mirror of `ld_gemm_args` (include/landiff_b200.h)."""

    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("epilogue", C.c_int32),
        ("A", C.c_void_p), ("W", C.c_void_p), ("bias", C.c_void_p), ("out", C.c_void_p),
        ("ld_out", C.c_int64),
        ("rows_per_batch", C.c_int32), ("out_rows_per_batch", C.c_int32), ("out_row_offset", C.c_int32),
        ("tok_offset", C.c_int32), ("text_len", C.c_int32),
        ("resid", C.c_void_p), ("add2", C.c_void_p), ("gate_img", C.c_void_p), ("gate_txt", C.c_void_p),
        ("mod_batch_stride", C.c_int64),
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p),
        ("q_ln_w", C.c_void_p), ("q_ln_b", C.c_void_p), ("k_ln_w", C.c_void_p), ("k_ln_b", C.c_void_p),
        ("ln_eps", C.c_float), ("q_scale", C.c_float),
        ("heads", C.c_int32), ("qkv_rows", C.c_int32), ("qkv_row_offset", C.c_int32),
        ("pos", C.c_void_p),
        ("T", C.c_int32), ("Hp", C.c_int32), ("Wp", C.c_int32), ("C", C.c_int32),
        ("resid_f32", C.c_int32), ("out_f32", C.c_int32),
        ("conv_F", C.c_int32), ("conv_H", C.c_int32), ("conv_W", C.c_int32), ("conv_C", C.c_int32),
    ]


class TokenBlocks(C.Structure):
    """Mirror of `ld_token_blocks` (include/landiff_b200.h)."""

    _fields_ = [("ptr", C.c_void_p * 16), ("row", C.c_int32 * 16), ("g0", C.c_int32 * 16), ("count", C.c_int32 * 16),
                ("n", C.c_int32)]


class KvShard(C.Structure):
    """Mirror of `ld_kv_shard` (include/landiff_b200.h)."""

    _fields_ = [("k", C.c_void_p), ("v", C.c_void_p), ("nkv", C.c_int32), ("kv_rows", C.c_int32),
                ("ready_flag", C.c_void_p), ("ready_value", C.c_uint32)]


# name -> (restype, argtypes); the CPU test-suite checks every one of these is exported.
_vp, _i, _f, _i64, _fp = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_void_p
SIGNATURES = {
    "ld_last_error": (C.c_char_p, []),
    "ld_abi_version": (C.c_int, []),
    "ld_struct_sizes": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ld_device_check": (C.c_int, [C.POINTER(C.c_int)]),
    "ld_gemm_bf16": (C.c_int, [C.POINTER(GemmArgs), _vp]),
    "ld_attention_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "ld_attention_shards_bf16": (C.c_int, [_vp, C.POINTER(KvShard), _i, _vp, _fp, _fp, _i, _i, _i, _i, _i, _vp]),
    "ld_attention_workspace_bytes": (C.c_size_t, [C.POINTER(KvShard), _i, _i, _i, _i]),
    "ld_attention_shards_ws_bf16": (C.c_int, [_vp, C.POINTER(KvShard), _i, _vp, _fp, _fp, _i, _i, _i, _i, _i, _vp, C.c_size_t,
                                              _vp]),
    "ld_attention_status": (C.c_int, [C.POINTER(C.c_uint), _i]),
    "ld_attention_merge": (C.c_int, [_fp, _fp, _fp, _fp, _vp, _i, _i, _i, _vp]),
    "ld_nchw_to_nhwc": (C.c_int, [_vp, _i, _vp, _i, _i, _i, _vp]),
    "ld_groupnorm_stats": (C.c_int, [_vp, _fp, _fp, _i, _i, _i, _i, C.c_float, _vp]),
    "ld_groupnorm_apply": (C.c_int, [_vp, _vp, _fp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ld_im2col3x3": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _fp, _vp, _vp, _i, _i, _vp]),
    "ld_pixel_shuffle2": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "ld_conv3x3_to_nchw16": (C.c_int, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "ld_unpatchify_blocks": (C.c_int, [C.POINTER(TokenBlocks), _vp, _i, _i, _i, _i, _vp]),
    "ld_ipc_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]),
    "ld_ipc_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "ld_ipc_close": (C.c_int, [_vp]),
    "ld_ipc_free": (C.c_int, [_vp]),
    "ld_copy_async": (C.c_int, [_vp, _vp, C.c_size_t, _vp]),
    "ld_stream_write_u32": (C.c_int, [_vp, C.c_uint, _vp]),
    "ld_stream_wait_geq_u32": (C.c_int, [_vp, C.c_uint, _vp]),
    "ld_layernorm_modulate": (C.c_int, [_vp, _i, _vp, _vp, _vp, _f, _fp, _fp, _fp, _fp, _i64, _i, _i, _i, _i, _i, _vp]),
    "ld_final_norm_modulate": (C.c_int, [_vp, _i, _vp, _vp, _vp, _f, _vp, _vp, _f, _fp, _fp, _i64, _i, _i, _i, _i, _i, _vp]),
    "ld_patchify": (C.c_int, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "ld_small_linear": (C.c_int, [_fp, _vp, _vp, _fp, _i, _i, _i, _i, _i, _i, _vp]),
    "ld_small_linear_batched": (C.c_int, [_fp, _vp, _vp, _fp, _i, _i, _i, _i, _i, _i, _vp]),
    "ld_timestep_embedding": (C.c_int, [_fp, _fp, _i, _i, _f, _i, _vp]),
    "ld_sampler_update": (C.c_int, [_fp, _vp, _vp, _fp, _fp, _fp, _fp, _i64] + [_f] * 8 + [_i, _i, _vp]),
}

_lib = None


def lib_path() -> Path:
    return _build.LIB_PATH


ABI_VERSION = 7   # must equal ld_abi_version() of the loaded library (bumped whenever a signature or struct changes)


def load() -> C.CDLL:
    """Build (if stale/missing) and load the library; bind signatures.  Raises on any failure.  The build step is a
    cheap source-hash stamp check when the library is current; a library that was built from other sources (a stale
    git-ignored .so carried over from an older snapshot) is rebuilt, and one that cannot be rebuilt (no nvcc) but
    reports another ABI version is refused."""
    global _lib
    if _lib is not None:
        return _lib
    try:
        path = _build.build()
    except Exception:
        path = _build.LIB_PATH
        if not path.exists():
            raise
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    got = lib.ld_abi_version()
    if got != ABI_VERSION:
        raise LanDiffB200Error(f"{path} reports ABI version {got}, the Python binding expects {ABI_VERSION}: "
                               "stale library — rebuild with `python -m landiff_b200.build --force`")
    sz = [C.c_int(0) for _ in range(3)]
    lib.ld_struct_sizes(*[C.byref(v) for v in sz])
    for name, got_c, mirror in (("ld_gemm_args", sz[0].value, GemmArgs), ("ld_kv_shard", sz[1].value, KvShard),
                                ("ld_token_blocks", sz[2].value, TokenBlocks)):
        if got_c != C.sizeof(mirror):
            raise LanDiffB200Error(f"{name}: the library's struct is {got_c} bytes, the Python mirror {C.sizeof(mirror)} — "
                                   "include/landiff_b200.h and landiff_b200/_C.py have drifted apart")
    _lib = lib
    return lib


class LanDiffB200Error(RuntimeError):
    pass


LAUNCHES = [0]  # KERNELS launched through the C-ABI (every compute entry point launches exactly one kernel)
# entry points that launch no kernel: queries, allocation, copy-engine copies and stream memory operations
_NO_KERNEL = {"ld_device_check", "ld_ipc_alloc", "ld_ipc_open", "ld_ipc_close", "ld_ipc_free", "ld_copy_async",
              "ld_stream_write_u32", "ld_stream_wait_geq_u32", "ld_attention_status", "ld_attention_workspace_bytes", "ld_struct_sizes", "w"}


def check(rc: int, what: str) -> None:
    if rc == LD_OK and what not in _NO_KERNEL:
        LAUNCHES[0] += 1
    if rc != LD_OK:
        msg = load().ld_last_error().decode(errors="replace")
        raise LanDiffB200Error(f"{what} failed (code {rc}): {msg}")
