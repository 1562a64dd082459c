/*
 * landiff_b200 — C-ABI of the B200-native (sm_100a) DiT hot path for LanDiff's diffusion stage.
 *
 * Plain C: pointers are DEVICE pointers unless stated otherwise, sizes are ints, `stream` is a cudaStream_t
 * passed as void*.  Every entry point returns LD_OK (0) or a negative error code; ld_last_error() gives the
 * thread-local message.  There is no CPU fallback: on a machine without an sm_100 device every compute entry
 * point fails with LD_ERR_DEVICE.
 *
 * The reference (pure Python, /root/reference) has no FFI; each entry point below names the reference
 * PyTorch expression it replaces (file:line relative to the reference root).  The Python host in
 * landiff_b200/ binds these with ctypes and registers them as torch.library custom ops; INTEGRATION.md shows
 * the binding a reference maintainer would add.
 *
 * Layout conventions: activations are row-major bf16 [rows, features]; a "row" is one token of one sample,
 * rows of sample b occupy [b*rows_per_batch, (b+1)*rows_per_batch).  Token t of a sample is a TEXT token iff
 * tok_offset + t < text_len (dit_video_concat.py:549-550).  Linear weights keep the nn.Linear layout
 * [out_features, in_features] bf16 (no copies, no re-packing).
 */
#ifndef LANDIFF_B200_H
#define LANDIFF_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LD_OK 0
#define LD_ERR_ARG (-1)    /* bad shape / alignment / null pointer */
#define LD_ERR_CUDA (-2)   /* CUDA runtime / driver error */
#define LD_ERR_DEVICE (-3) /* no sm_100 device */

const char* ld_last_error(void);
int ld_abi_version(void);
/* sizeof(ld_gemm_args), sizeof(ld_kv_shard), sizeof(ld_token_blocks): a binding checks its mirrors against these at load */
int ld_struct_sizes(int* gemm_args, int* kv_shard, int* token_blocks);
/* 0 iff the current device is compute capability 10.x; fills sm_count if non-null. */
int ld_device_check(int* sm_count);

/* ---- dense contractions: out = epilogue(A[M,K] @ W[N,K]^T), tcgen05/TMEM tiles fed by TMA ------------- */
enum ld_epilogue {
  LD_EPI_NONE = 0,        /* out = acc                      control zero_linears, dit_video_concat.py:1234-1237 */
  LD_EPI_BIAS = 1,        /* out = acc + bias               text_proj :55-58 (with row remap into hidden)       */
  LD_EPI_BIAS_GELU = 2,   /* out = gelu_tanh(acc + bias)    SAT MLP dense_h_to_4h + activation, :612, :731-733  */
  LD_EPI_GATED_RESID = 3, /* out = resid + gate[seg]*(acc+bias) (+ add2)   :593-598, :619-624, :1357-1370       */
  LD_EPI_QKV = 4,         /* split Q|K|V, per-head LayerNorm(64) on Q,K, head-major store   :636-664 + SAT      */
  LD_EPI_BIAS_POS = 5,    /* out = acc + bias + pos[token]  patch-embed conv as GEMM :47-62, pos add :227-231   */
  LD_EPI_UNPATCHIFY = 6,  /* out[b,t,c,2h+p,2w+q] = acc + bias   final linear + unpatchify :392-410, :453-456   */
  LD_EPI_BIAS_ADD = 7     /* out = acc + bias + add2 (bf16, indexed like out)   ResnetBlock `x + h` of the semantic
                             conditioner's conv decoder, semantic_models/modules/vq_gan_blocks.py:126-147           */
};

typedef struct ld_gemm_args {
  int32_t M, N, K;              /* K % 64 == 0; N % tile_n == 0 (tile_n: 192 for N%192==0, else 128, else 64) */
  int32_t epilogue;             /* enum ld_epilogue */
  const void* A;                /* [M, K] bf16 row-major */
  const void* W;                /* [N, K] bf16 row-major (nn.Linear weight) */
  const void* bias;             /* [N] bf16 or NULL */
  void* out;                    /* bf16; row stride ld_out elements (ignored by QKV) */
  int64_t ld_out;
  /* row bookkeeping: A row r belongs to sample b = r / rows_per_batch, token tok_offset + r % rows_per_batch;
     it is written to out row b*out_rows_per_batch + out_row_offset + r % rows_per_batch */
  int32_t rows_per_batch, out_rows_per_batch, out_row_offset;
  int32_t tok_offset, text_len;
  /* GATED_RESID: resid/add2 are indexed like out; gate_* point at sample 0's fp32 [N] vectors */
  const void* resid;
  const void* add2;
  const float* gate_img;
  const float* gate_txt;
  int64_t mod_batch_stride;     /* floats between consecutive samples' modulation vectors */
  /* QKV: N = 3*heads*64.  q/k/v are [B, heads, qkv_rows, 64] bf16; token t lands on row qkv_row_offset + t.
     Q is additionally multiplied by q_scale (softmax scale * log2 e) after its LayerNorm. */
  void* q; void* k; void* v;
  const void* q_ln_w; const void* q_ln_b; const void* k_ln_w; const void* k_ln_b; /* bf16 [64] */
  float ln_eps, q_scale;
  int32_t heads, qkv_rows, qkv_row_offset;
  /* BIAS_POS: pos is bf16 [>= text_len + image tokens, N]; the row added is pos[tok_offset + t] */
  const void* pos;
  /* UNPATCHIFY: out is bf16 [B, T, C, 2*Hp, 2*Wp]; image token index g = tok_offset + t - text_len */
  int32_t T, Hp, Wp, C;
  /* residual-stream dtype (GATED_RESID: resid/out, BIAS_POS: out): 0 = bf16 (the reference's rounding points),
     1 = fp32 (the main net keeps its 30-layer residual stream in fp32; add2 stays bf16) */
  int32_t resid_f32, out_f32;
  /* implicit 3x3 convolution (conv_W > 0): A is a channels-last activation tensor [conv_F, conv_H, conv_W, conv_C] bf16
     (conv_C % 64 == 0), W is [N, 9*conv_C] in (ky, kx, cin) order, K = 9*conv_C, M = conv_F*conv_H*conv_W output positions
     in raster order; stride 1, zero padding 1.  The A tiles are gathered by TMA in im2col mode (no im2col buffer). */
  int32_t conv_F, conv_H, conv_W, conv_C;
} ld_gemm_args;

int ld_gemm_bf16(const ld_gemm_args* args, void* stream);

/* ---- full (non-causal) self-attention, head_dim 64, flash-style online softmax on tcgen05 ------------- */
/* q: [BH, q_rows, 64], k/v: [BH, kv_rows, 64] bf16; scores are scaled by 1/sqrt(64) inside the kernel.
   Uses q rows [0,nq) and kv rows [0,nkv).  out: bf16 [B, nq, heads*64] (token-major, ready for `dense`).
   variant: 0 = default (one fixed reference maximum per row, 4 of 16 exponential pairs on the FMA pipe, row sums on the
   tensor core; CTAs whose reference overflowed re-run through the exact path inside the same launch), 1 = exact path only
   (per-block maxima, lazy rescaling), 2 / 3 / 4 = default with 0 / 5 / 6 of 16 pairs on the FMA pipe, 5 = default with a
   truncating bf16 pack (tuning variants), 6 = default without the tail split.
   Tail split: the query blocks of the last, partly empty wave of the grid are each served by several CTAs that take a share
   of the keys and write fp32 partial results, merged by a second small kernel.  It needs a workspace, so it is only
   available through ld_attention_shards_ws_bf16 (LD_ATTN_SPLIT=0 disables it; it is also off when lse is requested).
   If lse != NULL also writes fp32 log2-sum-exp [BH, nq] and, when out_f32 != NULL, the normalised fp32 output
   [BH, nq, 64].  Replaces SAT attention_fn_default -> F.scaled_dot_product_attention reached through
   dit_video_concat.py:655-664. */
int ld_attention_bf16(const void* q, const void* k, const void* v, void* out, float* lse, float* out_f32,
                      int batch, int heads, int nq, int q_rows, int nkv, int kv_rows, int variant, void* stream);

/* The same attention over K/V given as 1..4 SHARDS of the key sequence (ring sequence parallelism: the local shard and
   the shards the peers' copy engines deliver into IPC-mapped buffers, csrc/peer_ring.cu).  ONE launch walks all shards,
   accumulating in TMEM.  A shard with ready_flag != NULL is still in flight when the kernel starts: the kernel's TMA
   producer polls the 32-bit device word until (int32)(*ready_flag - ready_value) >= 0 (acquire, system scope) before its
   first load from that shard.  A wait that lasts longer than ~2 s gives up and raises bit 0 of the status word
   (ld_attention_status) instead of hanging the GPU.  The reference has no sequence parallelism. */
typedef struct ld_kv_shard {
  const void* k;              /* [BH, kv_rows, 64] bf16 */
  const void* v;
  int32_t nkv, kv_rows;
  const uint32_t* ready_flag; /* device pointer or NULL */
  uint32_t ready_value;
} ld_kv_shard;
int ld_attention_shards_bf16(const void* q, const ld_kv_shard* shards, int n_shards, void* out, float* lse, float* out_f32,
                             int batch, int heads, int nq, int q_rows, int variant, void* stream);

/* The same with a caller-owned workspace for the tail split (see ld_attention_bf16): ld_attention_workspace_bytes says how
   much this problem wants (0: nothing to split); with workspace == NULL or too small the launch uses one CTA per query
   block.  The workspace must stay untouched until the launch has finished on `stream`. */
size_t ld_attention_workspace_bytes(const ld_kv_shard* shards, int n_shards, int batch, int heads, int nq);
int ld_attention_shards_ws_bf16(const void* q, const ld_kv_shard* shards, int n_shards, void* out, float* lse, float* out_f32,
                                int batch, int heads, int nq, int q_rows, int variant, void* workspace,
                                size_t workspace_bytes, void* stream);
/* copies the per-device status word of the shard waits to *host_out (synchronising); reset != 0 clears it */
int ld_attention_status(unsigned int* host_out, int reset);

/* merge two partial attention results (ring hop):  (o_acc, lse_acc) <- merge((o_acc, lse_acc), (o_new, lse_new)).
   o_*: fp32 [BH, nq, 64]; lse_*: fp32 [BH, nq].  If out_bf16 != NULL also writes bf16 [B, nq, heads*64]. */
int ld_attention_merge(float* o_acc, float* lse_acc, const float* o_new, const float* lse_new, void* out_bf16,
                       int batch, int heads, int nq, void* stream);

/* ---- copy-engine ring transport (sequence-parallel attention, csrc/peer_ring.cu) ------------------------- */
/* K|V shards rotate between the GPUs of one NVLink domain as cudaMemcpyAsync peer copies (no SMs) into buffers the
   receiver exports with CUDA IPC; cross-process ordering by 32-bit stream memory operations.  Replaces the NCCL
   send/recv hop of landiff_b200/parallel.py; the reference has no sequence parallelism.
   ld_ipc_alloc: cudaMalloc + zero + export (64-byte cudaIpcMemHandle_t).  ld_ipc_open maps a peer's export into this
   process (lazy peer access).  ld_stream_wait_geq_u32 waits until (int32)(*addr - value) >= 0. */
int ld_ipc_alloc(size_t bytes, void** dev_ptr, unsigned char handle_out[64]);
int ld_ipc_open(const unsigned char handle[64], void** dev_ptr);
int ld_ipc_close(void* dev_ptr);
int ld_ipc_free(void* dev_ptr);
int ld_copy_async(void* dst, const void* src, size_t bytes, void* stream);
int ld_stream_write_u32(void* dev_addr, unsigned int value, void* stream);
int ld_stream_wait_geq_u32(void* dev_addr, unsigned int value, void* stream);

/* ---- fused memory-bound row kernels ------------------------------------------------------------------- */
/* out = LayerNorm(x; w, b, eps) * (1 + scale[seg]) + shift[seg]   (dit_video_concat.py:577-586, 601-611, :388)
   x: bf16 or fp32 (x_is_f32) [B*rows_per_batch, D]; out: bf16; w,b: bf16 [D]; shift_x / scale_x: fp32 [D] of sample 0.
   D % 8 == 0, D <= 2048 */
int ld_layernorm_modulate(const void* x, int x_is_f32, void* out, const void* w, const void* b, float eps,
                          const float* shift_img, const float* scale_img, const float* shift_txt,
                          const float* scale_txt, int64_t mod_batch_stride, int batch, int rows_per_batch,
                          int tok_offset, int text_len, int D, void* stream);

/* final layer front half: y = LN2(LN1(x[image rows]); eps2) * (1 + scale) + shift  -> bf16 [B*n_img, D]
   (SAT final_layernorm + FinalLayerMixin.final_forward, dit_video_concat.py:442-452).  x: [B, rows_per_batch, D];
   image rows are those with tok_offset + t >= text_len.  x is bf16 or fp32 (x_is_f32). */
int ld_final_norm_modulate(const void* x, int x_is_f32, void* out, const void* w1, const void* b1, float eps1, const void* w2,
                           const void* b2, float eps2, const float* shift, const float* scale,
                           int64_t mod_batch_stride, int batch, int rows_per_batch, int tok_offset, int text_len,
                           int D, void* stream);

/* im2col for the 2x2/stride-2 patch conv: cols[b*n_img + g, c*4 + p*2 + q] = bf16(x[b,t,c,2h+p,2w+q] (+ sem[t,c,..]))
   x: fp32 or bf16 [B,T,C,2Hp,2Wp] (x_is_f32), sem: bf16/fp32 [1,T,C,2Hp,2Wp] or NULL (same dtype flag sem_is_f32).
   Only image tokens g in [g0, g0+n) are produced (sequence shard).  dit_video_concat.py:47-54, :991 */
int ld_patchify(const void* x, int x_is_f32, const void* sem, int sem_is_f32, void* cols, int batch, int T, int C,
                int Hp, int Wp, int g0, int n, void* stream);

/* y[b, n] = act_out( sum_k act_in(x[b,k]) * W[n,k] + bias[n] ), tiny-batch GEMV (B <= 8), fp32 in/out, bf16 W.
   act codes: 0 none, 1 SiLU.  time_embed :764-768, adaLN_modulation :510-515/:555, final adaLN :434-436 */
int ld_small_linear(const float* x, const void* W, const void* bias, float* y, int batch, int N, int K, int act_in,
                    int act_out, int round_bf16, void* stream);

/* the same GEMV for `layers` weight matrices of identical shape that share the input x, ONE launch:
   y[l, b, n] = sum_k act_in(x[b,k]) * Ws[l][n,k] + biases[l][n].  Ws / biases: DEVICE arrays of `layers` device pointers
   (biases[l] may be NULL).  All adaLN_modulation projections of a network in one pass (dit_video_concat.py:510-515/:555
   evaluates them layer by layer from the same time embedding). */
int ld_small_linear_batched(const float* x, const void* const* Ws, const void* const* biases, float* y, int layers, int batch,
                            int N, int K, int act_in, int round_bf16, void* stream);

/* sinusoidal timestep embedding, [cos | sin] (sgm/modules/diffusionmodules/util.py:207-233) -> fp32 [B, dim] */
int ld_timestep_embedding(const float* t, float* out, int batch, int dim, float max_period, int round_bf16,
                          void* stream);

/* Token-major network outputs -> latent layout.  Under CFG / sequence parallelism every rank produces the final linear's
   output for its own (batch row, image-token shard) as a contiguous bf16 [count, 64] block (LD_EPI_BIAS), the blocks travel
   to every rank by copy-engine peer copies (landiff_b200/dma_ring.py), and this kernel scatters up to 16 of them into
   out [rows, T, C, 2*Hp, 2*Wp]: block b covers image tokens [g0, g0 + count) of output row `row` (unpatchify,
   dit_video_concat.py:392-410).  Replaces the zero-fill + mask + all-reduce assembly; the reference has no parallelism. */
typedef struct ld_token_blocks {
  const void* ptr[16];
  int32_t row[16], g0[16], count[16];
  int32_t n;
} ld_token_blocks;
int ld_unpatchify_blocks(const ld_token_blocks* blocks, void* out, int T, int Hp, int Wp, int C, void* stream);

/* fused denoiser scaling + CFG + DPM-Solver++(2M) SDE update, all fp32, n elements:
     den_u = c_out*net_u + c_skip*x ; den_c likewise        (denoiser.py:38-41, denoiser_scaling.py:62-70)
     den   = den_u + cfg*(den_c - den_u)                     (guiders.py:75-79, sampling_utils.py:8-13)
     mode 0 (first step):  x' = m1*x - m2*den + mn*eps                        (sampling.py:771-774)
     mode 1 (middle):      x' = m1*x - m2*(m3*den - m4*old) + mn*eps          (sampling.py:776-781)
     mode 2 (last):        x' = den                                           (sampling.py:750-751)
   net_u/net_c: bf16 network outputs (net_is_f32 = 0) or fp32 rows (net_is_f32 = 1, e.g. already-denoised rows with
   c_skip = 0, c_out = 1 when the reference DiscreteDenoiser stays in the loop); x, old, eps, x_out, den_out: fp32. */
int ld_sampler_update(const float* x, const void* net_u, const void* net_c, const float* old_den,
                      const float* eps, float* x_out, float* den_out, int64_t n, float c_skip, float c_out,
                      float cfg, float m1, float m2, float m3, float m4, float mn, int mode, int net_is_f32,
                      void* stream);

/* ---- semantic conditioner, upsample path ------------------------------------------------
   The conv decoder between the semantic tokenizer's features and the control network's latent add
   (landiff/diffusion/semantic_models/condition.py:86-137; modules/vq_gan_blocks.py:30-147, 480-606).  Activations are
   channels-last bf16 [frames, H, W, C]; a 3x3 convolution = ld_im2col3x3 (GroupNorm + swish of the input applied on the
   fly) + ld_gemm_bf16 with weights reordered to [Cout, (ky, kx, cin)] and epilogue LD_EPI_BIAS / LD_EPI_BIAS_ADD. */

/* out = swish(GroupNorm(x)) (swish optional) on channels-last frames [frames, P, C] bf16 with ld_groupnorm_stats' (mean, rstd)
   and gamma/beta bf16 [C]: the activation in front of every convolution (vq_gan_blocks.py:128-140, 599-604) */
int ld_groupnorm_apply(const void* x, void* out, const float* stats, const void* gamma, const void* beta, int frames, int P,
                       int C, int groups, int swish, void* stream);

/* x [frames, C, P] (NCHW with P = H*W; bf16 or fp32) -> out [frames, P, C] bf16   (condition.py:104-107 input cast + layout) */
int ld_nchw_to_nhwc(const void* x, int x_is_f32, void* out, int frames, int C, int P, void* stream);

/* GroupNorm statistics (vq_gan_blocks.py:35-38: 32 groups, eps 1e-6) of channels-last frames x [frames, P, C] bf16:
   stats [frames, groups, 2] fp32 = (mean, rstd), two passes (sum, then centred squares) reduced in a fixed order, so the
   result is bit-reproducible.  scratch: fp32 [frames * ceil(P / 64) * groups]. */
int ld_groupnorm_stats(const void* x, float* stats, float* scratch, int frames, int P, int C, int groups, float eps,
                       void* stream);

/* im2col of a 3x3 / stride 1 / padding 1 convolution: out [(frame, y, x), (ky, kx, c)] bf16.  gn_stats != NULL: the input is
   first normalised with ld_groupnorm_stats' (mean, rstd), gamma/beta (bf16 [C]) and passed through swish when swish != 0
   (norm -> nonlinearity -> conv, vq_gan_blocks.py:128-140, 599-604); padding taps are zeros of the ACTIVATED tensor. */
int ld_im2col3x3(const void* x, void* out, int frames, int H, int W, int C, const float* gn_stats, const void* gamma,
                 const void* beta, int groups, int swish, void* stream);

/* torch.nn.PixelShuffle(2) on channels-last frames: x [frames, H, W, 4*C_out] -> out [frames, 2H, 2W, C_out]
   (vq_gan_blocks.py:47-48, 63-64) */
int ld_pixel_shuffle2(const void* x, void* out, int frames, int H, int W, int C_out, void* stream);

/* direct 3x3 convolution to 16 channels: x channels-last [frames, H, W, Cin] bf16, w bf16 in torch layout [16, Cin, 3, 3],
   bias bf16 [16] or NULL, out bf16 NCHW [frames, 16, H, W] — SemanticCond.conv_out, condition.py:49-56, 132-136 */
int ld_conv3x3_to_nchw16(const void* x, const void* w, const void* bias, void* out, int frames, int H, int W, int Cin,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif
