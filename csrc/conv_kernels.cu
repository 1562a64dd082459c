// Kernels of the semantic conditioner's upsample path: the VQGAN-style conv decoder that turns
// the semantic tokenizer's [B*T, 768, H/16, W/16] features into the [B, T, 16, H/8, W/8] tensor the control network adds
// to its latent (reference: landiff/diffusion/semantic_models/condition.py:86-137, modules/vq_gan_blocks.py:30-147,
// 480-606).  It runs ONCE per video, not per sampler step, so the design goal is "every FLOP on the tcgen05 GEMM of
// gemm_tcgen05.cu, every byte moved once": activations are channels-last bf16 [frame, y, x, c]; a 3x3 convolution is an
// im2col gather ([positions, 9*C], taps-major) with the preceding GroupNorm + swish applied on the fly, followed by the
// GEMM with bias (+ residual) epilogue; GroupNorm statistics are two small reductions (sum, then centred squares, like
// torch's two-pass kernel); PixelShuffle(2) and the NCHW <-> channels-last changes are index kernels.  The last
// convolution (64 -> 16 channels, zero-initialised in the reference) is too narrow for a 64-wide MMA tile and is done
// directly, writing the NCHW layout `patchify` reads.
#include "host_util.h"
#include "ptx.cuh"

namespace ld {

using bf16 = __nv_bfloat16;

__device__ __forceinline__ void cv_unpack8(const uint4& v, float* f) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 cv_pack8(const float* f) {
  uint4 v;
  v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
  v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
  return v;
}

// ---- [F, C, P] (NCHW, bf16 or fp32) -> [F, P, C] bf16 ------------------------------------------------------------------
template <typename TX>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const TX* __restrict__ x, bf16* __restrict__ out, int C, int P) {
  __shared__ float tile[32][33];
  const int f = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const TX* xf = x + (int64_t)f * C * P;
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, p = p0 + tx;
    tile[j][tx] = (c < C && p < P) ? (float)xf[(int64_t)c * P + p] : 0.f;
  }
  __syncthreads();
  bf16* of = out + (int64_t)f * P * C;
  for (int j = ty; j < 32; j += 8) {
    const int p = p0 + j, c = c0 + tx;
    if (p < P && c < C) of[(int64_t)p * C + c] = __float2bfloat16_rn(tile[tx][j]);
  }
}

// ---- GroupNorm statistics over channels-last frames -------------------------------------------------------------------
// x: [F, P, C] -> stats [F, G, 2] = (mean, rstd), two passes like torch's kernel (sum, then centred squares) and
// DETERMINISTIC: every CTA reduces its 256-row chunk in a fixed order into partial[f, chunk, g], a one-CTA-per-frame
// kernel adds the chunks in order.  (vq_gan_blocks.py:35-38: 32 groups, eps 1e-6)
constexpr int kGnRows = 64;

template <int PASS>
__global__ void __launch_bounds__(256) gn_partial_kernel(const bf16* __restrict__ x, const float* __restrict__ stats,
                                                         float* __restrict__ partial, int P, int C, int G) {
  extern __shared__ float part_sm[];   // [rstep][C]
  const int f = blockIdx.y;
  const int nvec = C >> 3;
  const int cpg = C / G;
  const int v = threadIdx.x % nvec, rl = threadIdx.x / nvec, rstep = blockDim.x / nvec;
  float mean[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) mean[i] = PASS == 1 ? stats[(f * G + (8 * v + i) / cpg) * 2] : 0.f;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int p0 = blockIdx.x * kGnRows;
  const int p1 = min(P, p0 + kGnRows);
  const uint4* xf = reinterpret_cast<const uint4*>(x + (int64_t)f * P * C);
  for (int pb = p0 + rl; pb < p1; pb += 4 * rstep) {   // four independent 16-byte loads in flight per thread
    uint4 raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = pb + u * rstep;
      raw[u] = p < p1 ? xf[(int64_t)p * nvec + v] : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (pb + u * rstep >= p1) break;
      float e[8];
      cv_unpack8(raw[u], e);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (PASS == 1) {
          const float d = e[i] - mean[i];
          acc[i] = fmaf(d, d, acc[i]);
        } else {
          acc[i] += e[i];
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) part_sm[rl * C + 8 * v + i] = acc[i];
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < rstep; ++r)
      for (int c = 0; c < cpg; ++c) s += part_sm[r * C + g * cpg + c];
    partial[((int64_t)f * gridDim.x + blockIdx.x) * G + g] = s;
  }
}

template <int PASS>
__global__ void gn_finalize_kernel(const float* __restrict__ partial, float* __restrict__ stats, int nchunk, int G, float count,
                                   float eps) {
  const int f = blockIdx.x;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < nchunk; ++k) s += partial[((int64_t)f * nchunk + k) * G + g];
    if (PASS == 0) stats[(f * G + g) * 2] = s / count;
    else stats[(f * G + g) * 2 + 1] = rsqrtf(s / count + eps);
  }
}

// ---- out = swish(GroupNorm(x)): the activation in front of every convolution, applied once --------------------------------
__global__ void __launch_bounds__(256) gn_apply_kernel(const bf16* __restrict__ x, bf16* __restrict__ out,
                                                       const float* __restrict__ stats, const bf16* __restrict__ gamma,
                                                       const bf16* __restrict__ beta, int64_t P, int C, int G, int swish,
                                                       int64_t total_vec) {
  const int nvec = C >> 3;
  const int cpg = C / G;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    const int f = (int)(i / (P * nvec));
    float e[8], gm[8], bt[8];
    cv_unpack8(reinterpret_cast<const uint4*>(x)[i], e);
    cv_unpack8(reinterpret_cast<const uint4*>(gamma)[v], gm);
    cv_unpack8(reinterpret_cast<const uint4*>(beta)[v], bt);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float2 ms = reinterpret_cast<const float2*>(stats)[f * G + (8 * v + k) / cpg];
      float y = fmaf((e[k] - ms.x) * ms.y, gm[k], bt[k]);
      if (swish) y = y / (1.0f + __expf(-y));
      e[k] = y;
    }
    reinterpret_cast<uint4*>(out)[i] = cv_pack8(e);
  }
}

// ---- im2col of a 3x3 / stride 1 / zero-padded convolution over channels-last frames -----------------------------------
// out[(f, y, x), (ky, kx, c)] = act(x[f, y+ky-1, x+kx-1, c]); act = identity, or GroupNorm (+ swish) from (mean, rstd).
// The padding is applied AFTER the activation, as in the reference (norm -> swish -> conv with padding=1).
__global__ void __launch_bounds__(256) im2col3x3_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int H, int W, int C,
                                                        int64_t total_vec, const float* __restrict__ gn_stats,
                                                        const bf16* __restrict__ gamma, const bf16* __restrict__ beta, int G,
                                                        int swish) {
  const int nvec = C >> 3;
  const int cpg = C / (G > 0 ? G : 1);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    const int64_t j = i / nvec;
    const int tap = (int)(j % 9);
    const int64_t pos = j / 9;
    const int xx = (int)(pos % W);
    const int64_t q = pos / W;
    const int yy = (int)(q % H);
    const int f = (int)(q / H);
    const int sy = yy + tap / 3 - 1, sx = xx + tap % 3 - 1;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (sy >= 0 && sy < H && sx >= 0 && sx < W) {
      const uint4 raw = reinterpret_cast<const uint4*>(x)[(((int64_t)f * H + sy) * W + sx) * nvec + v];
      if (gn_stats != nullptr) {
        float e[8], gm[8], bt[8];
        cv_unpack8(raw, e);
        cv_unpack8(reinterpret_cast<const uint4*>(gamma)[v], gm);
        cv_unpack8(reinterpret_cast<const uint4*>(beta)[v], bt);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float2 ms = reinterpret_cast<const float2*>(gn_stats)[f * G + (8 * v + k) / cpg];
          float y = fmaf((e[k] - ms.x) * ms.y, gm[k], bt[k]);
          if (swish) y = y / (1.0f + __expf(-y));
          e[k] = y;
        }
        o = cv_pack8(e);
      } else {
        o = raw;
      }
    }
    reinterpret_cast<uint4*>(out)[i] = o;
  }
}

// ---- PixelShuffle(2) on channels-last frames: in [F, H, W, 4*Co] -> out [F, 2H, 2W, Co] --------------------------------
// torch.nn.PixelShuffle: out[c, 2h+dy, 2w+dx] = in[4c + 2dy + dx, h, w]      (vq_gan_blocks.py:47-48, 63-64)
__global__ void __launch_bounds__(256) pixel_shuffle2_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int H, int W,
                                                             int Co, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Co);
    const int64_t pos = i / Co;
    const int ox = (int)(pos % (2 * W));
    const int64_t q = pos / (2 * W);
    const int oy = (int)(q % (2 * H));
    const int f = (int)(q / (2 * H));
    out[i] = x[(((int64_t)f * H + (oy >> 1)) * W + (ox >> 1)) * (4 * Co) + 4 * c + 2 * (oy & 1) + (ox & 1)];
  }
}

// ---- direct 3x3 convolution with 16 output channels, channels-last in, NCHW out ---------------------------------------
// SemanticCond.conv_out (condition.py:49-56, 132-136): Cin = 64 -> 16, the tensor the control net adds to its latent.
__global__ void __launch_bounds__(256) conv3x3_to_nchw16_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                                const bf16* __restrict__ bias, bf16* __restrict__ out, int F,
                                                                int H, int W, int Cin) {
  extern __shared__ float wsm[];   // [9][Cin][16]
  for (int i = threadIdx.x; i < 9 * Cin * 16; i += blockDim.x) {
    const int co = i & 15, ci = (i >> 4) % Cin, tap = (i >> 4) / Cin;
    wsm[i] = __bfloat162float(w[((int64_t)co * Cin + ci) * 9 + tap]);   // torch layout [16, Cin, 3, 3]
  }
  __syncthreads();
  const int64_t total = (int64_t)F * H * W;
  const int nvec = Cin >> 3;
  for (int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pos < total; pos += (int64_t)gridDim.x * blockDim.x) {
    const int xx = (int)(pos % W);
    const int64_t q = pos / W;
    const int yy = (int)(q % H);
    const int f = (int)(q / H);
    float acc[16];
#pragma unroll
    for (int co = 0; co < 16; ++co) acc[co] = bias ? __bfloat162float(bias[co]) : 0.f;
    for (int tap = 0; tap < 9; ++tap) {
      const int sy = yy + tap / 3 - 1, sx = xx + tap % 3 - 1;
      if (sy < 0 || sy >= H || sx < 0 || sx >= W) continue;
      const uint4* src = reinterpret_cast<const uint4*>(x) + (((int64_t)f * H + sy) * W + sx) * nvec;
      const float* wt = wsm + (size_t)tap * Cin * 16;
      for (int v = 0; v < nvec; ++v) {
        float e[8];
        cv_unpack8(src[v], e);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4* wr = reinterpret_cast<const float4*>(wt + (8 * v + k) * 16);
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const float4 ww = wr[c4];
            acc[4 * c4 + 0] = fmaf(e[k], ww.x, acc[4 * c4 + 0]);
            acc[4 * c4 + 1] = fmaf(e[k], ww.y, acc[4 * c4 + 1]);
            acc[4 * c4 + 2] = fmaf(e[k], ww.z, acc[4 * c4 + 2]);
            acc[4 * c4 + 3] = fmaf(e[k], ww.w, acc[4 * c4 + 3]);
          }
        }
      }
    }
    const int64_t hw = (int64_t)H * W;
    bf16* o = out + (int64_t)f * 16 * hw + (int64_t)yy * W + xx;
#pragma unroll
    for (int co = 0; co < 16; ++co) o[co * hw] = __float2bfloat16_rn(acc[co]);
  }
}

static int grid_1d(int64_t work, int threads) {
  const int64_t want = (work + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count() * 16;
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace ld

extern "C" int ld_nchw_to_nhwc(const void* x, int x_is_f32, void* out, int frames, int C, int P, void* stream) {
  using namespace ld;
  int rc = check_device();
  if (rc != LD_OK) return rc;
  LD_CHECK_ARG(x && out && frames > 0 && C > 0 && P > 0, "ld_nchw_to_nhwc: bad arguments");
  LD_CHECK_ARG(frames <= 65535, "ld_nchw_to_nhwc: frames=%d too large", frames);
  const dim3 grid((P + 31) / 32, (C + 31) / 32, frames);
  cudaStream_t st = (cudaStream_t)stream;
  if (x_is_f32)
    nchw_to_nhwc_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (bf16*)out, C, P);
  else
    nchw_to_nhwc_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, (bf16*)out, C, P);
  LD_CHECK_CUDA(cudaGetLastError());
  return LD_OK;
}

extern "C" int ld_groupnorm_stats(const void* x, float* stats, float* scratch, int frames, int P, int C, int groups, float eps,
                                  void* stream) {
  using namespace ld;
  int rc = check_device();
  if (rc != LD_OK) return rc;
  LD_CHECK_ARG(x && stats && scratch && frames > 0 && P > 0, "ld_groupnorm_stats: bad arguments");
  LD_CHECK_ARG(C % 8 == 0 && groups > 0 && C % groups == 0 && C / 8 <= 256 && 256 % (C / 8) == 0,
               "ld_groupnorm_stats: C=%d must be a multiple of 8 and of groups=%d, with C/8 dividing 256", C, groups);
  LD_CHECK_ARG(frames <= 65535, "ld_groupnorm_stats: frames=%d too large", frames);
  cudaStream_t st = (cudaStream_t)stream;
  const int nchunk = (P + kGnRows - 1) / kGnRows;
  const dim3 grid(nchunk, frames);
  const size_t smem = (size_t)256 * 8 * sizeof(float);
  const float count = (float)P * (float)(C / groups);
  gn_partial_kernel<0><<<grid, 256, smem, st>>>((const bf16*)x, stats, scratch, P, C, groups);
  gn_finalize_kernel<0><<<frames, 64, 0, st>>>(scratch, stats, nchunk, groups, count, eps);
  gn_partial_kernel<1><<<grid, 256, smem, st>>>((const bf16*)x, stats, scratch, P, C, groups);
  gn_finalize_kernel<1><<<frames, 64, 0, st>>>(scratch, stats, nchunk, groups, count, eps);
  LD_CHECK_CUDA(cudaGetLastError());
  return LD_OK;
}

extern "C" int ld_groupnorm_apply(const void* x, void* out, const float* stats, const void* gamma, const void* beta, int frames,
                                  int P, int C, int groups, int swish, void* stream) {
  using namespace ld;
  int rc = check_device();
  if (rc != LD_OK) return rc;
  LD_CHECK_ARG(x && out && stats && gamma && beta && frames > 0 && P > 0, "ld_groupnorm_apply: bad arguments");
  LD_CHECK_ARG(C % 8 == 0 && groups > 0 && C % groups == 0, "ld_groupnorm_apply: C=%d must be a multiple of 8 and of groups", C);
  const int64_t total_vec = (int64_t)frames * P * (C / 8);
  gn_apply_kernel<<<grid_1d(total_vec, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)out, stats, (const bf16*)gamma,
                                                                             (const bf16*)beta, P, C, groups, swish, total_vec);
  LD_CHECK_CUDA(cudaGetLastError());
  return LD_OK;
}

extern "C" int ld_im2col3x3(const void* x, void* out, int frames, int H, int W, int C, const float* gn_stats, const void* gamma,
                            const void* beta, int groups, int swish, void* stream) {
  using namespace ld;
  int rc = check_device();
  if (rc != LD_OK) return rc;
  LD_CHECK_ARG(x && out && frames > 0 && H > 0 && W > 0, "ld_im2col3x3: bad arguments");
  LD_CHECK_ARG(C % 8 == 0 && C > 0, "ld_im2col3x3: C=%d must be a multiple of 8", C);
  if (gn_stats != nullptr)
    LD_CHECK_ARG(gamma && beta && groups > 0 && C % groups == 0, "ld_im2col3x3: GroupNorm needs stats, gamma, beta, groups");
  const int64_t total_vec = (int64_t)frames * H * W * 9 * (C / 8);
  im2col3x3_kernel<<<grid_1d(total_vec, 256), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)x, (bf16*)out, H, W, C, total_vec, gn_stats, (const bf16*)gamma, (const bf16*)beta, gn_stats ? groups : 0,
      swish);
  LD_CHECK_CUDA(cudaGetLastError());
  return LD_OK;
}

extern "C" int ld_pixel_shuffle2(const void* x, void* out, int frames, int H, int W, int C_out, void* stream) {
  using namespace ld;
  int rc = check_device();
  if (rc != LD_OK) return rc;
  LD_CHECK_ARG(x && out && frames > 0 && H > 0 && W > 0 && C_out > 0, "ld_pixel_shuffle2: bad arguments");
  const int64_t total = (int64_t)frames * 4 * H * W * C_out;
  pixel_shuffle2_kernel<<<grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)out, H, W, C_out, total);
  LD_CHECK_CUDA(cudaGetLastError());
  return LD_OK;
}

extern "C" int ld_conv3x3_to_nchw16(const void* x, const void* w, const void* bias, void* out, int frames, int H, int W, int Cin,
                                    void* stream) {
  using namespace ld;
  int rc = check_device();
  if (rc != LD_OK) return rc;
  LD_CHECK_ARG(x && w && out && frames > 0 && H > 0 && W > 0, "ld_conv3x3_to_nchw16: bad arguments");
  LD_CHECK_ARG(Cin % 8 == 0 && Cin > 0 && Cin <= 256, "ld_conv3x3_to_nchw16: Cin=%d must be a multiple of 8 and <= 256", Cin);
  const size_t smem = (size_t)9 * Cin * 16 * sizeof(float);
  static bool attr_set[64] = {};
  int dev = 0;
  LD_CHECK_CUDA(cudaGetDevice(&dev));
  if (smem > 48 * 1024 && dev >= 0 && dev < 64 && !attr_set[dev]) {
    LD_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_to_nchw16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set[dev] = true;
  }
  const int64_t total = (int64_t)frames * H * W;
  conv3x3_to_nchw16_kernel<<<grid_1d(total, 256), 256, smem, (cudaStream_t)stream>>>((const bf16*)x, (const bf16*)w,
                                                                                     (const bf16*)bias, (bf16*)out, frames, H, W,
                                                                                     Cin);
  LD_CHECK_CUDA(cudaGetLastError());
  return LD_OK;
}
