// Persistent warp-specialised bf16 GEMM for sm_100a:  out = epilogue(A[M,K] @ W[N,K]^T).
//
//   warp 0      : TMA producer  (A 128x64 and W BNx64 boxes, SWIZZLE_128B, STAGES-deep mbarrier ring)
//   warp 1      : tcgen05.mma issuer (one elected thread; UMMA 128 x BN x 16, fp32 accumulators in TMEM,
//                 two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 2..9  : epilogue (tcgen05.ld 32x32b -> registers -> fused epilogue -> global); two warps per TMEM
//                 lane quadrant split the tile's columns; global operands are fetched before the TMEM wait
//
// Both operands are K-major (activations row-major, nn.Linear weights [out,in]) so no transposes are needed.
// The fused epilogues implement the reference's per-layer elementwise work (bias, GELU-tanh, adaLN-gated
// residual, control add, QKV split + per-head QK-LayerNorm, patch-embed position add, unpatchify); see
// include/landiff_b200.h for the reference lines each one replaces.
#include <cstdlib>
#include "host_util.h"
#include "ptx.cuh"

namespace ld {

using bf16 = __nv_bfloat16;

template <int BN>
struct GemmCfg {
  static constexpr int BM = 128;
  static constexpr int BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 192) ? 5 : (BN >= 128 ? 6 : 8);
  static constexpr uint32_t TMEM_COLS = (BN <= 64) ? 128 : (BN <= 128 ? 256 : 512);
  static constexpr int ACC_STRIDE = TMEM_COLS / 2;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

constexpr int kGemmThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quadrant)

struct RowCtx {
  int row;        // A row
  bool valid;
  int b;          // sample
  int t;          // token within the local shard
  bool is_text;
  int64_t out_row;
};

__device__ __forceinline__ void load8_bf16(const bf16* p, float* f) {
  uint4 v = *reinterpret_cast<const uint4*>(p);
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ void store8_bf16(bf16* p, const float* f) {
  uint4 v;
  v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
  v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = v;
}

// residual stream may be kept in fp32 (main net) or bf16 (control net): element-index based typed access
__device__ __forceinline__ void load8_stream(const void* base, bool is_f32, int64_t idx, float* f) {
  if (is_f32) {
    const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
    const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    load8_bf16(reinterpret_cast<const bf16*>(base) + idx, f);
  }
}
__device__ __forceinline__ void store8_stream(void* base, bool is_f32, int64_t idx, const float* f) {
  if (is_f32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx + 4) = make_float4(f[4], f[5], f[6], f[7]);
  } else {
    store8_bf16(reinterpret_cast<bf16*>(base) + idx, f);
  }
}

// ---- epilogues on a 32-column chunk held by one thread (one output row) -------------------------------------
// Operands are fetched from global BEFORE the TMEM load is waited on, so their latency overlaps it.
template <int EPI>
struct EpiOperands {
  uint4 bias[4];    // 32 bf16
  uint4 resid[8];   // 32 bf16 (first 4) or 32 fp32 (all 8)
  uint4 add2[4];    // 32 bf16
  float4 gate[8];   // 32 fp32
  uint4 pos[4];     // 32 bf16
};

template <int EPI>
__device__ __forceinline__ void epi_fetch(const ld_gemm_args& p, const RowCtx& rc, int col, EpiOperands<EPI>& o) {
  if constexpr (EPI != LD_EPI_NONE) {
    if (p.bias != nullptr) {
      const uint4* b = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.bias) + col);
#pragma unroll
      for (int g = 0; g < 4; ++g) o.bias[g] = b[g];
    } else {
#pragma unroll
      for (int g = 0; g < 4; ++g) o.bias[g] = make_uint4(0, 0, 0, 0);
    }
  }
  if (!rc.valid) return;
  if constexpr (EPI == LD_EPI_GATED_RESID) {
    const int64_t eidx = rc.out_row * p.ld_out + col;
    if (p.resid_f32) {
      const uint4* r = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.resid) + eidx);
#pragma unroll
      for (int g = 0; g < 8; ++g) o.resid[g] = r[g];
    } else {
      const uint4* r = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.resid) + eidx);
#pragma unroll
      for (int g = 0; g < 4; ++g) o.resid[g] = r[g];
    }
    if (p.add2 != nullptr) {
      const uint4* a2 = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.add2) + eidx);
#pragma unroll
      for (int g = 0; g < 4; ++g) o.add2[g] = a2[g];
    }
    const float4* gt = reinterpret_cast<const float4*>((rc.is_text ? p.gate_txt : p.gate_img) +
                                                       (int64_t)rc.b * p.mod_batch_stride + col);
#pragma unroll
    for (int g = 0; g < 8; ++g) o.gate[g] = gt[g];
  }
  if constexpr (EPI == LD_EPI_BIAS_ADD) {
    const uint4* a2 = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.add2) + rc.out_row * p.ld_out + col);
#pragma unroll
    for (int g = 0; g < 4; ++g) o.add2[g] = a2[g];
  }
  if constexpr (EPI == LD_EPI_BIAS_POS) {
    const uint4* ps = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.pos) +
                                                     (int64_t)(p.tok_offset + rc.t) * p.N + col);
#pragma unroll
    for (int g = 0; g < 4; ++g) o.pos[g] = ps[g];
  }
}

__device__ __forceinline__ void unpack8_u4(const uint4& v, float* f) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}

template <int EPI>
__device__ __forceinline__ void epilogue32(const ld_gemm_args& p, const RowCtx& rc, int col, const uint32_t* r,
                                           const EpiOperands<EPI>& o) {
  if (!rc.valid) return;
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = __uint_as_float(r[i]);
  if constexpr (EPI != LD_EPI_NONE) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float bv[8];
      unpack8_u4(o.bias[g], bv);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[g * 8 + i] += bv[i];
    }
  }

  if constexpr (EPI == LD_EPI_NONE || EPI == LD_EPI_BIAS || EPI == LD_EPI_BIAS_GELU || EPI == LD_EPI_BIAS_ADD) {
    if constexpr (EPI == LD_EPI_BIAS_GELU) {
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = gelu_tanh(acc[i]);
    }
    if constexpr (EPI == LD_EPI_BIAS_ADD) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float av[8];
        unpack8_u4(o.add2[g], av);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[g * 8 + i] += av[i];
      }
    }
    bf16* op = reinterpret_cast<bf16*>(p.out) + rc.out_row * p.ld_out + col;
#pragma unroll
    for (int g = 0; g < 4; ++g) store8_bf16(op + g * 8, acc + g * 8);
  } else if constexpr (EPI == LD_EPI_GATED_RESID) {
    const int64_t eidx = rc.out_row * p.ld_out + col;
    const bool rf32 = p.resid_f32 != 0, of32 = p.out_f32 != 0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float rv[8];
      if (rf32) {
        const uint4 a = o.resid[2 * g], b = o.resid[2 * g + 1];
        rv[0] = __uint_as_float(a.x); rv[1] = __uint_as_float(a.y); rv[2] = __uint_as_float(a.z); rv[3] = __uint_as_float(a.w);
        rv[4] = __uint_as_float(b.x); rv[5] = __uint_as_float(b.y); rv[6] = __uint_as_float(b.z); rv[7] = __uint_as_float(b.w);
      } else {
        unpack8_u4(o.resid[g], rv);
      }
      const float4 g0 = o.gate[2 * g], g1 = o.gate[2 * g + 1];
      const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      float ov[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) ov[i] = fmaf(gv[i], acc[g * 8 + i], rv[i]);
      if (p.add2 != nullptr) {
        float av[8];
        unpack8_u4(o.add2[g], av);
#pragma unroll
        for (int i = 0; i < 8; ++i) ov[i] += av[i];
      }
      store8_stream(p.out, of32, eidx + g * 8, ov);
    }
  } else if constexpr (EPI == LD_EPI_BIAS_POS) {
    const int64_t eidx = rc.out_row * p.ld_out + col;
    const bool of32 = p.out_f32 != 0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float pv[8];
      unpack8_u4(o.pos[g], pv);
      float ov[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) ov[i] = acc[g * 8 + i] + pv[i];
      store8_stream(p.out, of32, eidx + g * 8, ov);
    }
  } else if constexpr (EPI == LD_EPI_UNPATCHIFY) {
    // col = c*4 + pp*2 + qq  ->  out[b, t, c, 2h+pp, 2w+qq]      (dit_video_concat.py:392-410)
    const int g = p.tok_offset + rc.t - p.text_len;
    const int hw = p.Hp * p.Wp;
    const int tt = g / hw, rem = g - tt * hw;
    const int h = rem / p.Wp, w = rem - h * p.Wp;
    const int H = 2 * p.Hp, W = 2 * p.Wp;
    bf16* op = reinterpret_cast<bf16*>(p.out);
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const int cc = col + i;
      const int c = cc >> 2, pp = (cc >> 1) & 1;
      const int64_t idx = (((int64_t)(rc.b * p.T + tt) * p.C + c) * H + (2 * h + pp)) * W + 2 * w;
      *reinterpret_cast<uint32_t*>(op + idx) = pack_bf16x2(acc[i], acc[i + 1]);
    }
  }
}

// QKV epilogue on one 64-column head slice: split, per-head LayerNorm for Q/K, head-major store.
__device__ __forceinline__ void epilogue_qkv64(const ld_gemm_args& p, const RowCtx& rc, int col, const uint32_t* r0,
                                               const uint32_t* r1) {
  float acc[64];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    acc[i] = __uint_as_float(r0[i]);
    acc[32 + i] = __uint_as_float(r1[i]);
  }
  const bf16* bias = reinterpret_cast<const bf16*>(p.bias) + col;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    float bv[8];
    load8_bf16(bias + g * 8, bv);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[g * 8 + i] += bv[i];
  }
  if (!rc.valid) return;
  const int hd_all = p.heads * 64;
  const int which = col / hd_all;  // 0 q, 1 k, 2 v
  const int head = (col - which * hd_all) >> 6;
  if (which < 2) {
    // the reference rounds the QKV projection to bf16 before the LayerNorm; keep that rounding point
    float mean = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      acc[i] = __bfloat162float(__float2bfloat16_rn(acc[i]));
      mean += acc[i];
    }
    mean *= (1.0f / 64.0f);
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      const float d = acc[i] - mean;
      var = fmaf(d, d, var);
    }
    const float rstd = rsqrtf(var * (1.0f / 64.0f) + p.ln_eps);
    const bf16* w = reinterpret_cast<const bf16*>(which == 0 ? p.q_ln_w : p.k_ln_w);
    const bf16* b = reinterpret_cast<const bf16*>(which == 0 ? p.q_ln_b : p.k_ln_b);
    const float post = (which == 0) ? p.q_scale : 1.0f;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float wv[8], bv[8];
      load8_bf16(w + g * 8, wv);
      load8_bf16(b + g * 8, bv);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float y = fmaf((acc[g * 8 + i] - mean) * rstd, wv[i], bv[i]);
        if (which == 0) y = __bfloat162float(__float2bfloat16_rn(y)) * post;  // bf16 LN output, then scale
        acc[g * 8 + i] = y;
      }
    }
  }
  bf16* base = reinterpret_cast<bf16*>(which == 0 ? p.q : (which == 1 ? p.k : p.v));
  bf16* o = base + (((int64_t)rc.b * p.heads + head) * p.qkv_rows + p.qkv_row_offset + rc.t) * 64;
#pragma unroll
  for (int g = 0; g < 8; ++g) store8_bf16(o + g * 8, acc + g * 8);
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const ld_gemm_args p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m = (p.M + Cfg::BM - 1) / Cfg::BM;
  const int num_n = p.N / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = p.K / Cfg::BK;

  // Producer and MMA warps run warp-uniform code and issue through an elect.sync leader, which ptxas compiles to
  // back-to-back UTMALDG / UTCHMMA (a `lane == 0` branch costs a divergence loop of ~12 instructions per MMA).
  if (warp == 0) {
    const bool leader = elect_one();
    int s = 0;
    uint32_t ph = 0;
    const bool conv = p.conv_W > 0;   // implicit 3x3 convolution: A tiles gathered by TMA in im2col mode
    const int cblocks = conv ? p.conv_C / Cfg::BK : 1;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / num_n) * Cfg::BM;
      const int n0 = (tile % num_n) * BN;
      // first output position of the tile -> base pixel in bounding-box coordinates (lower corner = -padding)
      const int cx = conv ? m0 % p.conv_W - 1 : 0;
      const int cy = conv ? (m0 / p.conv_W) % p.conv_H - 1 : 0;
      const int cf = conv ? m0 / (p.conv_W * p.conv_H) : 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (leader) {
          mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
          if (conv) {
            const int tap = kb / cblocks;
            tma_load_im2col_4d(sA + s * Cfg::A_BYTES, &tmap_a, &full_bar[s], (kb - tap * cblocks) * Cfg::BK, cx, cy, cf,
                               (uint16_t)(tap % 3), (uint16_t)(tap / 3));
          } else {
            tma_load_2d(sA + s * Cfg::A_BYTES, &tmap_a, &full_bar[s], kb * Cfg::BK, m0);
          }
          tma_load_2d(sB + s * Cfg::B_BYTES, &tmap_b, &full_bar[s], kb * Cfg::BK, n0);
        }
        if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_bf16(Cfg::BM, BN);
    const uint64_t adesc0 = make_sdesc_sw128(smem_u32(sA));
    const uint64_t bdesc0 = make_sdesc_sw128(smem_u32(sB));
    int s = 0;
    uint32_t ph = 0;
    int as = 0;
    uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[as], aph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * Cfg::ACC_STRIDE;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint64_t adesc = adesc0 + uint32_t(s * (Cfg::A_BYTES >> 4));
        const uint64_t bdesc = bdesc0 + uint32_t(s * (Cfg::B_BYTES >> 4));
        if (leader) {
#pragma unroll
          for (int k = 0; k < Cfg::BK / 16; ++k) {
            // +32 bytes (encoded >>4 = 2) per 16-element K step inside the 128-byte swizzle span
            umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
      }
      if (leader) umma_commit(&tfull_bar[as]);
      if (++as == 2) { as = 0; aph ^= 1; }
    }
  } else {
    const int quad = warp & 3;        // TMEM lane quadrant this warp may access (warp id % 4)
    const int half = (warp - 2) >> 2;  // two warps share a quadrant and split the tile's columns
    const int row_in_tile = quad * 32 + lane;
    // column split: QKV works on whole 64-wide heads, everything else on 32-column chunks
    constexpr int kSplit = (EPI == LD_EPI_QKV) ? ((BN >= 128) ? (BN / 128) * 64 : 64) : (BN / 64) * 32;
    const int c_begin = half == 0 ? 0 : kSplit;
    const int c_end = half == 0 ? kSplit : BN;
    int as = 0;
    uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / num_n) * Cfg::BM;
      const int n0 = (tile % num_n) * BN;
      RowCtx rc;
      rc.row = m0 + row_in_tile;
      rc.valid = rc.row < p.M;
      const int rr = rc.valid ? rc.row : 0;
      rc.b = rr / p.rows_per_batch;
      rc.t = rr - rc.b * p.rows_per_batch;
      rc.is_text = (p.tok_offset + rc.t) < p.text_len;
      rc.out_row = (int64_t)rc.b * p.out_rows_per_batch + p.out_row_offset + rc.t;

      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + as * Cfg::ACC_STRIDE;
      if constexpr (EPI == LD_EPI_QKV) {
#pragma unroll 1
        for (int c = c_begin; c < c_end; c += 64) {
          uint32_t r0[32], r1[32];
          LD_TMEM_LD32(taddr + c, r0);
          LD_TMEM_LD32(taddr + c + 32, r1);
          tmem_ld_wait();
          epilogue_qkv64(p, rc, n0 + c, r0, r1);
        }
      } else {
#pragma unroll 1
        for (int c = c_begin; c < c_end; c += 32) {
          EpiOperands<EPI> ops;
          epi_fetch<EPI>(p, rc, n0 + c, ops);
          uint32_t r[32];
          LD_TMEM_LD32(taddr + c, r);
          tmem_ld_wait();
          epilogue32<EPI>(p, rc, n0 + c, r, ops);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aph ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------
// CTA-pair GEMM (round 2): 256 x 192 output tiles computed by two CTAs of a cluster with tcgen05.mma.cta_group::2.
//
// With 128 x 192 tiles one k-block of 64 moves (128 + 192) x 128 B = 40 KB from L2 into shared memory for 3.1 MFLOP —
// 78 FLOP per byte, i.e. ~17 TB/s of L2 -> SM traffic at 1330 TFLOP/s, which is what the fabric delivers: the 1-CTA kernel
// sits at 0.76-0.89 of the measured cuBLAS peak with its tensor pipe 70-90 % busy (profiles/r1_ncu_gemm_v2.csv).  A CTA
// pair shares the W tile: each CTA loads its 128 rows of A and HALF of the W tile (96 rows), the UMMA reads both halves
// from the two shared memories -> 28 KB per CTA and k-block, 112 FLOP per byte, and room for 7 stages instead of 5.
//   per CTA: warp 0 TMA producer (both CTAs), warp 1 MMA issuer (leader CTA only), warps 2..9 epilogue (own 128 rows)
//   full[s]   leader's barrier: ONE arrival, the leader's arrive.expect_tx of the bytes of BOTH CTAs.  The peer's TMA
//             completes on it too but does not arrive: a remote mbarrier.arrive per k-block costs a cluster-scope release
//             fence (MEMBAR + ERRBAR, ~5 % of all stall samples and a producer slower than the MMAs — measured, 605
//             TFLOP/s).  Bytes of the peer that land before the leader's expect_tx only drive the pending count negative;
//             they cannot complete a phase (the leader's arrival is still outstanding) and cannot hit the previous phase
//             (the peer refills a stage only after the multicast commit of the MMAs that consumed it).
//   empty[s], tfull[a]   one per CTA, signalled in both by multicast tcgen05.commit
//   tempty[a] leader's barrier: 8 epilogue warps of each CTA arrive (the peer's through mapa)
struct Gemm2Cfg {
  static constexpr int BN = 192;
  static constexpr int BM = 128;           // rows per CTA; 256 per pair
  static constexpr int BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = 7;
  static constexpr uint32_t TMEM_COLS = 512;
  static constexpr int ACC_STRIDE = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
             const ld_gemm_args p) {
  using Cfg = Gemm2Cfg;
  constexpr int BN = Cfg::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();      // 0 = leader
  const bool is_leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 16);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();      // barrier inits and TMEM allocations of both CTAs are visible before any remote access
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m2 = (p.M + 2 * Cfg::BM - 1) / (2 * Cfg::BM);   // 256-row pair tiles
  const int num_n = p.N / BN;
  const int num_tiles = num_m2 * num_n;
  const int num_kb = p.K / Cfg::BK;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    const bool leader_lane = elect_one();
    int s = 0;
    uint32_t ph = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int m0 = (tile / num_n) * (2 * Cfg::BM) + (int)rank * Cfg::BM;
      const int n0 = (tile % num_n) * BN + (int)rank * (BN / 2);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (leader_lane) {
          const uint32_t full0 = mapa_shared(smem_u32(&full_bar[s]), 0);   // the leader's full barrier
          if (is_leader) mbar_expect_tx(&full_bar[s], 2 * Cfg::STAGE_BYTES);
          tma_load_2d_2cta(sA + s * Cfg::A_BYTES, &tmap_a, full0, kb * Cfg::BK, m0);
          tma_load_2d_2cta(sB + s * Cfg::B_BYTES, &tmap_b, full0, kb * Cfg::BK, n0);
        }
        if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (is_leader) {
      const bool leader_lane = elect_one();
      constexpr uint32_t idesc = make_idesc_bf16(2 * Cfg::BM, BN);
      const uint64_t adesc0 = make_sdesc_sw128(smem_u32(sA));
      const uint64_t bdesc0 = make_sdesc_sw128(smem_u32(sB));
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * Cfg::ACC_STRIDE;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint64_t adesc = adesc0 + uint32_t(s * (Cfg::A_BYTES >> 4));
          const uint64_t bdesc = bdesc0 + uint32_t(s * (Cfg::B_BYTES >> 4));
          if (leader_lane) {
#pragma unroll
            for (int k = 0; k < Cfg::BK / 16; ++k)
              umma_ss_2cta(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_2cta(&empty_bar[s]);
          }
          if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
        }
        if (leader_lane) umma_commit_2cta(&tfull_bar[as]);
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (each CTA: its own 128 rows)
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row_in_tile = quad * 32 + lane;
    constexpr int kSplit = (EPI == LD_EPI_QKV) ? 64 : 96;
    const int c_begin = half == 0 ? 0 : kSplit;
    const int c_end = half == 0 ? kSplit : BN;
    int as = 0;
    uint32_t aph = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int m0 = (tile / num_n) * (2 * Cfg::BM) + (int)rank * Cfg::BM;
      const int n0 = (tile % num_n) * BN;
      RowCtx rc;
      rc.row = m0 + row_in_tile;
      rc.valid = rc.row < p.M;
      const int rr = rc.valid ? rc.row : 0;
      rc.b = rr / p.rows_per_batch;
      rc.t = rr - rc.b * p.rows_per_batch;
      rc.is_text = (p.tok_offset + rc.t) < p.text_len;
      rc.out_row = (int64_t)rc.b * p.out_rows_per_batch + p.out_row_offset + rc.t;

      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(quad * 32) << 16) + as * Cfg::ACC_STRIDE;
      if constexpr (EPI == LD_EPI_QKV) {
#pragma unroll 1
        for (int c = c_begin; c < c_end; c += 64) {
          uint32_t r0[32], r1[32];
          LD_TMEM_LD32(taddr + c, r0);
          LD_TMEM_LD32(taddr + c + 32, r1);
          tmem_ld_wait();
          epilogue_qkv64(p, rc, n0 + c, r0, r1);
        }
      } else {
#pragma unroll 1
        for (int c = c_begin; c < c_end; c += 32) {
          EpiOperands<EPI> ops;
          epi_fetch<EPI>(p, rc, n0 + c, ops);
          uint32_t r[32];
          LD_TMEM_LD32(taddr + c, r);
          tmem_ld_wait();
          epilogue32<EPI>(p, rc, n0 + c, r, ops);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[as]), 0));   // the leader's MMA warp waits
      if (++as == 2) { as = 0; aph ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();      // no CTA leaves (or frees TMEM) while its partner may still touch its shared memory / TMEM
  if (warp == 1) tmem_dealloc_2cta<Cfg::TMEM_COLS>(tmem_base);
}

template <int EPI>
static int launch_gemm2(const ld_gemm_args& a, cudaStream_t stream) {
  using Cfg = Gemm2Cfg;
  CUtensorMap ta, tb;
  {
    const uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.M};
    const uint64_t str[1] = {(uint64_t)a.K * 2};
    const uint32_t box[2] = {64, 128};
    int rc = make_tmap_bf16(&ta, a.A, 2, dims, str, box);
    if (rc != LD_OK) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.N};
    const uint64_t str[1] = {(uint64_t)a.K * 2};
    const uint32_t box[2] = {64, (uint32_t)(Cfg::BN / 2)};
    int rc = make_tmap_bf16(&tb, a.W, 2, dims, str, box);
    if (rc != LD_OK) return rc;
  }
  auto kern = gemm2_kernel<EPI>;
  static bool attr_set[64] = {false};  // per template instantiation and device
  int dev = 0;
  LD_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    LD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set[dev] = true;
  }
  const int num_tiles = ((a.M + 255) / 256) * (a.N / Cfg::BN);
  const int max_clusters = sm_count() / 2;
  const int clusters = num_tiles < max_clusters ? num_tiles : max_clusters;
  kern<<<2 * clusters, kGemmThreads, Cfg::SMEM_BYTES, stream>>>(ta, tb, a);
  LD_CHECK_CUDA(cudaGetLastError());
  return LD_OK;
}

template <int BN, int EPI>
static int launch_gemm(const ld_gemm_args& a, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap ta, tb;
  if (a.conv_W > 0) {
    int rc = make_tmap_im2col3x3_bf16(&ta, a.A, a.conv_C, a.conv_W, a.conv_H, a.conv_F, 128);
    if (rc != LD_OK) return rc;
  } else {
    const uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.M};
    const uint64_t str[1] = {(uint64_t)a.K * 2};
    const uint32_t box[2] = {64, 128};
    int rc = make_tmap_bf16(&ta, a.A, 2, dims, str, box);
    if (rc != LD_OK) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a.K, (uint64_t)a.N};
    const uint64_t str[1] = {(uint64_t)a.K * 2};
    const uint32_t box[2] = {64, (uint32_t)BN};
    int rc = make_tmap_bf16(&tb, a.W, 2, dims, str, box);
    if (rc != LD_OK) return rc;
  }
  auto kern = gemm_kernel<BN, EPI>;
  static bool attr_set[64] = {false};  // per template instantiation and device
  int dev = 0;
  LD_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    LD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set[dev] = true;
  }
  const int num_tiles = ((a.M + 127) / 128) * (a.N / BN);
  const int grid = num_tiles < sm_count() ? num_tiles : sm_count();
  kern<<<grid, kGemmThreads, Cfg::SMEM_BYTES, stream>>>(ta, tb, a);
  LD_CHECK_CUDA(cudaGetLastError());
  return LD_OK;
}

static bool use_pair_kernel() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LD_GEMM_1CTA");   // tuning / A-B switch: force the 1-CTA kernel
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

template <int EPI>
static int dispatch_bn(const ld_gemm_args& a, cudaStream_t stream) {
  if constexpr (EPI != LD_EPI_UNPATCHIFY && EPI != LD_EPI_BIAS_POS) {
    // the five per-layer GEMMs (QKV, out-proj, fc1, fc2, zero-linear): CTA pairs, 256 x 192 tiles
    if (a.N % 192 == 0 && a.M >= 256 && a.conv_W == 0 && use_pair_kernel()) return launch_gemm2<EPI>(a, stream);
  }
  if (a.N % 192 == 0) return launch_gemm<192, EPI>(a, stream);
  if (a.N % 128 == 0) return launch_gemm<128, EPI>(a, stream);
  return launch_gemm<64, EPI>(a, stream);
}

}  // namespace ld

extern "C" int ld_gemm_bf16(const ld_gemm_args* args, void* stream) {
  using namespace ld;
  int rc = check_device();
  if (rc != LD_OK) return rc;
  LD_CHECK_ARG(args != nullptr, "ld_gemm_bf16: null args");
  const ld_gemm_args& a = *args;
  LD_CHECK_ARG(a.M > 0 && a.N > 0 && a.K > 0, "ld_gemm_bf16: bad shape M=%d N=%d K=%d", a.M, a.N, a.K);
  LD_CHECK_ARG(a.K % 64 == 0, "ld_gemm_bf16: K=%d must be a multiple of 64", a.K);
  LD_CHECK_ARG(a.N % 64 == 0, "ld_gemm_bf16: N=%d must be a multiple of 64", a.N);
  LD_CHECK_ARG(a.A && a.W, "ld_gemm_bf16: null operand");
  LD_CHECK_ARG(a.rows_per_batch > 0 && a.M % a.rows_per_batch == 0,
               "ld_gemm_bf16: M=%d not a multiple of rows_per_batch=%d", a.M, a.rows_per_batch);
  if (a.conv_W > 0) {
    LD_CHECK_ARG(a.conv_F > 0 && a.conv_H > 0 && a.conv_C > 0 && a.conv_C % 64 == 0 && a.K == 9 * a.conv_C &&
                     (int64_t)a.M == (int64_t)a.conv_F * a.conv_H * a.conv_W,
                 "ld_gemm_bf16: implicit convolution needs conv_C %% 64 == 0, K == 9*conv_C and M == conv_F*conv_H*conv_W");
    LD_CHECK_ARG(a.epilogue == LD_EPI_NONE || a.epilogue == LD_EPI_BIAS || a.epilogue == LD_EPI_BIAS_GELU ||
                     a.epilogue == LD_EPI_BIAS_ADD, "ld_gemm_bf16: implicit convolution supports the plain epilogues only");
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a.epilogue != LD_EPI_QKV) {
    LD_CHECK_ARG(a.out != nullptr, "ld_gemm_bf16: null out");
    LD_CHECK_ARG(a.epilogue == LD_EPI_UNPATCHIFY || (a.ld_out >= a.N && a.ld_out % 8 == 0),
                 "ld_gemm_bf16: ld_out=%lld must be >= N and a multiple of 8", (long long)a.ld_out);
  }
  switch (a.epilogue) {
    case LD_EPI_NONE: return dispatch_bn<LD_EPI_NONE>(a, st);
    case LD_EPI_BIAS: return dispatch_bn<LD_EPI_BIAS>(a, st);
    case LD_EPI_BIAS_GELU: return dispatch_bn<LD_EPI_BIAS_GELU>(a, st);
    case LD_EPI_BIAS_ADD:
      LD_CHECK_ARG(a.add2 != nullptr, "ld_gemm_bf16: BIAS_ADD needs add2");
      return dispatch_bn<LD_EPI_BIAS_ADD>(a, st);
    case LD_EPI_GATED_RESID:
      LD_CHECK_ARG(a.resid && a.gate_img && a.gate_txt, "ld_gemm_bf16: GATED_RESID needs resid and gates");
      return dispatch_bn<LD_EPI_GATED_RESID>(a, st);
    case LD_EPI_QKV:
      LD_CHECK_ARG(a.q && a.k && a.v && a.bias && a.q_ln_w && a.q_ln_b && a.k_ln_w && a.k_ln_b,
                   "ld_gemm_bf16: QKV needs q/k/v, bias and LayerNorm parameters");
      LD_CHECK_ARG(a.heads > 0 && a.N == 3 * a.heads * 64, "ld_gemm_bf16: QKV needs N == 3*heads*64");
      LD_CHECK_ARG(a.qkv_rows >= a.qkv_row_offset + a.rows_per_batch, "ld_gemm_bf16: qkv_rows too small");
      return dispatch_bn<LD_EPI_QKV>(a, st);
    case LD_EPI_BIAS_POS:
      LD_CHECK_ARG(a.pos != nullptr, "ld_gemm_bf16: BIAS_POS needs pos");
      return dispatch_bn<LD_EPI_BIAS_POS>(a, st);
    case LD_EPI_UNPATCHIFY:
      LD_CHECK_ARG(a.N == 64 && a.C == 16 && a.T > 0 && a.Hp > 0 && a.Wp > 0,
                   "ld_gemm_bf16: UNPATCHIFY needs N=64, C=16 and the patch grid");
      return dispatch_bn<LD_EPI_UNPATCHIFY>(a, st);
    default:
      set_error("ld_gemm_bf16: unknown epilogue %d", a.epilogue);
      return LD_ERR_ARG;
  }
}
