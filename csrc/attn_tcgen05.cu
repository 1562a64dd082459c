// Full (non-causal) self-attention for head_dim 64 on sm_100a: flash-style online softmax with both
// contractions on tcgen05 tensor cores and S / P / O resident in TMEM.
//
// One CTA owns 256 query rows of one (sample, head) as two 128-row tiles that ping-pong on the tensor pipe:
//   warp 0        TMA producer: Q once, then K_j / V_j 128x64 boxes through KS-deep mbarrier rings
//   warp 1        tcgen05.mma issuer (one thread):  S_t = Q_t K_j^T  (128x128x64, SS),
//                                                   O_t += P_t V_j   (128x64x128, P from TMEM or smem; V N-major)
//   warps 4-7     softmax for tile 0 (one thread per query row: tcgen05.ld S row -> max / exp2 / sum ->
//   warps 8-11    softmax for tile 1   bf16 P -> tcgen05.st over S (or swizzled smem) ; lazy O rescale ; epilogue)
// The softmax of tile t overlaps the MMAs of tile 1-t.  K/V tail columns are masked to -inf; TMA zero-fills
// out-of-range rows.  Replaces F.scaled_dot_product_attention as reached from dit_video_concat.py:655-664.
#include "host_util.h"
#include "ptx.cuh"

namespace ld {

using bf16 = __nv_bfloat16;

constexpr int kAttnThreads = 384;
constexpr int kKS = 3;                       // K/V ring depth
constexpr int kTileBytes = 128 * 64 * 2;     // 16 KB: one 128x64 bf16 box
constexpr int kPBytes = 128 * 128 * 2;       // 32 KB: one P tile in smem (variant 1)
constexpr int kAttnSmem = 2 * kTileBytes + 2 * kKS * kTileBytes + 2 * kPBytes + 1024 + 256;
constexpr float kRescaleThreshold = 8.0f;    // log2 units

struct AttnParams {
  bf16* out;        // [B, nq, heads*64]
  float* lse;       // [BH, nq] or null
  float* out_f32;   // [BH, nq, 64] or null
  int heads, nq, nkv;
  float scale_log2;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for x <= ~8 on the FMA/ALU pipes (Cody-Waite range reduction + degree-3 minimax polynomial, max rel. error
// 8.8e-5 — far below the bf16 rounding of P): floor via round-down magic add, fraction in [0,1), exponent
// re-inserted by an integer add.  Offloads part of the softmax exponentials from the 16-op/clk MUFU unit.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -127.0f);
  float xr;
  asm("add.rm.ftz.f32 %0, %1, %2;" : "=f"(xr) : "f"(x), "f"(12582912.0f));  // 1.5 * 2^23: low mantissa bits = floor(x)
  const float f = x - (xr - 12582912.0f);
  float p = fmaf(f, 0.077119089663028717f, 0.227564394474029541f);
  p = fmaf(p, f, 0.695146143436431885f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xr) << 23));
}

// POLY_EVERY = n > 0: every n-th exponential of a row goes to ex2_poly, the others to MUFU.EX2; 0: all MUFU.
template <bool P_IN_TMEM, int POLY_EVERY>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
            const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                // 2 tiles
  uint8_t* sK = sQ + 2 * kTileBytes;                 // kKS tiles
  uint8_t* sV = sK + kKS * kTileBytes;               // kKS tiles
  uint8_t* sP = sV + kKS * kTileBytes;               // 2 x 32 KB (variant 1 only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kPBytes);
  uint64_t* q_full = bars;               // 1
  uint64_t* k_full = bars + 1;           // kKS
  uint64_t* k_empty = k_full + kKS;
  uint64_t* v_full = k_empty + kKS;
  uint64_t* v_empty = v_full + kKS;
  uint64_t* s_full = v_empty + kKS;      // 2
  uint64_t* p_full = s_full + 2;         // 2
  uint64_t* o_done = p_full + 2;         // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_blocks = (p.nq + 255) / 256;
  const int bh = blockIdx.x / q_blocks;
  const int q0 = (blockIdx.x % q_blocks) * 256;
  const int n_tiles = (p.nkv + 127) / 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKS; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 128);
      mbar_init(&o_done[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384); P_t aliases the first 64 columns of S_t

  if (warp < 4) {
    reg_dealloc<56>();
    if (warp == 0 && lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      mbar_expect_tx(q_full, 2 * kTileBytes);
      tma_load_3d(sQ, &tmap_q, q_full, 0, q0, bh);
      tma_load_3d(sQ + kTileBytes, &tmap_q, q_full, 0, q0 + 128, bh);
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_expect_tx(&k_full[s], kTileBytes);
        tma_load_3d(sK + s * kTileBytes, &tmap_k, &k_full[s], 0, j * 128, bh);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_expect_tx(&v_full[s], kTileBytes);
        tma_load_3d(sV + s * kTileBytes, &tmap_v, &v_full[s], 0, j * 128, bh);
        if (++s == kKS) { s = 0; ph ^= 1; }
      }
    } else if (warp == 1 && lane == 0) {
      // ---------------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, false, true);  // B (=V) is N-major
      auto issue_s = [&](int t, int stage) {
        const uint64_t adesc = make_sdesc_sw128(smem_u32(sQ + t * kTileBytes));
        const uint64_t bdesc = make_sdesc_sw128(smem_u32(sK + stage * kTileBytes));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem_base + t * 128, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0);
        umma_commit(&s_full[t]);
      };
      auto issue_pv = [&](int t, int stage, bool first) {
        const uint64_t bdesc = make_sdesc_sw128(smem_u32(sV + stage * kTileBytes));
        const uint32_t d = tmem_base + 256 + t * 64;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t acc = (first && k == 0) ? 0u : 1u;
          // 16 keys per step: V rows advance 16*128 B = 2048 B (encoded 128)
          if constexpr (P_IN_TMEM) {
            umma_ts(d, tmem_base + t * 128 + k * 8, bdesc + 128 * k, idesc_o, acc);
          } else {
            const uint64_t adesc = make_sdesc_sw128(smem_u32(sP + t * kPBytes + (k >> 2) * kTileBytes)) + 2 * (k & 3);
            umma_ss(d, adesc, bdesc + 128 * k, idesc_o, acc);
          }
        }
        umma_commit(&o_done[t]);
      };

      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      umma_commit(&k_empty[0]);
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < n_tiles; ++j) {
        int sn = s + 1;
        uint32_t phn = ph;
        if (sn == kKS) { sn = 0; phn ^= 1; }
        const bool has_next = (j + 1) < n_tiles;
        const uint32_t jpar = j & 1;
        // tile 0
        mbar_wait(&p_full[0], jpar);
        mbar_wait(&v_full[s], ph);
        tc_fence_after();
        issue_pv(0, s, j == 0);
        if (has_next) {
          mbar_wait(&k_full[sn], phn);
          tc_fence_after();
          issue_s(0, sn);
        }
        // tile 1
        mbar_wait(&p_full[1], jpar);
        tc_fence_after();
        issue_pv(1, s, j == 0);
        umma_commit(&v_empty[s]);
        if (has_next) {
          issue_s(1, sn);
          umma_commit(&k_empty[sn]);
        }
        s = sn;
        ph = phn;
      }
    }
  } else {
    reg_alloc<208>();
    // ------------------------------------------------------------------ softmax / correction / epilogue
    const int t = (warp - 4) >> 2;        // query tile 0/1
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const int q_row = q0 + t * 128 + row_in_tile;
    const uint32_t lane_base = uint32_t(quad * 32) << 16;
    const uint32_t ts = tmem_base + lane_base + t * 128;       // S_t (and P_t)
    const uint32_t to = tmem_base + lane_base + 256 + t * 64;  // O_t
    uint8_t* sPt = sP + t * kPBytes;
    const float sl2 = p.scale_log2;

    float m_used = -INFINITY;  // raw-score units
    float l = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      uint32_t s[128];
      LD_TMEM_LD32(ts + 0, (s + 0));
      LD_TMEM_LD32(ts + 32, (s + 32));
      LD_TMEM_LD32(ts + 64, (s + 64));
      LD_TMEM_LD32(ts + 96, (s + 96));
      tmem_ld_wait();
      const int valid = p.nkv - j * 128;
      if (valid < 128) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= valid) s[i] = 0xff800000u;  // -inf
      }
      float mx4[4] = {__uint_as_float(s[0]), __uint_as_float(s[1]), __uint_as_float(s[2]), __uint_as_float(s[3])};
#pragma unroll
      for (int i = 4; i < 128; i += 4) {  // four independent chains (the compiler fuses pairs into FMNMX3)
        mx4[0] = fmaxf(mx4[0], __uint_as_float(s[i]));
        mx4[1] = fmaxf(mx4[1], __uint_as_float(s[i + 1]));
        mx4[2] = fmaxf(mx4[2], __uint_as_float(s[i + 2]));
        mx4[3] = fmaxf(mx4[3], __uint_as_float(s[i + 3]));
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));

      bool need = (j > 0) && ((mx - m_used) * sl2 > kRescaleThreshold);
      if (j == 0) m_used = mx;
      if (__any_sync(0xffffffffu, need)) {
        // lazy correction: bring O_t and l to the new reference maximum (whole warp, tcgen05.ld/st are collective)
        const float m_new = fmaxf(m_used, mx);
        const float alpha = ex2((m_used - m_new) * sl2);
        mbar_wait(&o_done[t], (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 64; c += 32) {
          uint32_t o[32];
          LD_TMEM_LD32(to + c, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          LD_TMEM_ST32(to + c, o);
        }
        l *= alpha;
        m_used = m_new;
      }
      const float msc = m_used * sl2;
      float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
      uint32_t pk[64];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float pv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int idx = 4 * i + e;
          const float x = fmaf(__uint_as_float(s[idx]), sl2, -msc);
          if constexpr (POLY_EVERY > 0) {
            pv[e] = ((idx % POLY_EVERY) == POLY_EVERY - 1) ? ex2_poly(x) : ex2(x);
          } else {
            pv[e] = ex2(x);
          }
        }
        sum0 += pv[0];
        sum1 += pv[1];
        sum2 += pv[2];
        sum3 += pv[3];
        pk[2 * i] = pack_bf16x2(pv[0], pv[1]);
        pk[2 * i + 1] = pack_bf16x2(pv[2], pv[3]);
      }
      l += (sum0 + sum1) + (sum2 + sum3);
      if constexpr (P_IN_TMEM) {
        LD_TMEM_ST32(ts + 0, (pk + 0));
        LD_TMEM_ST32(ts + 32, (pk + 32));
        tmem_st_wait();
        tc_fence_before();
      } else {
        if (j > 0) mbar_wait(&o_done[t], (j - 1) & 1);  // P_t smem still being read by PV_t(j-1)
        tmem_st_wait();
        // row r of panel kp (64 keys) lives at kp*16KB + r*128 B, 16-byte chunk c stored at (c ^ (r & 7))
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int kp = c >> 3, cc = c & 7;
          uint4 v = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          *reinterpret_cast<uint4*>(sPt + kp * kTileBytes + row_in_tile * 128 + ((cc ^ (row_in_tile & 7)) << 4)) = v;
        }
        fence_proxy_async_smem();
        tc_fence_before();
      }
      mbar_arrive(&p_full[t]);
    }

    // epilogue: O / l
    mbar_wait(&o_done[t], (n_tiles - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const bool valid_row = q_row < p.nq;
    const int b = bh / p.heads, h = bh - b * p.heads;
    bf16* orow = p.out + ((int64_t)b * p.nq + q_row) * (p.heads * 64) + h * 64;
#pragma unroll
    for (int c = 0; c < 64; c += 32) {
      uint32_t o[32];
      LD_TMEM_LD32(to + c, o);
      tmem_ld_wait();
      if (valid_row) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(o[g * 8 + i]) * inv_l;
          uint4 v;
          v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
          v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
          *reinterpret_cast<uint4*>(orow + c + g * 8) = v;
          if (p.out_f32 != nullptr) {
            float* of = p.out_f32 + ((int64_t)bh * p.nq + q_row) * 64 + c + g * 8;
            *reinterpret_cast<float4*>(of) = make_float4(f[0], f[1], f[2], f[3]);
            *reinterpret_cast<float4*>(of + 4) = make_float4(f[4], f[5], f[6], f[7]);
          }
        }
      }
    }
    if (valid_row && p.lse != nullptr) p.lse[(int64_t)bh * p.nq + q_row] = m_used * sl2 + log2f(l);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// (o_acc, lse_acc) <- merge with (o_new, lse_new); log2-domain LSE
__global__ void __launch_bounds__(256) attn_merge_kernel(float* __restrict__ o_acc, float* __restrict__ lse_acc,
                                                         const float* __restrict__ o_new,
                                                         const float* __restrict__ lse_new, bf16* __restrict__ out_bf16,
                                                         int heads, int nq, int64_t rows) {
  // 16 threads per row (4 floats each)
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = gid >> 4;
  const int c = (int)(gid & 15) * 4;
  if (row >= rows) return;
  const float la = lse_acc[row], lb = lse_new[row];
  const float m = fmaxf(la, lb);
  const float wa = exp2f(la - m), wb = exp2f(lb - m);
  const float inv = 1.0f / (wa + wb);
  const float4 a = *reinterpret_cast<const float4*>(o_acc + row * 64 + c);
  const float4 b = *reinterpret_cast<const float4*>(o_new + row * 64 + c);
  float4 r;
  r.x = (a.x * wa + b.x * wb) * inv; r.y = (a.y * wa + b.y * wb) * inv;
  r.z = (a.z * wa + b.z * wb) * inv; r.w = (a.w * wa + b.w * wb) * inv;
  *reinterpret_cast<float4*>(o_acc + row * 64 + c) = r;
  if (out_bf16 != nullptr) {
    const int64_t bhi = row / nq, qi = row - bhi * nq;
    const int64_t bi = bhi / heads, hi = bhi - bi * heads;
    bf16* o = out_bf16 + (bi * nq + qi) * (heads * 64) + hi * 64 + c;
    uint2 v;
    v.x = pack_bf16x2(r.x, r.y);
    v.y = pack_bf16x2(r.z, r.w);
    *reinterpret_cast<uint2*>(o) = v;
  }
  __syncwarp();
  if ((gid & 15) == 0) lse_acc[row] = m + log2f(wa + wb);
}

template <bool P_IN_TMEM, int POLY_EVERY>
static int launch_attn(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& prm,
                       int grid, cudaStream_t st) {
  auto kern = attn_kernel<P_IN_TMEM, POLY_EVERY>;
  static bool attr_set = false;
  if (!attr_set) {
    LD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    attr_set = true;
  }
  kern<<<grid, kAttnThreads, kAttnSmem, st>>>(tq, tk, tv, prm);
  LD_CHECK_CUDA(cudaGetLastError());
  return LD_OK;
}

}  // namespace ld

using namespace ld;

extern "C" int ld_attention_bf16(const void* q, const void* k, const void* v, void* out, float* lse, float* out_f32,
                                 int batch, int heads, int nq, int q_rows, int nkv, int kv_rows, int variant,
                                 void* stream) {
  int rc = check_device();
  if (rc != LD_OK) return rc;
  LD_CHECK_ARG(q && k && v && out, "ld_attention_bf16: null pointer");
  LD_CHECK_ARG(batch > 0 && heads > 0 && nq > 0 && nkv > 0, "ld_attention_bf16: empty problem");
  LD_CHECK_ARG(nq <= q_rows && nkv <= kv_rows, "ld_attention_bf16: nq/nkv exceed buffer rows");
  LD_CHECK_ARG(out_f32 == nullptr || lse != nullptr, "ld_attention_bf16: out_f32 requires lse");
  const int BH = batch * heads;
  CUtensorMap tq, tk, tv;
  const uint32_t box[3] = {64, 128, 1};
  {
    // dims limited to the rows actually used so TMA zero-fills the tail
    const uint64_t dims[3] = {64, (uint64_t)nq, (uint64_t)BH};
    const uint64_t str[2] = {128, (uint64_t)q_rows * 128};
    rc = make_tmap_bf16(&tq, q, 3, dims, str, box);
    if (rc != LD_OK) return rc;
  }
  {
    const uint64_t dims[3] = {64, (uint64_t)nkv, (uint64_t)BH};
    const uint64_t str[2] = {128, (uint64_t)kv_rows * 128};
    rc = make_tmap_bf16(&tk, k, 3, dims, str, box);
    if (rc != LD_OK) return rc;
    rc = make_tmap_bf16(&tv, v, 3, dims, str, box);
    if (rc != LD_OK) return rc;
  }
  AttnParams prm;
  prm.out = (bf16*)out;
  prm.lse = lse;
  prm.out_f32 = out_f32;
  prm.heads = heads;
  prm.nq = nq;
  prm.nkv = nkv;
  prm.scale_log2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
  const int grid = BH * ((nq + 255) / 256);
  // variant: bit 0 = P through shared memory instead of TMEM; bits 1.. = exponential split
  //   0: default (P in TMEM, every 4th exp on the FMA pipe)   1: P via smem, all MUFU
  //   2: P in TMEM, all MUFU   4: every 3rd exp polynomial   6: every 2nd   8: every 4th (same as 0)
  cudaStream_t st = (cudaStream_t)stream;
  switch (variant) {
    case 0:
    case 8: return launch_attn<true, 4>(tq, tk, tv, prm, grid, st);
    case 1: return launch_attn<false, 0>(tq, tk, tv, prm, grid, st);
    case 2: return launch_attn<true, 0>(tq, tk, tv, prm, grid, st);
    case 4: return launch_attn<true, 3>(tq, tk, tv, prm, grid, st);
    case 6: return launch_attn<true, 2>(tq, tk, tv, prm, grid, st);
    default:
      set_error("ld_attention_bf16: unknown variant %d", variant);
      return LD_ERR_ARG;
  }
}

extern "C" int ld_attention_merge(float* o_acc, float* lse_acc, const float* o_new, const float* lse_new, void* out_bf16,
                                  int batch, int heads, int nq, void* stream) {
  int rc = check_device();
  if (rc != LD_OK) return rc;
  LD_CHECK_ARG(o_acc && lse_acc && o_new && lse_new, "ld_attention_merge: null pointer");
  const int64_t rows = (int64_t)batch * heads * nq;
  const int64_t threads = rows * 16;
  attn_merge_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(o_acc, lse_acc, o_new, lse_new,
                                                                                         (bf16*)out_bf16, heads, nq, rows);
  LD_CHECK_CUDA(cudaGetLastError());
  return LD_OK;
}
