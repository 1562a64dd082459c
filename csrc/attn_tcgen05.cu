// Full (non-causal) self-attention for head_dim 64 on sm_100a: flash-style online softmax with both
// contractions on tcgen05 tensor cores and Q / S / P / O resident in TMEM.  Replaces
// F.scaled_dot_product_attention as reached from dit_video_concat.py:655-664 (SAT attention_fn_default).
//
// Two kernels (history and measurements: DESIGN.md section 3a):
//   attn4_kernel   the product path: per CTA 256 query rows as two 128-row tiles, double-buffered 64-key score
//                  blocks, 16 softmax warps in column-split pairs, Q in TMEM (TS-mode S = Q K^T), two MMA issuer
//                  warps, one fixed reference maximum per row (first sub-block) with overflow detection.
//   attn3_kernel   the exact kernel (per-block maxima, lazy rescaling, four independent softmax streams); launched
//                  after attn4_kernel to recompute only the CTAs that flagged an overflow, and on its own as
//                  variant 1.
// K/V tail columns are masked to -inf; TMA zero-fills out-of-range rows.
#include <cstdlib>
#include "host_util.h"
#include "ptx.cuh"

namespace ld {

using bf16 = __nv_bfloat16;

constexpr int kTileBytes = 128 * 64 * 2;     // 16 KB: one 128x64 bf16 box
constexpr float kRescaleThreshold = 8.0f;    // log2 units (attn3_kernel's lazy rescaling)

struct AttnParams {
  bf16* out;        // [B, nq, heads*64]
  float* lse;       // [BH, nq] or null
  float* out_f32;   // [BH, nq, 64] or null
  int heads, nq, nkv;
  float scale_log2;
  const bf16* q;     // [BH, q_rows, 64] (attn4_kernel reads Q rows directly)
  int q_rows;
  int* redo;         // [grid] attn4_kernel: 1 = recompute this CTA with the exact kernel
  int redo_only;     // attn3_kernel: run only the CTAs with redo[cta] != 0
  long long* prof;   // per-(CTA, warp) phase cycle counters (profiling variants only)
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA/ALU pipes (Cody-Waite range reduction + degree-3 minimax polynomial, max rel. error 8.8e-5 — far
// below the bf16 rounding of P): floor via round-down magic add, fraction in [0,1), exponent re-inserted by an
// integer add.  Offloads part of the softmax exponentials from the 16-op/clk MUFU unit.  x >= 128 yields a NaN
// bit pattern (exponent field 255), which attn4_kernel's overflow detection catches like MUFU's +inf.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fminf(fmaxf(x, -127.0f), 128.0f);
  float xr;
  asm("add.rm.ftz.f32 %0, %1, %2;" : "=f"(xr) : "f"(x), "f"(12582912.0f));  // 1.5 * 2^23: low mantissa bits = floor(x)
  const float f = x - (xr - 12582912.0f);
  float p = fmaf(f, 0.077119089663028717f, 0.227564394474029541f);
  p = fmaf(p, f, 0.695146143436431885f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xr) << 23));
}

// POLY_EVERY = n > 0: every n-th exponential of a row goes to ex2_poly, the others to MUFU.EX2; 0: all MUFU.
// PROF kernels accumulate per-phase cycle counters (tools/attn_phase_prof.py).
#define LD_PROF(slot)                                  \
  if constexpr (PROF) {                                \
    const long long now = clock64();                   \
    prof_acc[slot] += now - tp;                        \
    tp = now;                                          \
  }

// ------------------------------------------------------------------------------------------------------------
// Third-generation kernel: FOUR independent online-softmax streams per CTA, 16 softmax warps (4 per SMSP).
//
// Measured on B200 (tools/softmax_mix_bench.cu): one warp alone runs the exponential mix (FFMA, MUFU.EX2, FADD,
// F2FP) at ~12 clk/element because its in-order issue cannot overlap MUFU with the dependent ops; two warps per
// SMSP reach 8.5 and three or more 8.2 — the MUFU limit (8).  So the kernel needs >= 3 warps per SMSP that have
// exponentials to do at any time.  TMEM (512 columns) cannot hold more than two 128-row query tiles with their
// accumulators, so each query tile is served by TWO streams that split the keys by 64-key sub-block parity:
// stream (t, b) owns sub-blocks i = b, b+2, ... of query tile t with its own score buffer S, its own output
// accumulator O and its own running (max, sum); the two streams of a tile are merged by log-sum-exp in the
// epilogue.  While one stream waits for the tensor pipe (PV then the next S) the other three keep MUFU busy.
//   warp 0: TMA producer   warp 1: MMA issuer   (2, 3 idle)   warps 4-19: softmax, stream = (warp-4)/4, TMEM
//   quadrant = warp%4.  640 threads start at 96 registers; the control warpgroup drops to 48 and the softmax
//   warpgroups rise to 104 (setmaxnreg.inc only draws on registers released inside the CTA: 6144 >= 4096).
//   TMEM columns: S/P(stream) at 64*stream [0,256) ; O(stream) at 256 + 64*stream [256,512)
constexpr int kAttn3Threads = 640;
constexpr int kKS3 = 4;
constexpr int kAttn3Smem = 2 * kTileBytes + 2 * kKS3 * kTileBytes + 2 * 128 * 2 * 8 /*(m,l) exchange*/ + 1024 + 256;

template <int POLY_EVERY, bool PROF>
__global__ void __launch_bounds__(kAttn3Threads, 1)
attn3_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
             const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                // 2 tiles
  uint8_t* sK = sQ + 2 * kTileBytes;                 // kKS3 tiles
  uint8_t* sV = sK + kKS3 * kTileBytes;              // kKS3 tiles
  float2* sML = reinterpret_cast<float2*>(sV + kKS3 * kTileBytes);   // [stream][128] (m * scale_log2, l)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sML + 4 * 128);
  uint64_t* q_full = bars;               // 1
  uint64_t* k_full = bars + 1;           // kKS3
  uint64_t* k_empty = k_full + kKS3;
  uint64_t* v_full = k_empty + kKS3;
  uint64_t* v_empty = v_full + kKS3;
  uint64_t* s_full = v_empty + kKS3;     // [stream] = 4
  uint64_t* p_full = s_full + 4;         // [stream] = 4
  uint64_t* o_done = p_full + 4;         // [stream] = 4
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 4);

  // fix-up launch after attn4_kernel: only the CTAs whose fixed reference maximum overflowed are recomputed
  if (p.redo_only && p.redo[blockIdx.x] == 0) return;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_blocks = (p.nq + 255) / 256;
  const int bh = blockIdx.x / q_blocks;
  const int q0 = (blockIdx.x % q_blocks) * 256;
  const int n_tiles = (p.nkv + 127) / 128;   // TMA boxes
  const int n_sub = (p.nkv + 63) / 64;       // 64-key sub-blocks

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKS3; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Warps 0 and 1 run warp-uniform code and issue through an elect.sync leader: ptxas then emits back-to-back
  // UTMALDG / UTCHMMA instead of a per-instruction divergence loop (measured: ~100 clk -> 32 clk per MMA).
  if (warp < 4) reg_dealloc<48>();
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(q_full, 2 * kTileBytes);
      tma_load_3d(sQ, &tmap_q, q_full, 0, q0, bh);
      tma_load_3d(sQ + kTileBytes, &tmap_q, q_full, 0, q0 + 128, bh);
    }
    int s = 0;
    uint32_t ph = 0;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(&k_empty[s], ph ^ 1);
      if (leader) {
        mbar_expect_tx(&k_full[s], kTileBytes);
        tma_load_3d(sK + s * kTileBytes, &tmap_k, &k_full[s], 0, j * 128, bh);
      }
      mbar_wait(&v_empty[s], ph ^ 1);
      if (leader) {
        mbar_expect_tx(&v_full[s], kTileBytes);
        tma_load_3d(sV + s * kTileBytes, &tmap_v, &v_full[s], 0, j * 128, bh);
      }
      if (++s == kKS3) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 64);
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, false, true);  // B (=V) is N-major
    const uint64_t qdesc = make_sdesc_sw128(smem_u32(sQ));
    const uint64_t kdesc = make_sdesc_sw128(smem_u32(sK));
    const uint64_t vdesc = make_sdesc_sw128(smem_u32(sV));
    // descriptor address units are 16 B: ring stage = 1024, 64-row half = 512, query tile = 1024
    // S(t, i) = Q_t K_i^T into the score buffer of stream 2t + (i&1); K sub-block i = rows [64*(i&1), +64) of ring
    // stage (i>>1) % kKS3
    auto issue_s = [&](int t, int i) {
      const uint64_t adesc = qdesc + uint32_t(t * (kTileBytes >> 4));
      const uint64_t bdesc = kdesc + uint32_t(((i >> 1) % kKS3) * (kTileBytes >> 4) + (i & 1) * (kTileBytes >> 5));
      const int st = 2 * t + (i & 1);
      const uint32_t d = tmem_base + st * 64;
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(d, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0);
        umma_commit(&s_full[st]);
      }
    };
    // O(stream) += P(t,i) V_i ; P: 128 lanes x 64 keys bf16 = 32 TMEM columns over S(stream)
    auto issue_pv = [&](int t, int i) {
      const uint64_t bdesc = vdesc + uint32_t(((i >> 1) % kKS3) * (kTileBytes >> 4) + (i & 1) * (kTileBytes >> 5));
      const int st = 2 * t + (i & 1);
      const uint32_t d = tmem_base + 256 + st * 64;
      const uint32_t a = tmem_base + st * 64;
      if (leader) {
        // 16 keys per step: 16 rows x 128 B = 2048 B (encoded 128)
        umma_ts(d, a, bdesc, idesc_o, i >= 2);
#pragma unroll
        for (int k = 1; k < 4; ++k) umma_ts(d, a + k * 8, bdesc + 128 * k, idesc_o, 1u);
        umma_commit(&o_done[st]);
      }
    };
    auto k_wait = [&](int tile) { mbar_wait(&k_full[tile % kKS3], (tile / kKS3) & 1); };
    auto v_wait = [&](int tile) { mbar_wait(&v_full[tile % kKS3], (tile / kKS3) & 1); };

    mbar_wait(q_full, 0);
    k_wait(0);
    tc_fence_after();
    issue_s(0, 0);
    issue_s(1, 0);
    if (n_sub > 1) {
      issue_s(0, 1);
      issue_s(1, 1);
    }
    if (leader) umma_commit(&k_empty[0]);
    long long w_kv = 0, w_p0 = 0, w_p1 = 0, t_all = 0;
    if constexpr (PROF) t_all = clock64();
    for (int i = 0; i < n_sub; ++i) {
      const int b = i & 1;
      const uint32_t par = (i >> 1) & 1;
      const bool has_next = (i + 2) < n_sub;
      long long c0 = 0;
      if constexpr (PROF) c0 = clock64();
      if (b == 0) v_wait(i >> 1);
      if (has_next && b == 0) k_wait((i + 2) >> 1);
      if constexpr (PROF) { const long long c1 = clock64(); w_kv += c1 - c0; c0 = c1; }
      // tile 0
      mbar_wait(&p_full[0 + b], par);
      if constexpr (PROF) { const long long c1 = clock64(); w_p0 += c1 - c0; c0 = c1; }
      tc_fence_after();
      issue_pv(0, i);
      if (has_next) issue_s(0, i + 2);
      // tile 1
      if constexpr (PROF) c0 = clock64();
      mbar_wait(&p_full[2 + b], par);
      if constexpr (PROF) w_p1 += clock64() - c0;
      tc_fence_after();
      issue_pv(1, i);
      if (leader && (b == 1 || i == n_sub - 1)) umma_commit(&v_empty[(i >> 1) % kKS3]);
      if (has_next) {
        issue_s(1, i + 2);
        if (leader && (b == 1 || i + 2 == n_sub - 1)) umma_commit(&k_empty[((i + 2) >> 1) % kKS3]);
      }
    }
    if constexpr (PROF) {
      if (leader && p.prof != nullptr) {
        long long* d = p.prof + ((int64_t)blockIdx.x * 20 + 1) * 8;
        d[0] = w_kv; d[1] = w_p0; d[2] = w_p1; d[3] = clock64() - t_all;
      }
    }
  } else if (warp >= 4) {
    reg_alloc<104>();   // 16 warps x 32 x 8 = 4096 <= the 6144 registers released by the control warpgroup
    // -------------------------------------------------------------------- softmax / correction / epilogue
    const int st = (warp - 4) >> 2;       // stream
    const int t = st >> 1;                // query tile
    const int b = st & 1;                 // sub-block parity served by this stream
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const int q_row = q0 + t * 128 + row_in_tile;
    const uint32_t lane_base = uint32_t(quad * 32) << 16;
    const uint32_t ts = tmem_base + lane_base + st * 64;         // S / P of this stream
    const uint32_t to = tmem_base + lane_base + 256 + st * 64;   // O of this stream
    const float sl2 = p.scale_log2;

    float m_used = -INFINITY;  // raw-score units
    float l = 0.f;
    long long prof_acc[6] = {0, 0, 0, 0, 0, 0};
    long long tp = 0;
    if constexpr (PROF) tp = clock64();
    int kk = 0;   // blocks done by this stream
    for (int i = b; i < n_sub; i += 2, ++kk) {
      mbar_wait(&s_full[st], kk & 1);
      tc_fence_after();
      LD_PROF(0);
      uint32_t s[64];
      LD_TMEM_LD32(ts + 0, (s + 0));
      LD_TMEM_LD32(ts + 32, (s + 32));
      tmem_ld_wait();
      LD_PROF(1);
      const int valid = p.nkv - i * 64;
      if (valid < 64) {
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (c >= valid) s[c] = 0xff800000u;  // -inf
      }
      float mx4[4] = {__uint_as_float(s[0]), __uint_as_float(s[1]), __uint_as_float(s[2]), __uint_as_float(s[3])};
#pragma unroll
      for (int c = 4; c < 64; c += 4) {
        mx4[0] = fmaxf(mx4[0], __uint_as_float(s[c]));
        mx4[1] = fmaxf(mx4[1], __uint_as_float(s[c + 1]));
        mx4[2] = fmaxf(mx4[2], __uint_as_float(s[c + 2]));
        mx4[3] = fmaxf(mx4[3], __uint_as_float(s[c + 3]));
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));

      bool need = (kk > 0) && ((mx - m_used) * sl2 > kRescaleThreshold);
      if (kk == 0) m_used = mx;
      if (__any_sync(0xffffffffu, need)) {
        // lazy correction: bring O and l of this stream to the new reference maximum (whole warp: tcgen05.ld/st are
        // collective).  s_full(kk) completing implies PV(kk-1) completed (commit order), so O is quiescent.
        const float m_new = fmaxf(m_used, mx);
        const float alpha = ex2((m_used - m_new) * sl2);
#pragma unroll 1
        for (int c = 0; c < 64; c += 8) {   // rare path: 8 columns at a time keeps it out of the main loop's registers
          uint32_t o[8];
          LD_TMEM_LD8(to + c, o);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
          LD_TMEM_ST8(to + c, o);
        }
        l *= alpha;
        m_used = m_new;
      }
      const float msc = m_used * sl2;
      LD_PROF(2);
      float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        float pv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int idx = 4 * c + e;
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
        pk[2 * c] = pack_bf16x2(pv[0], pv[1]);
        pk[2 * c + 1] = pack_bf16x2(pv[2], pv[3]);
      }
      l += (sum0 + sum1) + (sum2 + sum3);
      LD_PROF(3);
      LD_TMEM_ST32(ts, pk);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[st]);
      LD_PROF(4);
    }

    // ---- epilogue: merge the two streams of this query tile by log-sum-exp; stream b writes output columns
    //      [32b, 32b+32) of the head
    if (kk > 0) {
      mbar_wait(&o_done[st], (kk - 1) & 1);
      tc_fence_after();
    }
    if constexpr (PROF) {
      if (lane == 0 && p.prof != nullptr) {
        long long* d = p.prof + ((int64_t)blockIdx.x * 20 + warp) * 8;
        for (int e = 0; e < 6; ++e) d[e] = prof_acc[e];
      }
    }
    sML[st * 128 + row_in_tile] = make_float2(m_used * sl2, l);
    tc_fence_before();
    named_bar_sync(1 + t, 256);     // the 8 warps of query tile t
    tc_fence_after();
    const float2 other = sML[(st ^ 1) * 128 + row_in_tile];
    const bool other_has = n_sub > (b ^ 1);   // the sibling stream processed at least one sub-block (uniform)
    const float m_mine = m_used * sl2, m_oth = other_has ? other.x : -INFINITY;
    const float m_all = fmaxf(m_mine, m_oth);
    const float w_mine = (kk > 0) ? ex2(m_mine - m_all) : 0.f;
    const float w_oth = other_has ? ex2(m_oth - m_all) : 0.f;
    const float l_all = w_mine * l + w_oth * (other_has ? other.y : 0.f);
    const float inv_l = 1.0f / l_all;
    const float f_mine = w_mine * inv_l, f_oth = w_oth * inv_l;
    const bool valid_row = q_row < p.nq;
    const int bb = bh / p.heads, h = bh - bb * p.heads;
    const int c0 = 32 * b;
    uint32_t oa[32], ob[32];
    const uint32_t to_oth = tmem_base + lane_base + 256 + (st ^ 1) * 64;
    if (kk > 0) {
      LD_TMEM_LD32(to + c0, oa);
    }
    if (other_has) {
      LD_TMEM_LD32(to_oth + c0, ob);
    }
    tmem_ld_wait();
    if (valid_row) {
      bf16* orow = p.out + ((int64_t)bb * p.nq + q_row) * (p.heads * 64) + h * 64 + c0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float v = 0.f;
          if (kk > 0) v = __uint_as_float(oa[g * 8 + e]) * f_mine;
          if (other_has) v = fmaf(__uint_as_float(ob[g * 8 + e]), f_oth, v);
          f[e] = v;
        }
        uint4 v;
        v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
        v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
        *reinterpret_cast<uint4*>(orow + g * 8) = v;
        if (p.out_f32 != nullptr) {
          float* of = p.out_f32 + ((int64_t)bh * p.nq + q_row) * 64 + c0 + g * 8;
          *reinterpret_cast<float4*>(of) = make_float4(f[0], f[1], f[2], f[3]);
          *reinterpret_cast<float4*>(of + 4) = make_float4(f[4], f[5], f[6], f[7]);
        }
      }
      if (b == 0 && p.lse != nullptr) p.lse[(int64_t)bh * p.nq + q_row] = m_all + log2f(l_all);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}


// ------------------------------------------------------------------------------------------------------------
// Fourth-generation kernel: double-buffered 64-key score blocks + 16 softmax warps + Q in TMEM,
// and NO per-block row maximum.
//
// Two 128-row query tiles per CTA; each tile has two 64-column score buffers in TMEM so S(t,i+2) is issued right
// after PV(t,i) and the softmax never waits for the tensor pipe.  Each 32-row TMEM quadrant of a tile is served by
// a PAIR of warps that split the 64 columns of a sub-block (warp h exponentiates columns [32h, 32h+32)), which
// puts 4 warps with exponentials on every SMSP (the MUFU unit needs >= 3 to saturate, tools/softmax_mix_bench.cu).
//
// Reference maximum: floating point is scale-invariant, so the online-softmax reference only has to prevent
// overflow, not track the running maximum.  Each row takes the maximum of its FIRST sub-block as the reference
// for the whole row (both warps of a pair load that sub-block entirely, so they agree without communicating) and
// never rescales: later scores may exceed the reference by up to 2^127 before exp2 overflows, and terms far below
// it flush to zero exactly as their true weight demands.  That removes the 64 FMNMX + vote + correction logic per
// row and sub-block.  Overflow (a score more than ~127 log2-units above the first block's maximum — never seen on
// LayerNormed q/k, but constructible) makes the row sum or the output non-finite; the CTA then raises redo[cta],
// and a second launch of the exact kernel (attn3_kernel, per-block maxima and rescaling) recomputes only the
// flagged CTAs.  tests/test_kernels_gpu.py::test_attention_overflow_fixup covers that path.
//
// Q is stored once into TMEM (bf16 pairs, one row per lane) by the softmax threads, so S = Q K^T runs as a TS-mode
// MMA whose only shared-memory operand is the K sub-block: 32 clk per 128x64x16 instead of 48 (tools/mma_bench.cu).
// P has its own 32-column buffer per tile; a warp waits for PV(t,i-1) before overwriting it.
//   warps 0-15: softmax (tile w>>3, column half (w>>2)&1, TMEM quadrant w&3)   (16 idle)   warp 17: TMA producer
//   warps 18, 19: MMA issuers of query tile 0 and 1.  One issuer for both tiles spends ~1900 clk per sub-block in
//   its serial scalar code (barrier polls, descriptor arithmetic, R2UR) and was the bottleneck; the two tiles are
//   independent instruction streams for the tensor pipe, so each gets its own issuing warp (K/V ring stages are
//   released by one commit from each).
//   TMEM columns: S(t,b) at 64(2t+b) [0,256) ; O_t at 256+64t [256,384) ; Q_t at 384+32t [384,448) ;
//                 P_t at 448+32t [448,512)
constexpr int kAttn4Threads = 640;
constexpr int kKS4 = 4;
constexpr int kAttn4Smem = 2 * kKS4 * kTileBytes + 2 * 2 * 128 * 4 /* row-sum exchange */ + 1024 + 256;

template <int POLY_EVERY, bool PROF, bool PACKED>
__global__ void __launch_bounds__(kAttn4Threads, 1)
attn4_kernel(const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
             const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;                                // kKS4 tiles
  uint8_t* sV = sK + kKS4 * kTileBytes;              // kKS4 tiles
  float* sL = reinterpret_cast<float*>(sV + kKS4 * kTileBytes);   // [tile][half][128] partial row sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(sL + 2 * 2 * 128);
  uint64_t* q_ready = bars;              // [tile] = 2
  uint64_t* k_full = bars + 2;           // kKS4
  uint64_t* k_empty = k_full + kKS4;
  uint64_t* v_full = k_empty + kKS4;
  uint64_t* v_empty = v_full + kKS4;
  uint64_t* s_full = v_empty + kKS4;     // [t][b] = 4
  uint64_t* p_full = s_full + 4;         // [t][b] = 4
  uint64_t* o_done = p_full + 4;         // [t] = 2
  uint64_t* all_done = o_done + 2;       // 1: every MMA of this CTA has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(all_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_blocks = (p.nq + 255) / 256;
  const int bh = blockIdx.x / q_blocks;
  const int q0 = (blockIdx.x % q_blocks) * 256;
  const int n_tiles = (p.nkv + 127) / 128;   // TMA boxes
  const int n_sub = (p.nkv + 63) / 64;       // 64-key sub-blocks

  if (warp == 17 && lane == 0) {
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    for (int s = 0; s < kKS4; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 2);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 2);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 256);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_ready[i], 128);
      mbar_init(&o_done[i], 1);
    }
    mbar_init(all_done, 2);
    fence_barrier_init();
  }
  if (warp == 19) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  bool bad = false;   // softmax threads: non-finite row sum or output (overflow of the fixed reference maximum)

  // The producer and issuer warps run warp-uniform code and issue through an elect.sync leader: ptxas then emits
  // back-to-back UTMALDG / UTCHMMA instead of a per-instruction divergence loop (measured: ~100 clk -> 32 clk).
  if (warp >= 16) reg_dealloc<56>();   // releases 4 x 32 x 40 = 5120 registers
  if (warp == 17) {
    // ------------------------------------------------------------------ TMA producer
    const bool leader = elect_one();
    int s = 0;
    uint32_t ph = 0;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(&k_empty[s], ph ^ 1);
      if (leader) {
        mbar_expect_tx(&k_full[s], kTileBytes);
        tma_load_3d(sK + s * kTileBytes, &tmap_k, &k_full[s], 0, j * 128, bh);
      }
      mbar_wait(&v_empty[s], ph ^ 1);
      if (leader) {
        mbar_expect_tx(&v_full[s], kTileBytes);
        tma_load_3d(sV + s * kTileBytes, &tmap_v, &v_full[s], 0, j * 128, bh);
      }
      if (++s == kKS4) { s = 0; ph ^= 1; }
    }
  } else if (warp >= 18) {
    // ------------------------------------------------------------------ MMA issuer of query tile t
    const int t = warp - 18;
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 64);
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, false, true);  // B (=V) is N-major
    const uint64_t kdesc = make_sdesc_sw128(smem_u32(sK));
    const uint64_t vdesc = make_sdesc_sw128(smem_u32(sV));
    const uint32_t tm_s = tmem_base + 128 * t;        // S(t,0); S(t,1) 64 columns further
    const uint32_t tm_o = tmem_base + 256 + 64 * t;
    const uint32_t tm_q = tmem_base + 384 + 32 * t;
    const uint32_t tm_p = tmem_base + 448 + 32 * t;
    // descriptor address units are 16 B: ring stage = 1024, 64-row half = 512
    // S(t, i) = Q_t K_i^T into buffer i&1; Q_t from TMEM (8 columns per 16-dim K step)
    auto issue_s = [&](int i) {
      const uint64_t bdesc = kdesc + uint32_t(((i >> 1) % kKS4) * (kTileBytes >> 4) + (i & 1) * (kTileBytes >> 5));
      const uint32_t d = tm_s + (i & 1) * 64;
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ts(d, tm_q + 8 * k, bdesc + 2 * k, idesc_s, k != 0);
        umma_commit(&s_full[2 * t + (i & 1)]);
      }
    };
    // O_t += P(t,i) V_i ; P: 128 lanes x 64 keys bf16 = the tile's 32-column P buffer
    auto issue_pv = [&](int i) {
      const uint64_t bdesc = vdesc + uint32_t(((i >> 1) % kKS4) * (kTileBytes >> 4) + (i & 1) * (kTileBytes >> 5));
      if (leader) {
        // 16 keys per step: 16 rows x 128 B = 2048 B (encoded 128)
        umma_ts(tm_o, tm_p, bdesc, idesc_o, i != 0);
#pragma unroll
        for (int k = 1; k < 4; ++k) umma_ts(tm_o, tm_p + k * 8, bdesc + 128 * k, idesc_o, 1u);
        umma_commit(&o_done[t]);
      }
    };
    auto k_wait = [&](int tile) { mbar_wait(&k_full[tile % kKS4], (tile / kKS4) & 1); };
    auto v_wait = [&](int tile) { mbar_wait(&v_full[tile % kKS4], (tile / kKS4) & 1); };

    mbar_wait(&q_ready[t], 0);
    k_wait(0);
    tc_fence_after();
    issue_s(0);
    if (n_sub > 1) issue_s(1);
    if (leader) umma_commit(&k_empty[0]);
    long long w_kv = 0, w_p0 = 0, t_all = 0, w_ipv = 0, w_is = 0;
    if constexpr (PROF) t_all = clock64();
    for (int i = 0; i < n_sub; ++i) {
      const int b = i & 1;
      const bool has_next = (i + 2) < n_sub;
      long long c0 = 0;
      if constexpr (PROF) c0 = clock64();
      if (b == 0) v_wait(i >> 1);
      if (has_next && b == 0) k_wait((i + 2) >> 1);
      if constexpr (PROF) { const long long c1 = clock64(); w_kv += c1 - c0; c0 = c1; }
      mbar_wait(&p_full[2 * t + b], (i >> 1) & 1);
      if constexpr (PROF) { const long long c1 = clock64(); w_p0 += c1 - c0; c0 = c1; }
      tc_fence_after();
      issue_pv(i);
      if constexpr (PROF) { const long long c1 = clock64(); w_ipv += c1 - c0; c0 = c1; }
      if (leader && (b == 1 || i == n_sub - 1)) umma_commit(&v_empty[(i >> 1) % kKS4]);
      if (has_next) {
        issue_s(i + 2);
        if (leader && (b == 1 || i + 2 == n_sub - 1)) umma_commit(&k_empty[((i + 2) >> 1) % kKS4]);
      }
      if constexpr (PROF) w_is += clock64() - c0;
    }
    if (leader) umma_commit(all_done);
    if constexpr (PROF) {
      if (leader && p.prof != nullptr) {
        long long* d = p.prof + ((int64_t)blockIdx.x * 20 + warp) * 8;
        d[0] = w_kv; d[1] = w_p0; d[2] = 0; d[3] = clock64() - t_all; d[4] = w_ipv; d[5] = w_is;
      }
    }
  } else if (warp < 16) {
    reg_alloc<104>();   // 16 warps x 32 x 8 = 4096 <= the 5120 registers released by the control warpgroup
    // -------------------------------------------------------------------- softmax / correction / epilogue
    const int t = warp >> 3;                    // query tile
    const int h = (warp >> 2) & 1;              // column half of every sub-block exponentiated by this warp
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const int q_row = q0 + t * 128 + row_in_tile;
    const uint32_t lane_base = uint32_t(quad * 32) << 16;
    const uint32_t ts = tmem_base + lane_base + t * 128;         // S(t,0); S(t,1) is 64 columns further
    const uint32_t to = tmem_base + lane_base + 256 + t * 64;    // O_t
    const float sl2 = p.scale_log2;

    if (h == 0) {
      // Q row -> TMEM (bf16 pairs: column c holds dims 2c, 2c+1), zero beyond nq
      uint32_t qr[32];
      if (q_row < p.nq) {
        const uint4* src = reinterpret_cast<const uint4*>(p.q + ((int64_t)bh * p.q_rows + q_row) * 64);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 v = __ldg(src + c);
          qr[4 * c] = v.x; qr[4 * c + 1] = v.y; qr[4 * c + 2] = v.z; qr[4 * c + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) qr[c] = 0u;
      }
      LD_TMEM_ST32(tmem_base + lane_base + 384 + 32 * t, qr);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&q_ready[t]);
    }

    float msc = 0.f;   // reference maximum of the row (first sub-block) * scale_log2
    float l = 0.f;
    long long prof_acc[6] = {0, 0, 0, 0, 0, 0};
    long long tp = 0;
    if constexpr (PROF) tp = clock64();
    const uint32_t tp_addr = tmem_base + lane_base + 448 + 32 * t + 16 * h;
    bool s_ready = false;   // s_full(i) already observed complete by the probe of the previous iteration
    for (int i = 0; i < n_sub; ++i) {
      const int b = i & 1;
      const uint32_t tsb = ts + b * 64;
      if (!s_ready) mbar_wait(&s_full[2 * t + b], (i >> 1) & 1);
      tc_fence_after();
      LD_PROF(0);
      uint32_t s[32];   // this warp's column half
      LD_TMEM_LD32(tsb + 32 * h, s);
      const int valid = p.nkv - i * 64 - 32 * h;
      if (i == 0) {
        // reference maximum of the row = maximum of the whole first sub-block (identical in both warps of the pair)
        uint32_t so[32];
        LD_TMEM_LD32(tsb + 32 * (h ^ 1), so);
        tmem_ld_wait();
        const int valid_o = p.nkv - 32 * (h ^ 1);
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          if (c < valid) mx = fmaxf(mx, __uint_as_float(s[c]));
          if (c < valid_o) mx = fmaxf(mx, __uint_as_float(so[c]));
        }
        msc = mx * sl2;
      } else {
        tmem_ld_wait();
      }
      LD_PROF(1);
      if (valid < 32) {
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (c >= valid) s[c] = 0xff800000u;  // -inf
      }
      LD_PROF(2);
      // Probe the two barriers the end of this iteration and the start of the next one depend on now, so their
      // ~100-clk round trips overlap the exponentials (both are almost always complete already).
      const bool pv_done = (i == 0) || mbar_test_wait(&o_done[t], (i - 1) & 1);
      s_ready = (i + 1 < n_sub) && mbar_test_wait(&s_full[2 * t + (b ^ 1)], ((i + 1) >> 1) & 1);
      uint32_t pk[16];
      if constexpr (PACKED) {
        // packed f32x2 FMA / ADD: same FMA-pipe throughput but half the issue slots for the scale-subtract and the row
        // sums, which is what makes the polynomial split pay (tools/softmax_mix_bench.cu: 6.5 clk/element with every
        // 4th exponential on the FMA pipe at 4 warps per SMSP, against 8.2 for the plain mix)
        uint64_t sc2, nm2, sum01 = 0, sum23 = 0;
        asm("mov.b64 %0, {%1, %1};" : "=l"(sc2) : "f"(sl2));
        asm("mov.b64 %0, {%1, %1};" : "=l"(nm2) : "f"(-msc));
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint64_t a01, a23;
          asm("mov.b64 %0, {%1, %2};" : "=l"(a01) : "r"(s[4 * c]), "r"(s[4 * c + 1]));
          asm("mov.b64 %0, {%1, %2};" : "=l"(a23) : "r"(s[4 * c + 2]), "r"(s[4 * c + 3]));
          asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a01) : "l"(sc2), "l"(nm2));
          asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a23) : "l"(sc2), "l"(nm2));
          float x[4], pv[4];
          asm("mov.b64 {%0, %1}, %2;" : "=f"(x[0]), "=f"(x[1]) : "l"(a01));
          asm("mov.b64 {%0, %1}, %2;" : "=f"(x[2]), "=f"(x[3]) : "l"(a23));
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int idx = 4 * c + e;
            if constexpr (POLY_EVERY > 0) {
              pv[e] = ((idx % POLY_EVERY) == POLY_EVERY - 1) ? ex2_poly(x[e]) : ex2(x[e]);
            } else {
              pv[e] = ex2(x[e]);
            }
          }
          uint64_t p01, p23;
          asm("mov.b64 %0, {%1, %2};" : "=l"(p01) : "f"(pv[0]), "f"(pv[1]));
          asm("mov.b64 %0, {%1, %2};" : "=l"(p23) : "f"(pv[2]), "f"(pv[3]));
          asm("add.rn.f32x2 %0, %0, %1;" : "+l"(sum01) : "l"(p01));
          asm("add.rn.f32x2 %0, %0, %1;" : "+l"(sum23) : "l"(p23));
          pk[2 * c] = pack_bf16x2(pv[0], pv[1]);
          pk[2 * c + 1] = pack_bf16x2(pv[2], pv[3]);
        }
        float s0, s1, s2, s3;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(s0), "=f"(s1) : "l"(sum01));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(s2), "=f"(s3) : "l"(sum23));
        l += (s0 + s1) + (s2 + s3);
      } else {
        float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float pv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int idx = 4 * c + e;
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
          pk[2 * c] = pack_bf16x2(pv[0], pv[1]);
          pk[2 * c + 1] = pack_bf16x2(pv[2], pv[3]);
        }
        l += (sum0 + sum1) + (sum2 + sum3);
      }
      LD_PROF(3);
      if (!pv_done) mbar_wait(&o_done[t], (i - 1) & 1);   // PV(t,i-1) has consumed the P buffer
      LD_TMEM_ST16(tp_addr, pk);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[2 * t + b]);
      LD_PROF(4);
    }

    // ---- epilogue: total row sum = the two halves' partial sums (same reference maximum); warp h writes output
    //      columns [32h, 32h+32) of the head.  The softmax may run up to two PVs ahead of the tensor pipe, so a
    //      parity wait on o_done could alias here (both last phases may already have completed): the end of all
    //      MMAs has its own single-use barrier.
    mbar_wait(all_done, 0);
    tc_fence_after();
    if constexpr (PROF) {
      if (lane == 0 && p.prof != nullptr) {
        long long* d = p.prof + ((int64_t)blockIdx.x * 20 + warp) * 8;
        for (int e = 0; e < 6; ++e) d[e] = prof_acc[e];
      }
    }
    sL[(t * 2 + h) * 128 + row_in_tile] = l;
    named_bar_sync(1 + t, 256);     // the 8 warps of query tile t
    const float l_all = l + sL[(t * 2 + (h ^ 1)) * 128 + row_in_tile];
    const float inv_l = 1.0f / l_all;
    const bool valid_row = q_row < p.nq;
    const int bb = bh / p.heads, hd = bh - bb * p.heads;
    const int c0 = 32 * h;
    uint32_t o[32];
    LD_TMEM_LD32(to + c0, o);
    tmem_ld_wait();
    if (valid_row) {
      bad = !(l_all < INFINITY) || !(l_all > 0.f);   // inf / NaN row sum (or everything flushed to zero)
      bf16* orow = p.out + ((int64_t)bb * p.nq + q_row) * (p.heads * 64) + hd * 64 + c0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          f[e] = __uint_as_float(o[g * 8 + e]) * inv_l;
          bad |= !(fabsf(f[e]) < INFINITY);
        }
        uint4 v;
        v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
        v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
        *reinterpret_cast<uint4*>(orow + g * 8) = v;
        if (p.out_f32 != nullptr) {
          float* of = p.out_f32 + ((int64_t)bh * p.nq + q_row) * 64 + c0 + g * 8;
          *reinterpret_cast<float4*>(of) = make_float4(f[0], f[1], f[2], f[3]);
          *reinterpret_cast<float4*>(of + 4) = make_float4(f[4], f[5], f[6], f[7]);
        }
      }
      if (h == 0 && p.lse != nullptr) p.lse[(int64_t)bh * p.nq + q_row] = msc + log2f(l_all);
    }
  }

  tc_fence_before();
  const int any_bad = __syncthreads_or(bad ? 1 : 0);
  if (threadIdx.x == 0) p.redo[blockIdx.x] = any_bad;   // 1: the exact kernel recomputes this CTA's 256 rows
  if (warp == 19) tmem_dealloc<512>(tmem_base);
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

template <int POLY_EVERY, bool PROF = false>
static int launch_attn3(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& prm,
                        int grid, cudaStream_t st) {
  auto kern = attn3_kernel<POLY_EVERY, PROF>;
  static bool attr_set = false;
  if (!attr_set) {
    LD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttn3Smem));
    attr_set = true;
  }
  kern<<<grid, kAttn3Threads, kAttn3Smem, st>>>(tq, tk, tv, prm);
  LD_CHECK_CUDA(cudaGetLastError());
  return LD_OK;
}

template <int POLY_EVERY, bool PROF = false, bool PACKED = false>
static int launch_attn4(const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& prm, int grid,
                        cudaStream_t st) {
  auto kern = attn4_kernel<POLY_EVERY, PROF, PACKED>;
  static bool attr_set = false;
  if (!attr_set) {
    LD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttn4Smem));
    attr_set = true;
  }
  kern<<<grid, kAttn4Threads, kAttn4Smem, st>>>(tk, tv, prm);
  LD_CHECK_CUDA(cudaGetLastError());
  return LD_OK;
}

}  // namespace ld

using namespace ld;

// redo flags of the attn4 + fix-up pair: one int per CTA, per device, grown on demand (allocation only on first use
// or growth, so a warmed-up step allocates nothing)
static int* g_redo[64] = {nullptr};
static int g_redo_cap[64] = {0};
static int redo_buffer(int grid, int** out) {
  int dev = 0;
  LD_CHECK_CUDA(cudaGetDevice(&dev));
  LD_CHECK_ARG(dev >= 0 && dev < 64, "ld_attention_bf16: device index out of range");
  if (grid > g_redo_cap[dev]) {
    if (g_redo[dev] != nullptr) LD_CHECK_CUDA(cudaFree(g_redo[dev]));
    g_redo[dev] = nullptr;
    g_redo_cap[dev] = 0;
    const int cap = grid + grid / 2 + 1024;
    LD_CHECK_CUDA(cudaMalloc(&g_redo[dev], sizeof(int) * (size_t)cap));
    g_redo_cap[dev] = cap;
  }
  *out = g_redo[dev];
  return LD_OK;
}

static long long* g_attn_prof = nullptr;
// Debug hook (tools/attn_phase_prof.py): per-(CTA, warp) phase cycle counters for the next ld_attention_bf16 calls.
extern "C" void ld_debug_attn_prof(long long* buf) { g_attn_prof = buf; }

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
  prm.prof = nullptr;
  prm.redo = nullptr;
  prm.redo_only = 0;
  prm.q = (const bf16*)q;
  prm.q_rows = q_rows;
  prm.scale_log2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
  const int grid = BH * ((nq + 255) / 256);
  // variant 0: attn4_kernel (fixed first-block reference maximum) + exact fix-up launch   1: exact kernel only
  //         (attn3_kernel: per-block maxima, lazy rescaling)   2 / 3: attn4 with every 4th / 3rd exponential on the
  //         FMA pipe (polynomial)   4 / 5 / 6: attn4 with packed f32x2 FMA / ADD and every 4th / 8th / no polynomial
  cudaStream_t st = (cudaStream_t)stream;
  switch (variant) {
    case 1:
      if (g_attn_prof != nullptr) {
        prm.prof = g_attn_prof;
        return launch_attn3<0, true>(tq, tk, tv, prm, grid, st);
      }
      return launch_attn3<0>(tq, tk, tv, prm, grid, st);
    case 0:
    case 2:
    case 3:
    case 4:
    case 5:
    case 6: {
      rc = redo_buffer(grid, &prm.redo);
      if (rc != LD_OK) return rc;
      if (variant == 0 && g_attn_prof != nullptr) {
        prm.prof = g_attn_prof;
        rc = launch_attn4<0, true>(tk, tv, prm, grid, st);
      } else if (variant == 0) {
        rc = launch_attn4<0>(tk, tv, prm, grid, st);
      } else if (variant == 2) {
        rc = launch_attn4<4>(tk, tv, prm, grid, st);
      } else if (variant == 3) {
        rc = launch_attn4<3>(tk, tv, prm, grid, st);
      } else if (variant == 4) {
        rc = launch_attn4<4, false, true>(tk, tv, prm, grid, st);   // packed f32x2 + every 4th polynomial
      } else if (variant == 5) {
        rc = launch_attn4<8, false, true>(tk, tv, prm, grid, st);   // packed f32x2 + every 8th polynomial
      } else {
        rc = launch_attn4<0, false, true>(tk, tv, prm, grid, st);   // packed f32x2, all MUFU
      }
      if (rc != LD_OK) return rc;
      prm.prof = nullptr;
      prm.redo_only = 1;   // exact kernel, flagged CTAs only (all others exit at once)
      return launch_attn3<0>(tq, tk, tv, prm, grid, st);
    }
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
