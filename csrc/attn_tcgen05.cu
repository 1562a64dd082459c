// Full (non-causal) self-attention for head_dim 64 on sm_100a: flash-style online softmax with both contractions on
// tcgen05 tensor cores and Q / S / P / O / row sums resident in TMEM.  Replaces F.scaled_dot_product_attention as
// reached from dit_video_concat.py:655-664 (SAT attention_fn_default).
//
// K/V may arrive as up to four SHARDS (ring sequence parallelism: the local shard plus the peers' shards that the copy
// engines drop into IPC-mapped receive buffers): one launch walks all of them, accumulating in TMEM, and its TMA
// producer warp polls a per-shard arrival flag before the first load from a shard that is still in flight — no
// per-hop launches, no partial-result merges.
//
// One kernel, two code paths (history and measurements: DESIGN.md section 3a):
//   fast path   per CTA 256 query rows as two 128-row tiles, double-buffered 64-key score blocks, 16 softmax warps in
//               column-split pairs, Q in TMEM (TS-mode S = Q K^T), P written in place over S, row sums accumulated by
//               the tensor core (P x ones), a fraction of the exponentials on the FMA pipe (packed f32x2 polynomial),
//               ONE fixed reference maximum per row (first sub-block) with overflow detection.
//   exact path  per-block maxima with lazy rescaling, four independent softmax streams.  Runs in the SAME launch for
//               the (rare) CTAs whose fixed reference maximum overflowed, and on its own as variant 1.
// K/V tail columns are masked to -inf; TMA zero-fills out-of-range rows.
#include <cstdlib>
#include <cstring>
#include "host_util.h"
#include "ptx.cuh"

namespace ld {

using bf16 = __nv_bfloat16;

constexpr int kTileBytes = 128 * 64 * 2;     // 16 KB: one 128x64 bf16 box
constexpr float kRescaleThreshold = 8.0f;    // log2 units (exact path's lazy rescaling)
constexpr int kMaxShards = 4;
constexpr int kMaxSub = 2048;                // 64-key sub-blocks per launch: up to 131 072 keys over all shards
constexpr int kKS = 4;                       // K and V ring depth (128-key boxes)
constexpr int kAttnThreads = 640;
// Exponential pairs (of 16) on the FMA-pipe polynomial in the default variant.  In isolation (boost clock) 5 is the fastest;
// inside a sampler step the board runs at its power cap (~1.6 GHz) and the polynomial's extra FMA-pipe work costs clock:
// measured in-step 374.7 ms/step with 4 against 380.7 with 5 (profiles/r2_attn_kp_in_step.txt), sustained back-to-back
// launches 5.30 / 5.37 / 5.50 / 6.07 ms for 4 / 3 / 5 / 6 (profiles/r2_attn_sustained_bench.txt).
constexpr int kDefaultKP = 4;

struct alignas(64) AttnShards {
  CUtensorMap q;                         // [64, nq, BH], box 64 x 128 (exact path)
  CUtensorMap k[kMaxShards];             // [64, nkv_s, BH], box 64 x 128
  CUtensorMap v[kMaxShards];
  const uint32_t* ready[kMaxShards];     // arrival flag of shard s (null: already present): wait until
  uint32_t ready_val[kMaxShards];        //   (int32)(*ready[s] - ready_val[s]) >= 0   (acquire, system scope)
  int nkv[kMaxShards];
  int n;                                 // shards
  int n_sub;                             // 64-key sub-blocks over all shards
  int n_box;                             // 128-key TMA boxes over all shards
  uint32_t* status;                      // [0] |= 1 when a shard wait gave up after wait_cycles (result invalid)
  long long wait_cycles;
};

struct AttnParams {
  bf16* out;        // [B, nq, heads*64]
  float* lse;       // [BH, nq] or null
  float* out_f32;   // [BH, nq, 64] or null
  int heads, nq;
  float scale_log2;
  const bf16* q;     // [BH, q_rows, 64] (the fast path reads Q rows directly)
  int q_rows;
  int exact_only;    // variant 1: run the exact path for every CTA
  long long* prof;   // per-(CTA, warp) phase cycle counters (profiling build of the fast path only)
  // Tail split (wave quantisation): CTAs [0, n_main) each own one (head, 256-row query block) over ALL keys; the remaining
  // query blocks — the last, partial wave of the grid — are each served by n_split CTAs that take 1/n_split of the key
  // boxes and write normalised fp32 partial outputs + log-sum-exp, merged by attn_split_merge_kernel.  n_split <= 1: off.
  int n_main, n_split;
  float* part_o;     // [tail block][split][256][64]
  float* part_lse;   // [tail block][split][256]
};

// What one CTA works on: logical (head, query block) index, its range of key boxes, where its result goes.
struct AttnWork {
  int cta;           // logical work index: bh * q_blocks + query block
  int gb0, gb1;      // global box range [gb0, gb1) over the concatenated shards
  int n_sub;         // 64-key sub-blocks inside the range
  int64_t part_row;  // >= 0: first row of this CTA's slot in part_o / part_lse; < 0: final output
};

__device__ __forceinline__ int range_sub_blocks(const AttnShards& sh, int gb0, int gb1) {
  int n = 0, box0 = 0;
  for (int s = 0; s < sh.n; ++s) {
    const int nb = (sh.nkv[s] + 127) >> 7, ns = (sh.nkv[s] + 63) >> 6;
    const int lo = max(gb0, box0), hi = min(gb1, box0 + nb);
    if (hi > lo) n += min(ns, 2 * (hi - box0)) - 2 * (lo - box0);
    box0 += nb;
  }
  return n;
}

__device__ __forceinline__ AttnWork attn_work(const AttnShards& sh, const AttnParams& p) {
  AttnWork w;
  const int x = blockIdx.x;
  if (p.n_split <= 1 || x < p.n_main) {
    w.cta = x;
    w.gb0 = 0;
    w.gb1 = sh.n_box;
    w.n_sub = sh.n_sub;
    w.part_row = -1;
  } else {
    const int idx = x - p.n_main;
    const int tb = idx / p.n_split, j = idx - tb * p.n_split;
    w.cta = p.n_main + tb;
    w.gb0 = (int)((int64_t)sh.n_box * j / p.n_split);
    w.gb1 = (int)((int64_t)sh.n_box * (j + 1) / p.n_split);
    w.n_sub = range_sub_blocks(sh, w.gb0, w.gb1);
    w.part_row = (int64_t)idx * 256;
  }
  return w;
}

// dynamic shared memory (1024-byte aligned base)
constexpr int kOffK = 0;
constexpr int kOffV = kOffK + kKS * kTileBytes;
constexpr int kOffQ = kOffV + kKS * kTileBytes;          // exact path: two 128-row Q tiles
constexpr int kOffOnes = kOffQ + 2 * kTileBytes;         // 16 rows x 128 B of bf16 1.0 (B operand of the row-sum MMA)
constexpr int kOffTab = kOffOnes + 2048;                 // sub-block table
constexpr int kOffML = kOffTab + kMaxSub * 4;            // exact path: (m, l) exchange, float2 [4][128]
constexpr int kOffBarsFast = kOffML + 4 * 128 * 8;
constexpr int kOffBarsExact = kOffBarsFast + 32 * 8;
constexpr int kOffSlot = kOffBarsExact + 32 * 8;
constexpr int kAttnSmem = kOffSlot + 16 + 1024 /* alignment slack */;

// Sub-block table entry: box index (20 bits) | half of the box << 20 | last sub-block of its box << 21 | valid keys << 24.
// Shards are walked box by box (128 keys per TMA box, two 64-key sub-blocks); a shard whose key count is not a multiple
// of 128 ends with a one-sub-block box, so the (box, half) of sub-block i is not a function of i alone.
__device__ __forceinline__ void build_sub_table(const AttnShards& sh, uint32_t* tab, int lane, int gb0, int gb1) {
  int i0 = 0, box0 = 0;   // i0: entries written so far (sub-blocks inside [gb0, gb1)); box indices are relative to gb0
  for (int s = 0; s < sh.n; ++s) {
    const int nkv = sh.nkv[s];
    const int ns = (nkv + 63) >> 6, nb = (nkv + 127) >> 7;
    const int lo = max(gb0, box0), hi = min(gb1, box0 + nb);
    if (hi > lo) {
      const int l0 = 2 * (lo - box0), l1 = min(ns, 2 * (hi - box0));
      for (int l = l0 + lane; l < l1; l += 32) {
        const int valid = min(64, nkv - 64 * l);
        const uint32_t last = ((l & 1) || l == ns - 1) ? 1u : 0u;
        tab[i0 + l - l0] =
            uint32_t(box0 + (l >> 1) - gb0) | (uint32_t(l & 1) << 20) | (last << 21) | (uint32_t(valid) << 24);
      }
      i0 += l1 - l0;
    }
    box0 += nb;
  }
}
__device__ __forceinline__ int tab_box(uint32_t e) { return int(e & 0xFFFFFu); }
__device__ __forceinline__ int tab_half(uint32_t e) { return int((e >> 20) & 1u); }
__device__ __forceinline__ bool tab_last(uint32_t e) { return ((e >> 21) & 1u) != 0; }
__device__ __forceinline__ int tab_valid(uint32_t e) { return int(e >> 24); }

// TMA producer (one warp, warp-uniform, elect.sync leader): K and V boxes of every shard through kKS-deep mbarrier rings.
__device__ __forceinline__ void produce_kv(const AttnShards& sh, uint8_t* sK, uint8_t* sV, uint64_t* k_full,
                                           uint64_t* k_empty, uint64_t* v_full, uint64_t* v_empty, int bh, int gb0,
                                           int gb1) {
  const bool leader = elect_one();
  int g = 0;      // box counter of this CTA (ring stage and phase)
  int box0 = 0;   // global index of the shard's first box
  for (int s = 0; s < sh.n; ++s) {
    const int nb_s = (sh.nkv[s] + 127) >> 7;
    const int lo = max(gb0, box0), hi = min(gb1, box0 + nb_s);
    const int j0 = lo - box0, j1 = hi - box0;
    box0 += nb_s;
    if (j1 <= j0) continue;   // no box of this shard in the CTA's key range: its arrival is not waited for either
    if (sh.ready[s] != nullptr) {
      // the shard is being written by a peer's copy engine: wait for its arrival flag (written after the copy)
      if (leader) {
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys_u32(sh.ready[s]) - sh.ready_val[s]) < 0) {
          if (clock64() - t0 > sh.wait_cycles) {
            atomicOr(sh.status, 1u);
            break;
          }
          __nanosleep(100);
        }
        fence_proxy_async_all();
      }
      __syncwarp();
    }
    for (int j = j0; j < j1; ++j, ++g) {
      const int st = g % kKS;
      const uint32_t ph = (g / kKS) & 1;
      mbar_wait(&k_empty[st], ph ^ 1);
      if (leader) {
        mbar_expect_tx(&k_full[st], kTileBytes);
        tma_load_3d(sK + st * kTileBytes, &sh.k[s], &k_full[st], 0, j * 128, bh);
      }
      mbar_wait(&v_empty[st], ph ^ 1);
      if (leader) {
        mbar_expect_tx(&v_full[st], kTileBytes);
        tma_load_3d(sV + st * kTileBytes, &sh.v[s], &v_full[st], 0, j * 128, bh);
      }
    }
  }
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for a PAIR on the FMA / ALU pipes (Cody-Waite: floor by a round-down magic add, degree-3 minimax polynomial on
// the fraction, max rel. error 8.8e-5 — far below the bf16 rounding of P — exponent re-inserted by an integer add).
// Six packed f32x2 FMA-pipe ops per pair; offloads the 16-op/clk MUFU unit.  x is clamped to [-127, 128]: 2^-127 is a
// denormal (harmless), x >= 128 yields an exponent field of 255 (inf / NaN), which the overflow detection of the fast
// path catches exactly like MUFU's +inf.
__device__ __forceinline__ void ex2_poly_pair(uint64_t x2, float& p0, float& p1) {
  float x0, x1;
  unpack2(x2, x0, x1);
  x0 = fminf(fmaxf(x0, -127.0f), 128.0f);
  x1 = fminf(fmaxf(x1, -127.0f), 128.0f);
  const uint64_t xc = pack2(x0, x1);
  const uint64_t magic = pack2(12582912.0f, 12582912.0f);   // 1.5 * 2^23: low mantissa bits of the sum = floor(x)
  const uint64_t xr = add2_rm(xc, magic);
  const uint64_t f = sub2(xc, sub2(xr, magic));              // fraction in [0, 1)
  uint64_t p = fma2(f, pack2(0.077119089663028717f, 0.077119089663028717f),
                    pack2(0.227564394474029541f, 0.227564394474029541f));
  p = fma2(p, f, pack2(0.695146143436431885f, 0.695146143436431885f));
  p = fma2(p, f, pack2(1.0f, 1.0f));
  float q0, q1, r0, r1;
  unpack2(p, q0, q1);
  unpack2(xr, r0, r1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(r0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(r1) << 23));
}

#define LD_PROF(slot)                                  \
  if constexpr (PROF) {                                \
    const long long now = clock64();                   \
    prof_acc[slot] += now - tp;                        \
    tp = now;                                          \
  }

// ------------------------------------------------------------------------------------------------------------
// Exact path: FOUR independent online-softmax streams per CTA, 16 softmax warps (4 per SMSP).
//
// Measured on B200 (tools/softmax_mix_bench.cu): one warp alone runs the exponential mix at ~12 clk/element because
// its in-order issue cannot overlap MUFU with the dependent ops; three or more warps per SMSP reach the MUFU limit.
// TMEM (512 columns) cannot hold more than two 128-row query tiles with their accumulators, so each query tile is
// served by TWO streams that split the keys by 64-key sub-block parity: stream (t, b) owns sub-blocks i = b, b+2, ...
// of query tile t with its own score buffer S, its own output accumulator O and its own running (max, sum); the two
// streams of a tile are merged by log-sum-exp in the epilogue.
//   control warps: cw = TMA producer, cw + 1 = MMA issuer; softmax warps sw0 .. sw0 + 15: stream = (warp - sw0) / 4,
//   TMEM quadrant = warp % 4.  Standalone (variant 1): cw = 0, sw0 = 4; as the in-launch fallback of the fast path:
//   cw = 16, sw0 = 0 (the register split set up by the fast path — 104 for warps 0-15, 56 for 16-19 — already fits).
//   TMEM columns: S/P(stream) at 64*stream [0,256) ; O(stream) at 256 + 64*stream [256,512)
// All 640 threads must call this (it initialises its own barrier set and synchronises the CTA).
__device__ __forceinline__ void attn_exact_body(const AttnShards& sh, const AttnParams& p, uint8_t* smem,
                                                uint32_t tmem_base, const AttnWork& wk, int cw, int sw0) {
  uint8_t* sK = smem + kOffK;
  uint8_t* sV = smem + kOffV;
  uint8_t* sQ = smem + kOffQ;
  const uint32_t* tab = reinterpret_cast<const uint32_t*>(smem + kOffTab);
  float2* sML = reinterpret_cast<float2*>(smem + kOffML);   // [stream][128] (m * scale_log2, l)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBarsExact);
  uint64_t* q_full = bars;               // 1
  uint64_t* k_full = bars + 1;           // kKS
  uint64_t* k_empty = k_full + kKS;
  uint64_t* v_full = k_empty + kKS;
  uint64_t* v_empty = v_full + kKS;
  uint64_t* s_full = v_empty + kKS;      // [stream] = 4
  uint64_t* p_full = s_full + 4;         // [stream] = 4
  uint64_t* o_done = p_full + 4;         // [stream] = 4

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_blocks = (p.nq + 255) / 256;
  const int bh = wk.cta / q_blocks;
  const int q0 = (wk.cta % q_blocks) * 256;
  const int n_sub = wk.n_sub;

  if (warp == cw && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < kKS; ++s) {
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == cw) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_expect_tx(q_full, 2 * kTileBytes);
      tma_load_3d(sQ, &sh.q, q_full, 0, q0, bh);
      tma_load_3d(sQ + kTileBytes, &sh.q, q_full, 0, q0 + 128, bh);
    }
    __syncwarp();
    produce_kv(sh, sK, sV, k_full, k_empty, v_full, v_empty, bh, wk.gb0, wk.gb1);
  } else if (warp == cw + 1) {
    // ------------------------------------------------------------------ MMA issuer
    const bool leader = elect_one();
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 64);
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, false, true);  // B (=V) is N-major
    const uint64_t qdesc = make_sdesc_sw128(smem_u32(sQ));
    const uint64_t kdesc = make_sdesc_sw128(smem_u32(sK));
    const uint64_t vdesc = make_sdesc_sw128(smem_u32(sV));
    // descriptor address units are 16 B: ring stage = 1024, 64-row half = 512, query tile = 1024
    // S(t, i) = Q_t K_i^T into the score buffer of stream 2t + (i&1)
    auto issue_s = [&](int t, int i, uint32_t e) {
      const uint64_t adesc = qdesc + uint32_t(t * (kTileBytes >> 4));
      const uint64_t bdesc = kdesc + uint32_t((tab_box(e) % kKS) * (kTileBytes >> 4) + tab_half(e) * (kTileBytes >> 5));
      const int st = 2 * t + (i & 1);
      const uint32_t d = tmem_base + st * 64;
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(d, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0);
        umma_commit(&s_full[st]);
      }
    };
    // O(stream) += P(t,i) V_i ; P: 128 lanes x 64 keys bf16 = 32 TMEM columns over S(stream)
    auto issue_pv = [&](int t, int i, uint32_t e) {
      const uint64_t bdesc = vdesc + uint32_t((tab_box(e) % kKS) * (kTileBytes >> 4) + tab_half(e) * (kTileBytes >> 5));
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
    auto k_wait = [&](int box) { mbar_wait(&k_full[box % kKS], (box / kKS) & 1); };
    auto v_wait = [&](int box) { mbar_wait(&v_full[box % kKS], (box / kKS) & 1); };

    mbar_wait(q_full, 0);
    for (int i0 = 0; i0 < 2 && i0 < n_sub; ++i0) {
      const uint32_t e = tab[i0];
      if (tab_half(e) == 0) k_wait(tab_box(e));
      tc_fence_after();
      issue_s(0, i0, e);
      issue_s(1, i0, e);
      if (leader && tab_last(e)) umma_commit(&k_empty[tab_box(e) % kKS]);
    }
    for (int i = 0; i < n_sub; ++i) {
      const int b = i & 1;
      const uint32_t par = (i >> 1) & 1;
      const bool has_next = (i + 2) < n_sub;
      const uint32_t e = tab[i];
      const uint32_t en = has_next ? tab[i + 2] : 0u;
      if (tab_half(e) == 0) v_wait(tab_box(e));
      if (has_next && tab_half(en) == 0) k_wait(tab_box(en));
      // tile 0
      mbar_wait(&p_full[0 + b], par);
      tc_fence_after();
      issue_pv(0, i, e);
      if (has_next) issue_s(0, i + 2, en);
      // tile 1
      mbar_wait(&p_full[2 + b], par);
      tc_fence_after();
      issue_pv(1, i, e);
      if (leader && tab_last(e)) umma_commit(&v_empty[tab_box(e) % kKS]);
      if (has_next) {
        issue_s(1, i + 2, en);
        if (leader && tab_last(en)) umma_commit(&k_empty[tab_box(en) % kKS]);
      }
    }
  } else if (warp >= sw0 && warp < sw0 + 16) {
    // -------------------------------------------------------------------- softmax / correction / epilogue
    const int st = (warp - sw0) >> 2;     // stream
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
    int kk = 0;   // blocks done by this stream
    for (int i = b; i < n_sub; i += 2, ++kk) {
      mbar_wait(&s_full[st], kk & 1);
      tc_fence_after();
      uint32_t s[64];
      LD_TMEM_LD32(ts + 0, (s + 0));
      LD_TMEM_LD32(ts + 32, (s + 32));
      tmem_ld_wait();
      const int valid = tab_valid(tab[i]);
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
      float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        float pv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) pv[e] = ex2(fmaf(__uint_as_float(s[4 * c + e]), sl2, -msc));
        sum0 += pv[0];
        sum1 += pv[1];
        sum2 += pv[2];
        sum3 += pv[3];
        pk[2 * c] = pack_bf16x2(pv[0], pv[1]);
        pk[2 * c + 1] = pack_bf16x2(pv[2], pv[3]);
      }
      l += (sum0 + sum1) + (sum2 + sum3);
      LD_TMEM_ST32(ts, pk);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[st]);
    }

    // ---- epilogue: merge the two streams of this query tile by log-sum-exp; stream b writes output columns
    //      [32b, 32b+32) of the head
    if (kk > 0) {
      mbar_wait(&o_done[st], (kk - 1) & 1);
      tc_fence_after();
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
      const bool partial = wk.part_row >= 0;
      const int64_t prow = wk.part_row + t * 128 + row_in_tile;
      bf16* orow = p.out + ((int64_t)bb * p.nq + q_row) * (p.heads * 64) + h * 64 + c0;
      float* of_base = partial ? p.part_o + prow * 64 : (p.out_f32 ? p.out_f32 + ((int64_t)bh * p.nq + q_row) * 64 : nullptr);
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
        if (!partial) {
          uint4 v;
          v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
          v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
          *reinterpret_cast<uint4*>(orow + g * 8) = v;
        }
        if (of_base != nullptr) {
          float* of = of_base + c0 + g * 8;
          *reinterpret_cast<float4*>(of) = make_float4(f[0], f[1], f[2], f[3]);
          *reinterpret_cast<float4*>(of + 4) = make_float4(f[4], f[5], f[6], f[7]);
        }
      }
      if (b == 0) {
        if (partial) p.part_lse[prow] = m_all + log2f(l_all);
        else if (p.lse != nullptr) p.lse[(int64_t)bh * p.nq + q_row] = m_all + log2f(l_all);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Fast path (round 2, "attn5").
//
// Two 128-row query tiles per CTA, FOUR softmax streams: stream (t, b) exponentiates the 64-key sub-blocks i = b, b+2, ...
// of query tile t; each of its 4 warps owns a 32-row TMEM quadrant and all 64 columns of the sub-block.  Every SMSP
// thus holds one warp of each stream — four INDEPENDENT instruction streams, so while one waits for the tensor pipe
// (PV then the next S) the other three keep the exponential units busy (tools/softmax_mix_bench.cu: >= 3 warps per
// SMSP are needed to reach the MUFU limit), and a warp pays one barrier wait + one barrier arrive per 64 columns.
//
// What bounds the kernel is the exponential: 16 384 of them per 128x128 score tile at 16 / clk / SM on the MUFU unit
// is 1024 clk against 512 clk of tensor work.  Round-2 changes, all aimed at that (profiles/r2_softmax_mix_bench.txt:
// 8.1 -> 6.0 clk per warp-element with 4 warps per SMSP):
//   * the row sum is no longer an FADD per element: the tensor core accumulates it, L_t += P(t,i) x ones (an N = 16 MMA
//     against a constant tile, 12 clk), from exactly the bf16 P values the PV product uses;
//   * the scale-subtract is a packed FFMA2 per column pair;
//   * KP of every 16 column pairs are exponentiated by a packed f32x2 polynomial on the FMA pipe (ex2_poly_pair), the
//     rest by MUFU.EX2;
//   * no running maximum, hence no rescaling, hence the two streams of a tile accumulate into ONE output accumulator
//     O_t and ONE row-sum accumulator L_t (a sum is a sum), and there is nothing to merge in the epilogue.
//
// Reference maximum: floating point is scale-invariant, so the online-softmax reference only has to prevent overflow,
// not track the running maximum.  Each row takes the maximum of its FIRST sub-block as the reference for the whole row
// (stream (t,0) computes it and hands it to stream (t,1) through shared memory) and never rescales: later scores may
// exceed the reference by up to 2^127 before exp2 overflows, and terms far below it flush to zero exactly as their
// true weight demands.  Overflow (a score more than ~127 log2-units above the first block's maximum — never seen on
// LayerNormed q/k, but constructible) makes the row sum or the output non-finite; the CTA then re-runs its 256 rows
// through the exact path above in the same launch.  tests/test_kernels_gpu.py::test_attention_overflow_fixup.
//
// Q is stored once into TMEM (bf16 pairs, one row per lane) by the softmax threads, so S = Q K^T runs as a TS-mode MMA
// whose only shared-memory operand is the K sub-block: 32 clk per 128x64x16 instead of 48 (tools/mma_bench.cu).  P is
// written in place over the first 32 columns of the stream's score buffer (only the writing warp ever reads those).
//   warps 0-15: softmax (stream w>>2 = 2t+b, TMEM quadrant w&3)   16: builds the sub-block table   17: TMA producer
//   18, 19: MMA issuers of query tile 0 and 1 (independent instruction streams; K/V ring stages are released by one
//   commit from each).
//   TMEM columns: S/P(t,b) at 64(2t+b) [0,256) ; O_t at 256+64t [256,384) ; Q_t at 384+32t [384,448) ;
//                 L_t (row sums, 16 identical columns) at 448+16t [448,480)
template <int KP, bool TRUNC, bool PROF>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn5_kernel(const __grid_constant__ AttnShards sh, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem + kOffK;
  uint8_t* sV = smem + kOffV;
  uint32_t* sOnes = reinterpret_cast<uint32_t*>(smem + kOffOnes);
  uint32_t* tab = reinterpret_cast<uint32_t*>(smem + kOffTab);
  float* sRef = reinterpret_cast<float*>(smem + kOffML);   // [tile][128] reference maximum * scale_log2
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBarsFast);
  uint64_t* q_ready = bars;              // [tile] = 2
  uint64_t* k_full = bars + 2;           // kKS
  uint64_t* k_empty = k_full + kKS;
  uint64_t* v_full = k_empty + kKS;
  uint64_t* v_empty = v_full + kKS;
  uint64_t* s_full = v_empty + kKS;      // [stream] = 4   S(t,i) is in TMEM
  uint64_t* p_full = s_full + 4;         // [stream] = 4   P(t,i) is in TMEM
  uint64_t* all_done = p_full + 4;       // 1: every MMA of this CTA has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffSlot);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_blocks = (p.nq + 255) / 256;
  const AttnWork wk = attn_work(sh, p);
  const int bh = wk.cta / q_blocks;
  const int q0 = (wk.cta % q_blocks) * 256;
  const int n_sub = wk.n_sub;

  if (warp == 17 && lane == 0) {
    tma_prefetch_desc(&sh.q);
    for (int s = 0; s < sh.n; ++s) {
      tma_prefetch_desc(&sh.k[s]);
      tma_prefetch_desc(&sh.v[s]);
    }
    for (int s = 0; s < kKS; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 2);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 2);
    }
    for (int i = 0; i < 2; ++i) mbar_init(&q_ready[i], 128);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
    }
    mbar_init(all_done, 2);
    fence_barrier_init();
  }
  if (warp == 16) {
    build_sub_table(sh, tab, lane, wk.gb0, wk.gb1);
    for (int i = lane; i < 512; i += 32) sOnes[i] = 0x3F803F80u;   // bf16 1.0 pairs
    fence_proxy_async_smem();   // the ones tile is read by the async proxy (tcgen05.mma B operand)
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
  if (!p.exact_only) {
    if (warp == 17) {
      produce_kv(sh, sK, sV, k_full, k_empty, v_full, v_empty, bh, wk.gb0, wk.gb1);
    } else if (warp >= 18) {
      // ------------------------------------------------------------------ MMA issuer of query tile t
      const int t = warp - 18;
      const bool leader = elect_one();
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 64);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, false, true);  // B (=V) is N-major
      constexpr uint32_t idesc_l = make_idesc_bf16(128, 16);               // row sums: P x ones[64 keys x 16]
      const uint64_t kdesc = make_sdesc_sw128(smem_u32(sK));
      const uint64_t vdesc = make_sdesc_sw128(smem_u32(sV));
      const uint64_t odesc = make_sdesc_sw128(smem_u32(sOnes));
      const uint32_t tm_s = tmem_base + 128 * t;        // S/P(t,0); S/P(t,1) 64 columns further
      const uint32_t tm_o = tmem_base + 256 + 64 * t;
      const uint32_t tm_q = tmem_base + 384 + 32 * t;
      const uint32_t tm_l = tmem_base + 448 + 16 * t;
      long long w_k = 0, w_v = 0, w_p = 0, i_s = 0, i_pv = 0, t_all = 0;
      const uint32_t tab_s = smem_u32(tab);
      auto tab_at = [&](int i) {   // explicit ld.shared (the generic pointer would compile to a generic LD)
        uint32_t e;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(tab_s + 4u * uint32_t(i)));
        return e;
      };
      // descriptor address units are 16 B: ring stage = 1024, 64-row half = 512
      // S(t, i) = Q_t K_i^T into the buffer of stream (t, i&1); Q_t from TMEM (8 columns per 16-dim K step)
      auto k_wait = [&](uint32_t e) {
        if (tab_half(e) == 0) mbar_wait(&k_full[tab_box(e) % kKS], (tab_box(e) / kKS) & 1);
      };
      auto v_wait = [&](uint32_t e) {
        if (tab_half(e) == 0) mbar_wait(&v_full[tab_box(e) % kKS], (tab_box(e) / kKS) & 1);
      };
      auto issue_s = [&](int i, uint32_t e) {
        const int st = tab_box(e) % kKS;
        const uint64_t bdesc = kdesc + uint32_t(st * (kTileBytes >> 4) + tab_half(e) * (kTileBytes >> 5));
        const uint32_t d = tm_s + (i & 1) * 64;
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ts(d, tm_q + 8 * k, bdesc + 2 * k, idesc_s, k != 0);
          umma_commit(&s_full[2 * t + (i & 1)]);
          if (tab_last(e)) umma_commit(&k_empty[st]);
        }
      };
      // O_t += P(t,i) V_i and L_t += P(t,i) 1 ; P: 128 lanes x 64 keys bf16 = the first 32 columns of the stream's score
      // buffer; 16 keys (8 columns) per MMA
      auto issue_pv = [&](int i, uint32_t e) {
        const int st = tab_box(e) % kKS;
        const uint64_t bdesc = vdesc + uint32_t(st * (kTileBytes >> 4) + tab_half(e) * (kTileBytes >> 5));
        const uint32_t a = tm_s + (i & 1) * 64;
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // 16 keys per step: 16 rows x 128 B = 2048 B (encoded 128)
            umma_ts(tm_o, a + 8 * k, bdesc + 128 * k, idesc_o, (i | k) != 0);
            umma_ts(tm_l, a + 8 * k, odesc, idesc_l, (i | k) != 0);
          }
          if (tab_last(e)) umma_commit(&v_empty[st]);
        }
      };
      mbar_wait(&q_ready[t], 0);
      if constexpr (PROF) t_all = clock64();
      uint32_t e_pv = tab_at(0);                          // sub-block i (PV side)
      uint32_t e_s = n_sub > 1 ? tab_at(1) : 0u;          // sub-block i + 1, then i + 2 (S side)
      k_wait(e_pv);
      tc_fence_after();
      issue_s(0, e_pv);
      if (n_sub > 1) {
        k_wait(e_s);
        tc_fence_after();
        issue_s(1, e_s);
      }
      for (int i = 0; i < n_sub; ++i) {
        const bool has_next = i + 2 < n_sub;
        const uint32_t e_n = has_next ? tab_at(i + 2) : 0u;
        // The K / V stages needed by this iteration landed long ago (the rings run four boxes ahead), but even a
        // completed mbarrier wait costs ~100 clk of latency in this serial chain: take those waits BEFORE the wait for
        // the softmax (P(t,i)), where they overlap with time spent waiting anyway.
        long long c0 = 0;
        if constexpr (PROF) c0 = clock64();
        v_wait(e_pv);
        if constexpr (PROF) { const long long c1 = clock64(); w_v += c1 - c0; c0 = c1; }
        if (has_next) k_wait(e_n);
        if constexpr (PROF) { const long long c1 = clock64(); w_k += c1 - c0; c0 = c1; }
        mbar_wait(&p_full[2 * t + (i & 1)], (i >> 1) & 1);
        if constexpr (PROF) { const long long c1 = clock64(); w_p += c1 - c0; c0 = c1; }
        tc_fence_after();
        issue_pv(i, e_pv);
        if constexpr (PROF) { __syncwarp(); const long long c1 = clock64(); i_pv += c1 - c0; c0 = c1; }
        if (has_next) issue_s(i + 2, e_n);   // same buffer as P(t,i): the tensor pipe executes in issue order
        if constexpr (PROF) { __syncwarp(); i_s += clock64() - c0; }
        e_pv = e_s;
        e_s = e_n;
      }
      if (leader) umma_commit(all_done);
      if constexpr (PROF) {
        if (leader && p.prof != nullptr) {
          long long* d = p.prof + ((int64_t)blockIdx.x * 20 + warp) * 8;
          d[0] = w_k; d[1] = w_v; d[2] = w_p; d[3] = 0; d[4] = i_s; d[5] = i_pv; d[6] = clock64() - t_all;
        }
      }
    } else if (warp < 16) {
      reg_alloc<104>();   // 16 warps x 32 x 8 = 4096 <= the 5120 registers released by the control warpgroup
      // -------------------------------------------------------------------- softmax / epilogue
      const int st = warp >> 2;                   // stream
      const int t = st >> 1;                      // query tile
      const int b = st & 1;                       // sub-block parity served by this stream
      const int quad = warp & 3;
      const int row_in_tile = quad * 32 + lane;
      const int q_row = q0 + t * 128 + row_in_tile;
      const uint32_t lane_base = uint32_t(quad * 32) << 16;
      const uint32_t ts = tmem_base + lane_base + st * 64;         // S / P of this stream
      const uint32_t to = tmem_base + lane_base + 256 + t * 64;    // O_t
      const float sl2 = p.scale_log2;

      if (b == 0) {
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
      long long prof_acc[6] = {0, 0, 0, 0, 0, 0};
      long long tp = 0;
      if constexpr (PROF) tp = clock64();
      int kk = 0;
      for (int i = b; i < n_sub; i += 2, ++kk) {
        mbar_wait(&s_full[st], kk & 1);
        tc_fence_after();
        LD_PROF(0);
        uint32_t s[64];
        LD_TMEM_LD32(ts, s);
        LD_TMEM_LD32(ts + 32, (s + 32));
        tmem_ld_wait();
        LD_PROF(1);
        const int valid = tab_valid(tab[i]);
        if (valid < 64) {
#pragma unroll
          for (int c = 0; c < 64; ++c)
            if (c >= valid) s[c] = 0xff800000u;  // -inf
        }
        if (kk == 0) {
          // one reference per ROW for both streams of the tile (they share O_t and L_t): the maximum of sub-block 0
          if (b == 0) {
            float mx4[4] = {__uint_as_float(s[0]), __uint_as_float(s[1]), __uint_as_float(s[2]), __uint_as_float(s[3])};
#pragma unroll
            for (int c = 4; c < 64; c += 4) {
              mx4[0] = fmaxf(mx4[0], __uint_as_float(s[c]));
              mx4[1] = fmaxf(mx4[1], __uint_as_float(s[c + 1]));
              mx4[2] = fmaxf(mx4[2], __uint_as_float(s[c + 2]));
              mx4[3] = fmaxf(mx4[3], __uint_as_float(s[c + 3]));
            }
            msc = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * sl2;
            sRef[t * 128 + row_in_tile] = msc;
          }
          named_bar_sync(1 + t * 4 + quad, 64);   // the two warps (b = 0, 1) that own these 32 rows
          if (b == 1) msc = sRef[t * 128 + row_in_tile];
        }
        LD_PROF(2);
        uint32_t pk[32];
        const uint64_t sc2 = pack2(sl2, sl2), nm2 = pack2(-msc, -msc);
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const uint64_t x2 = fma2(pack2u(s[2 * c], s[2 * c + 1]), sc2, nm2);
          float p0, p1;
          if (KP > 0 && ((c * KP) % 16) < KP) {   // KP of 16 pairs, evenly spread, on the FMA pipe
            ex2_poly_pair(x2, p0, p1);
          } else {
            float x0, x1;
            unpack2(x2, x0, x1);
            p0 = ex2(x0);
            p1 = ex2(x1);
          }
          if constexpr (TRUNC) {
            // bf16 by truncation (one byte permute instead of the half-rate F2FP): the bias cancels because the row
            // sum is accumulated from the same truncated values
            asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(pk[c]) : "r"(__float_as_uint(p0)), "r"(__float_as_uint(p1)));
          } else {
            pk[c] = pack_bf16x2(p0, p1);
          }
        }
        LD_PROF(3);
        LD_TMEM_ST32(ts, pk);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[st]);
        LD_PROF(4);
      }
      if (n_sub == 1 && b == 1) named_bar_sync(1 + t * 4 + quad, 64);   // pair barrier of a stream without any sub-block

      // ---- epilogue: the row sum comes from the tensor core (L_t); stream b writes output columns [32b, 32b+32)
      mbar_wait(all_done, 0);
      tc_fence_after();
      if constexpr (PROF) {
        if (lane == 0 && p.prof != nullptr) {
          long long* d = p.prof + ((int64_t)blockIdx.x * 20 + warp) * 8;
          for (int e = 0; e < 6; ++e) d[e] = prof_acc[e];
        }
      }
      if (b == 1) msc = sRef[t * 128 + row_in_tile];
      uint32_t lbits;
      LD_TMEM_LD1(tmem_base + lane_base + 448 + 16 * t, lbits);
      const int c0 = 32 * b;
      uint32_t o[32];
      LD_TMEM_LD32(to + c0, o);
      tmem_ld_wait();
      const float l_all = __uint_as_float(lbits);
      const float inv_l = 1.0f / l_all;
      const bool valid_row = q_row < p.nq;
      const int bb = bh / p.heads, hd = bh - bb * p.heads;
      if (valid_row) {
        bad = !(l_all < INFINITY) || !(l_all > 0.f);   // inf / NaN row sum (or everything flushed to zero)
        const bool partial = wk.part_row >= 0;
        const int64_t prow = wk.part_row + t * 128 + row_in_tile;
        bf16* orow = p.out + ((int64_t)bb * p.nq + q_row) * (p.heads * 64) + hd * 64 + c0;
        float* of_base =
            partial ? p.part_o + prow * 64 : (p.out_f32 ? p.out_f32 + ((int64_t)bh * p.nq + q_row) * 64 : nullptr);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            f[e] = __uint_as_float(o[g * 8 + e]) * inv_l;
            bad |= !(fabsf(f[e]) < INFINITY);
          }
          if (!partial) {
            uint4 v;
            v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
            v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
            *reinterpret_cast<uint4*>(orow + g * 8) = v;
          }
          if (of_base != nullptr) {
            float* of = of_base + c0 + g * 8;
            *reinterpret_cast<float4*>(of) = make_float4(f[0], f[1], f[2], f[3]);
            *reinterpret_cast<float4*>(of + 4) = make_float4(f[4], f[5], f[6], f[7]);
          }
        }
        // truncated P values are low by 2^-9 on average (uniform mantissa tails): the normalisation above uses the same
        // values and needs no correction, the exported log-sum-exp does
        if (b == 0) {
          const float lse_v = msc + log2f(TRUNC ? l_all * 1.001953125f : l_all);
          if (partial) p.part_lse[prow] = lse_v;
          else if (p.lse != nullptr) p.lse[(int64_t)bh * p.nq + q_row] = lse_v;
        }
      }
    }
  } else {
    if (warp < 16) reg_alloc<104>();
  }

  tc_fence_before();
  const int redo = __syncthreads_or((bad || p.exact_only) ? 1 : 0);
  if (redo) {
    // the fixed reference maximum overflowed somewhere in this CTA's 256 rows (or variant 1): exact path, same launch
    tc_fence_after();
    attn_exact_body(sh, p, smem, tmem_base, wk, /*cw=*/16, /*sw0=*/0);
    tc_fence_before();
    __syncthreads();
  }
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

// Tail split: out rows of tail block tb = sum over its n_split partial results weighted by 2^(lse_j - max lse)
__global__ void __launch_bounds__(256) attn_split_merge_kernel(const float* __restrict__ part_o,
                                                               const float* __restrict__ part_lse, bf16* __restrict__ out,
                                                               int n_main, int n_split, int n_tail, int heads, int nq) {
  // 16 threads per row (4 floats each), 256 rows per tail block
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = gid >> 4;
  const int c = (int)(gid & 15) * 4;
  if (row >= (int64_t)n_tail * 256) return;
  const int tb = (int)(row >> 8), r = (int)(row & 255);
  const int q_blocks = (nq + 255) / 256;
  const int work = n_main + tb;
  const int bh = work / q_blocks;
  const int q_row = (work % q_blocks) * 256 + r;
  if (q_row >= nq) return;
  const int64_t base = ((int64_t)tb * n_split) * 256 + r;   // partial slot j: base + j * 256
  float m = -INFINITY;
  for (int j = 0; j < n_split; ++j) m = fmaxf(m, part_lse[base + (int64_t)j * 256]);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float wsum = 0.f;
  for (int j = 0; j < n_split; ++j) {
    const float w = exp2f(part_lse[base + (int64_t)j * 256] - m);
    const float4 o = *reinterpret_cast<const float4*>(part_o + (base + (int64_t)j * 256) * 64 + c);
    acc.x = fmaf(o.x, w, acc.x); acc.y = fmaf(o.y, w, acc.y); acc.z = fmaf(o.z, w, acc.z); acc.w = fmaf(o.w, w, acc.w);
    wsum += w;
  }
  const float inv = 1.0f / wsum;
  const int bb = bh / heads, hd = bh - bb * heads;
  bf16* o = out + ((int64_t)bb * nq + q_row) * (heads * 64) + hd * 64 + c;
  uint2 v;
  v.x = pack_bf16x2(acc.x * inv, acc.y * inv);
  v.y = pack_bf16x2(acc.z * inv, acc.w * inv);
  *reinterpret_cast<uint2*>(o) = v;
}

template <int KP, bool TRUNC, bool PROF = false>
static int launch_attn5(const AttnShards& sh, const AttnParams& prm, int grid, cudaStream_t st) {
  auto kern = attn5_kernel<KP, TRUNC, PROF>;
  // per template instantiation and device (cudaFuncSetAttribute is per device)
  static bool attr_set[64] = {false};
  int dev = 0;
  LD_CHECK_CUDA(cudaGetDevice(&dev));
  LD_CHECK_ARG(dev >= 0 && dev < 64, "ld_attention: device index out of range");
  if (!attr_set[dev]) {
    LD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    attr_set[dev] = true;
  }
  kern<<<grid, kAttnThreads, kAttnSmem, st>>>(sh, prm);
  LD_CHECK_CUDA(cudaGetLastError());
  if (prm.n_split > 1) {
    const int n_tail = (grid - prm.n_main) / prm.n_split;
    const int64_t threads = (int64_t)n_tail * 256 * 16;
    attn_split_merge_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(prm.part_o, prm.part_lse, prm.out, prm.n_main,
                                                                               prm.n_split, n_tail, prm.heads, prm.nq);
    LD_CHECK_CUDA(cudaGetLastError());
  }
  return LD_OK;
}

}  // namespace ld

using namespace ld;

static long long* g_attn_prof = nullptr;
// Debug hook (tools/attn_phase_prof.py): per-(CTA, warp) phase cycle counters for the next attention calls.
extern "C" void ld_debug_attn_prof(long long* buf) { g_attn_prof = buf; }

// per-device status word of the shard waits (bit 0: a wait timed out), allocated on first use
static uint32_t* g_status[64] = {nullptr};
static int status_word(uint32_t** out) {
  int dev = 0;
  LD_CHECK_CUDA(cudaGetDevice(&dev));
  LD_CHECK_ARG(dev >= 0 && dev < 64, "ld_attention: device index out of range");
  if (g_status[dev] == nullptr) {
    LD_CHECK_CUDA(cudaMalloc(&g_status[dev], 256));
    LD_CHECK_CUDA(cudaMemset(g_status[dev], 0, 256));
  }
  *out = g_status[dev];
  return LD_OK;
}

constexpr int kMaxSplitCtas = 320;   // bound on the CTAs of the split tail (workspace <= 21 MB)

// Tail split plan.  A grid of `grid` equal CTAs on `sms` SMs runs ceil(grid / sms) waves; the last one is partly empty.  Its
// n_tail query blocks are each split over S key ranges so that the tail costs ceil(n_tail * S / sms) / S (+ a fixed cost per
// CTA: prologue, first-block reference, epilogue) instead of one full wave.  Returns S (1 = no split).
static int plan_tail_split(int grid, int sms, int n_box, int* n_main) {
  static const int enabled = [] { const char* e = getenv("LD_ATTN_SPLIT"); return (e && e[0] == '0') ? 0 : 1; }();
  *n_main = grid;
  const int n_tail = grid % sms;
  if (!enabled || n_tail == 0) return 1;
  int best = 1;
  double best_cost = 0.97;   // a split must save at least 3 % of a wave to be worth the merge
  for (int S = 2; S <= 16; ++S) {
    if ((int64_t)n_tail * S > kMaxSplitCtas || n_box < 6 * S) break;
    const double per_cta = 1.0 / S + 0.025 + 1.0 / n_box;   // share of the keys + fixed cost (in units of a full CTA)
    const double cost = (double)((n_tail * S + sms - 1) / sms) * per_cta;
    if (cost < best_cost) { best_cost = cost; best = S; }
  }
  if (best > 1) *n_main = grid - n_tail;
  return best;
}

extern "C" int ld_attention_status(unsigned int* host_out, int reset) {
  uint32_t* w = nullptr;
  int rc = status_word(&w);
  if (rc != LD_OK) return rc;
  LD_CHECK_ARG(host_out != nullptr, "ld_attention_status: null pointer");
  LD_CHECK_CUDA(cudaMemcpy(host_out, w, 4, cudaMemcpyDeviceToHost));
  if (reset) LD_CHECK_CUDA(cudaMemset(w, 0, 4));
  return LD_OK;
}

static int total_boxes(const ld_kv_shard* shards, int n_shards) {
  int nb = 0;
  for (int s = 0; s < n_shards; ++s) nb += (shards[s].nkv + 127) / 128;
  return nb;
}

extern "C" size_t ld_attention_workspace_bytes(const ld_kv_shard* shards, int n_shards, int batch, int heads, int nq) {
  if (shards == nullptr || n_shards < 1 || n_shards > kMaxShards || batch <= 0 || heads <= 0 || nq <= 0) return 0;
  if (check_device() != LD_OK) return 0;
  const int grid = batch * heads * ((nq + 255) / 256);
  int n_main = grid;
  const int S = plan_tail_split(grid, sm_count(), total_boxes(shards, n_shards), &n_main);
  return S > 1 ? (size_t)(grid - n_main) * S * 256 * 65 * sizeof(float) : 0;
}

extern "C" int ld_attention_shards_bf16(const void* q, const ld_kv_shard* shards, int n_shards, void* out, float* lse,
                                        float* out_f32, int batch, int heads, int nq, int q_rows, int variant,
                                        void* stream) {
  return ld_attention_shards_ws_bf16(q, shards, n_shards, out, lse, out_f32, batch, heads, nq, q_rows, variant, nullptr, 0,
                                     stream);
}

extern "C" int ld_attention_shards_ws_bf16(const void* q, const ld_kv_shard* shards, int n_shards, void* out, float* lse,
                                           float* out_f32, int batch, int heads, int nq, int q_rows, int variant,
                                           void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_device();
  if (rc != LD_OK) return rc;
  LD_CHECK_ARG(q && shards && out, "ld_attention: null pointer");
  LD_CHECK_ARG(n_shards >= 1 && n_shards <= kMaxShards, "ld_attention: 1..%d K/V shards, got %d", kMaxShards, n_shards);
  LD_CHECK_ARG(batch > 0 && heads > 0 && nq > 0, "ld_attention: empty problem");
  LD_CHECK_ARG(nq <= q_rows, "ld_attention: nq exceeds the q buffer rows");
  LD_CHECK_ARG(out_f32 == nullptr || lse != nullptr, "ld_attention: out_f32 requires lse");
  const int BH = batch * heads;
  AttnShards sh;
  memset(&sh, 0, sizeof(sh));
  const uint32_t box[3] = {64, 128, 1};
  {
    // dims limited to the rows actually used so TMA zero-fills the tail
    const uint64_t dims[3] = {64, (uint64_t)nq, (uint64_t)BH};
    const uint64_t str[2] = {128, (uint64_t)q_rows * 128};
    rc = make_tmap_bf16(&sh.q, q, 3, dims, str, box);
    if (rc != LD_OK) return rc;
  }
  int n_sub = 0;
  for (int s = 0; s < n_shards; ++s) {
    const ld_kv_shard& d = shards[s];
    LD_CHECK_ARG(d.k && d.v && d.nkv > 0 && d.nkv <= d.kv_rows, "ld_attention: bad K/V shard %d", s);
    const uint64_t dims[3] = {64, (uint64_t)d.nkv, (uint64_t)BH};
    const uint64_t str[2] = {128, (uint64_t)d.kv_rows * 128};
    rc = make_tmap_bf16(&sh.k[s], d.k, 3, dims, str, box);
    if (rc != LD_OK) return rc;
    rc = make_tmap_bf16(&sh.v[s], d.v, 3, dims, str, box);
    if (rc != LD_OK) return rc;
    sh.nkv[s] = d.nkv;
    sh.ready[s] = d.ready_flag;
    sh.ready_val[s] = d.ready_value;
    n_sub += (d.nkv + 63) / 64;
    sh.n_box += (d.nkv + 127) / 128;
  }
  LD_CHECK_ARG(n_sub <= kMaxSub, "ld_attention: %d keys-blocks exceed the limit of %d (131072 keys)", n_sub, kMaxSub);
  sh.n = n_shards;
  sh.n_sub = n_sub;
  rc = status_word(&sh.status);
  if (rc != LD_OK) return rc;
  // a peer that never delivers must not hang the GPU: ~2 s at 1.9 GHz by default (LD_ATTN_WAIT_MS overrides; the
  // single-GPU multi-process tests, where the ranks time-slice one device, use a longer bound)
  static long long wait_cycles = 0;
  if (wait_cycles == 0) {
    const char* e = getenv("LD_ATTN_WAIT_MS");
    const long long ms = e ? atoll(e) : 2000;
    wait_cycles = (ms > 0 ? ms : 2000) * 2000000LL;
  }
  sh.wait_cycles = wait_cycles;
  AttnParams prm;
  prm.out = (bf16*)out;
  prm.lse = lse;
  prm.out_f32 = out_f32;
  prm.heads = heads;
  prm.nq = nq;
  prm.prof = nullptr;
  prm.exact_only = 0;
  prm.q = (const bf16*)q;
  prm.q_rows = q_rows;
  prm.scale_log2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
  int grid = BH * ((nq + 255) / 256);
  prm.n_main = grid;
  prm.n_split = 1;
  prm.part_o = prm.part_lse = nullptr;
  if (variant != 1 && variant != 6 && lse == nullptr && out_f32 == nullptr) {
    // callers that want the log-sum-exp (ring merges) get one CTA per query block; everyone else the tail split
    int n_main = grid;
    const int S = plan_tail_split(grid, sm_count(), sh.n_box, &n_main);
    const size_t need = (size_t)(grid - n_main) * S * 256 * 65 * sizeof(float);
    if (S > 1 && workspace != nullptr && workspace_bytes >= need) {   // no workspace: one CTA per query block
      LD_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "ld_attention: workspace must be 16-byte aligned");
      prm.part_o = reinterpret_cast<float*>(workspace);
      prm.part_lse = prm.part_o + (size_t)(grid - n_main) * S * 256 * 64;
      prm.n_main = n_main;
      prm.n_split = S;
      grid = n_main + (grid - n_main) * S;
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  // variant 0: fast path (4 of 16 column pairs on the FMA-pipe polynomial) with the exact path as in-launch fallback
  //         1: exact path only (per-block maxima, lazy rescaling)
  //         2 / 3 / 4: fast path with 0 / 5 / 6 of 16 pairs on the polynomial   5: variant 0 with truncating bf16 pack
  //         6: variant 0 with one CTA per query block even in the last wave (no tail split)
  // share of exponential pairs on the FMA-pipe polynomial for the default variants (0, 6): LD_ATTN_KP = 4 | 5 | 6 (A/B switch
  // between the shipped default, the boost-clock optimum and one step beyond; the sweep over 2..6 is in profiles/)
  static const int kp_default = [] { const char* e = getenv("LD_ATTN_KP"); const int v = e ? atoi(e) : 0; return (v >= 4 && v <= 6) ? v : kDefaultKP; }();
  switch (variant) {
    case 0:
    case 6:   // 6: variant 0 with one CTA per query block even in the last wave (no tail split)
      if (g_attn_prof != nullptr && variant == 0) {
        prm.prof = g_attn_prof;
        return launch_attn5<kDefaultKP, false, true>(sh, prm, grid, st);
      }
      switch (kp_default) {
        case 6: return launch_attn5<6, false>(sh, prm, grid, st);
        case 5: return launch_attn5<5, false>(sh, prm, grid, st);
        default: return launch_attn5<4, false>(sh, prm, grid, st);
      }
    case 1:
      prm.exact_only = 1;
      return launch_attn5<5, false>(sh, prm, grid, st);
    case 2: return launch_attn5<0, false>(sh, prm, grid, st);
    case 3: return launch_attn5<5, false>(sh, prm, grid, st);
    case 4: return launch_attn5<6, false>(sh, prm, grid, st);
    case 5: return launch_attn5<kDefaultKP, true>(sh, prm, grid, st);
    default:
      set_error("ld_attention: unknown variant %d", variant);
      return LD_ERR_ARG;
  }
}

extern "C" int ld_attention_bf16(const void* q, const void* k, const void* v, void* out, float* lse, float* out_f32,
                                 int batch, int heads, int nq, int q_rows, int nkv, int kv_rows, int variant,
                                 void* stream) {
  LD_CHECK_ARG(q && k && v && out, "ld_attention_bf16: null pointer");
  LD_CHECK_ARG(nkv > 0 && nkv <= kv_rows, "ld_attention_bf16: nkv exceeds the buffer rows");
  ld_kv_shard one;
  one.k = k;
  one.v = v;
  one.nkv = nkv;
  one.kv_rows = kv_rows;
  one.ready_flag = nullptr;
  one.ready_value = 0;
  return ld_attention_shards_bf16(q, &one, 1, out, lse, out_f32, batch, heads, nq, q_rows, variant, stream);
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
