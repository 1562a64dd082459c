// Throughput of the softmax exponential phase instruction mix on sm_100a, 1 or 2 warps per SMSP:
//   per element: FFMA (scale - max), MUFU.EX2, FADD (row sum); per pair: F2FP bf16x2 pack.
// MODE 0: plain mix   1: packed f32x2 FFMA/FADD   2: no pack   3: no FADD   4: MUFU only   5: every 4th exp polynomial
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -127.0f);
  float xr;
  asm("add.rm.ftz.f32 %0, %1, %2;" : "=f"(xr) : "f"(x), "f"(12582912.0f));
  const float f = x - (xr - 12582912.0f);
  float p = fmaf(f, 0.077119089663028717f, 0.227564394474029541f);
  p = fmaf(p, f, 0.695146143436431885f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xr) << 23));
}

template <int MODE>
__global__ void k(const float* in, uint32_t* out, unsigned long long* cyc, float sl2, float msc, int iters) {
  float s[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) s[i] = in[threadIdx.x * 64 + i];
  uint32_t acc = 0;
  float sum0 = 0, sum1 = 0, sum2 = 0, sum3 = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t pk[32];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      float pv[4];
      if (MODE == 1) {
        unsigned long long a01, a23, sc, ms;
        asm("mov.b64 %0, {%1, %2};" : "=l"(a01) : "f"(s[4 * c]), "f"(s[4 * c + 1]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(a23) : "f"(s[4 * c + 2]), "f"(s[4 * c + 3]));
        asm("mov.b64 %0, {%1, %1};" : "=l"(sc) : "f"(sl2));
        asm("mov.b64 %0, {%1, %1};" : "=l"(ms) : "f"(-msc));
        asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a01) : "l"(sc), "l"(ms));
        asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a23) : "l"(sc), "l"(ms));
        float x0, x1, x2, x3;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(a01));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x2), "=f"(x3) : "l"(a23));
        pv[0] = ex2(x0); pv[1] = ex2(x1); pv[2] = ex2(x2); pv[3] = ex2(x3);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float x = fmaf(s[4 * c + e], sl2, -msc);
          if (MODE == 5 && e == 3) pv[e] = ex2_poly(x); else pv[e] = ex2(x);
        }
      }
      if (MODE != 3 && MODE != 4) { sum0 += pv[0]; sum1 += pv[1]; sum2 += pv[2]; sum3 += pv[3]; }
      if (MODE != 2 && MODE != 4) {
        __nv_bfloat162 v0 = __floats2bfloat162_rn(pv[0], pv[1]);
        __nv_bfloat162 v1 = __floats2bfloat162_rn(pv[2], pv[3]);
        pk[2 * c] = *reinterpret_cast<uint32_t*>(&v0);
        pk[2 * c + 1] = *reinterpret_cast<uint32_t*>(&v1);
      } else {
        pk[2 * c] = __float_as_uint(pv[0]) ^ __float_as_uint(pv[1]);
        pk[2 * c + 1] = __float_as_uint(pv[2]) ^ __float_as_uint(pv[3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= pk[i];   // 32 LOP3-ish ops of overhead per 64 elements (reported separately)
    msc += 1e-3f;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __float_as_uint(sum0 + sum1 + sum2 + sum3);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// Software-pipelined variant: the FADD / F2FP consumers of chunk c are issued after the FFMA + MUFU of chunk c+1
// (CH elements per chunk), so a single in-order warp never waits on MUFU result latency.
// POLY: every POLY-th element uses the FMA-pipe polynomial (0 = none)
template <int CH, int POLY>
__global__ void kp(const float* in, uint32_t* out, unsigned long long* cyc, float sl2, float msc, int iters) {
  float s[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) s[i] = in[threadIdx.x * 64 + i];
  uint32_t acc = 0;
  float sum0 = 0, sum1 = 0, sum2 = 0, sum3 = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t pk[32];
    float pv[2][CH];
#pragma unroll
    for (int c = 0; c <= 64 / CH; ++c) {
      if (c < 64 / CH) {
#pragma unroll
        for (int e = 0; e < CH; ++e) {
          const float x = fmaf(s[c * CH + e], sl2, -msc);
          if (POLY > 0 && (e % POLY) == POLY - 1) pv[c & 1][e] = ex2_poly(x); else pv[c & 1][e] = ex2(x);
        }
      }
      if (c > 0) {
        const int cc = c - 1;
#pragma unroll
        for (int e = 0; e < CH; e += 4) {
          float* q = &pv[cc & 1][e];
          sum0 += q[0]; sum1 += q[1]; sum2 += q[2]; sum3 += q[3];
          __nv_bfloat162 v0 = __floats2bfloat162_rn(q[0], q[1]);
          __nv_bfloat162 v1 = __floats2bfloat162_rn(q[2], q[3]);
          pk[(cc * CH + e) / 2] = *reinterpret_cast<uint32_t*>(&v0);
          pk[(cc * CH + e) / 2 + 1] = *reinterpret_cast<uint32_t*>(&v1);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= pk[i];
    msc += 1e-3f;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __float_as_uint(sum0 + sum1 + sum2 + sum3);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int CH, int POLY>
void runp(const char* name, const float* in, uint32_t* out, unsigned long long* cyc) {
  printf("%-46s", name);
  const int iters = 2000;
  for (int threads : {128, 256, 384, 512}) {
    kp<CH, POLY><<<148, threads>>>(in, out, cyc, 0.18f, 3.0f, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf(" ERR %s", cudaGetErrorString(e)); continue; }
    unsigned long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    printf("  %dw/smsp: %6.2f cyc/elem/warp (%5.2f per SMSP)", threads / 128, avg / iters / 64, avg / iters / 64 / (threads / 128));
  }
  printf("\n");
}

// Candidate for the next attention iteration: packed f32x2 FMA / ADD (half the issue slots of the scale-subtract and
// the row sums) with every POLY-th exponential as an FMA-pipe polynomial, 32 elements per visit like attn4_kernel.
template <int POLY>
__global__ void kx(const float* in, uint32_t* out, unsigned long long* cyc, float sl2, float msc, int iters) {
  float s[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) s[i] = in[threadIdx.x * 32 + i];
  uint32_t acc = 0;
  unsigned long long sum01 = 0, sum23 = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t pk[16];
    unsigned long long sc, ms;
    asm("mov.b64 %0, {%1, %1};" : "=l"(sc) : "f"(sl2));
    asm("mov.b64 %0, {%1, %1};" : "=l"(ms) : "f"(-msc));
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      unsigned long long a01, a23;
      asm("mov.b64 %0, {%1, %2};" : "=l"(a01) : "f"(s[4 * c]), "f"(s[4 * c + 1]));
      asm("mov.b64 %0, {%1, %2};" : "=l"(a23) : "f"(s[4 * c + 2]), "f"(s[4 * c + 3]));
      asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a01) : "l"(sc), "l"(ms));
      asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a23) : "l"(sc), "l"(ms));
      float x[4], pv[4];
      asm("mov.b64 {%0, %1}, %2;" : "=f"(x[0]), "=f"(x[1]) : "l"(a01));
      asm("mov.b64 {%0, %1}, %2;" : "=f"(x[2]), "=f"(x[3]) : "l"(a23));
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = 4 * c + e;
        pv[e] = (POLY > 0 && (idx % (POLY > 0 ? POLY : 1)) == POLY - 1) ? ex2_poly(x[e]) : ex2(x[e]);
      }
      unsigned long long p01, p23;
      asm("mov.b64 %0, {%1, %2};" : "=l"(p01) : "f"(pv[0]), "f"(pv[1]));
      asm("mov.b64 %0, {%1, %2};" : "=l"(p23) : "f"(pv[2]), "f"(pv[3]));
      asm("add.rn.f32x2 %0, %0, %1;" : "+l"(sum01) : "l"(p01));
      asm("add.rn.f32x2 %0, %0, %1;" : "+l"(sum23) : "l"(p23));
      __nv_bfloat162 v0 = __floats2bfloat162_rn(pv[0], pv[1]);
      __nv_bfloat162 v1 = __floats2bfloat162_rn(pv[2], pv[3]);
      pk[2 * c] = *reinterpret_cast<uint32_t*>(&v0);
      pk[2 * c + 1] = *reinterpret_cast<uint32_t*>(&v1);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= pk[i];
    msc += 1e-3f;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (uint32_t)(sum01 ^ sum23);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int POLY>
void runx(const char* name, const float* in, uint32_t* out, unsigned long long* cyc) {
  printf("%-46s", name);
  const int iters = 4000;
  for (int threads : {128, 256, 384, 512}) {
    kx<POLY><<<148, threads>>>(in, out, cyc, 0.18f, 3.0f, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf(" ERR %s", cudaGetErrorString(e)); continue; }
    unsigned long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    printf("  %dw/smsp: %6.2f cyc/elem/warp (%5.2f per SMSP)", threads / 128, avg / iters / 32, avg / iters / 32 / (threads / 128));
  }
  printf("\n");
}

// Round-2 candidate (attn5): NO per-element FADD (the row sum comes from a tcgen05.mma against a ones tile), packed f32x2
// scale-subtract, KP of every 16 element PAIRS exponentiated by a packed f32x2 polynomial on the FMA pipe (the rest on
// MUFU), bf16 pack by F2FP (TRUNC = 0) or by a byte permute of the two high halves (TRUNC = 1; the truncation bias
// cancels because the row sum is taken from the same truncated values).
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
  unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ void ex2_poly2(unsigned long long x2, float& p0, float& p1) {
  float x0, x1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(x2));
  x0 = fminf(fmaxf(x0, -127.0f), 128.0f);
  x1 = fminf(fmaxf(x1, -127.0f), 128.0f);
  const unsigned long long xc = pack2(x0, x1), magic = pack2(12582912.0f, 12582912.0f), nmagic = pack2(-12582912.0f, -12582912.0f);
  unsigned long long xr, t, f, p;
  asm("add.rm.ftz.f32x2 %0, %1, %2;" : "=l"(xr) : "l"(xc), "l"(magic));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(xr), "l"(nmagic));
  const unsigned long long m1 = pack2(-1.0f, -1.0f);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(f) : "l"(t), "l"(m1), "l"(xc));
  const unsigned long long c3 = pack2(0.077119089663028717f, 0.077119089663028717f), c2 = pack2(0.227564394474029541f, 0.227564394474029541f),
                           c1 = pack2(0.695146143436431885f, 0.695146143436431885f), c0 = pack2(1.0f, 1.0f);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p) : "l"(f), "l"(c3), "l"(c2));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "+l"(p) : "l"(p), "l"(f), "l"(c1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "+l"(p) : "l"(p), "l"(f), "l"(c0));
  float q0, q1, r0, r1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(q0), "=f"(q1) : "l"(p));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(xr));
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(r0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(r1) << 23));
}

template <int KP, int TRUNC>
__global__ void ky(const float* in, uint32_t* out, unsigned long long* cyc, float sl2, float msc, int iters) {
  float s[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) s[i] = in[threadIdx.x * 32 + i];
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t pk[16];
    const unsigned long long sc = pack2(sl2, sl2), ms = pack2(-msc, -msc);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      unsigned long long a = pack2(s[2 * c], s[2 * c + 1]);
      asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(sc), "l"(ms));
      float p0, p1;
      // pairs c with (c * KP) % 16 < KP go to the polynomial: KP pairs of 16, evenly spread
      if (KP > 0 && ((c * KP) % 16) < KP) {
        ex2_poly2(a, p0, p1);
      } else {
        float x0, x1;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(a));
        p0 = ex2(x0); p1 = ex2(x1);
      }
      if (TRUNC) {
        asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(pk[c]) : "r"(__float_as_uint(p0)), "r"(__float_as_uint(p1)));
      } else {
        __nv_bfloat162 v = __floats2bfloat162_rn(p0, p1);
        pk[c] = *reinterpret_cast<uint32_t*>(&v);
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= pk[i];
    msc += 1e-3f;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KP, int TRUNC>
void runy(const char* name, const float* in, uint32_t* out, unsigned long long* cyc) {
  printf("%-46s", name);
  const int iters = 4000;
  for (int threads : {128, 256, 384, 512}) {
    ky<KP, TRUNC><<<148, threads>>>(in, out, cyc, 0.18f, 3.0f, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf(" ERR %s", cudaGetErrorString(e)); continue; }
    unsigned long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    printf("  %dw/smsp: %6.2f cyc/elem/warp (%5.2f per SMSP)", threads / 128, avg / iters / 32, avg / iters / 32 / (threads / 128));
  }
  printf("\n");
}

template <int MODE>
void run(const char* name, const float* in, uint32_t* out, unsigned long long* cyc) {
  printf("%-46s", name);
  const int iters = 2000;
  for (int threads : {128, 256, 384, 512}) {
    k<MODE><<<148, threads>>>(in, out, cyc, 0.18f, 3.0f, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf(" ERR %s", cudaGetErrorString(e)); continue; }
    unsigned long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    printf("  %dw/smsp: %6.2f cyc/elem/warp (%5.2f per SMSP)", threads / 128, avg / iters / 64, avg / iters / 64 / (threads / 128));
  }
  printf("\n");
}

int main() {
  float* in; uint32_t* out; unsigned long long* cyc;
  cudaMalloc(&in, 512 * 64 * 4); cudaMemset(in, 0, 512 * 64 * 4);
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);   // in: 512 threads x 64 floats (zeroed)
  run<0>("ffma + ex2 + fadd + cvt.bf16x2/2", in, out, cyc);
  run<1>("ffma2 + ex2 + fadd + cvt.bf16x2/2", in, out, cyc);
  run<2>("ffma + ex2 + fadd (no pack)", in, out, cyc);
  run<3>("ffma + ex2 + cvt (no fadd)", in, out, cyc);
  run<4>("ffma + ex2 only", in, out, cyc);
  run<5>("every 4th exp polynomial", in, out, cyc);
  runp<8, 0>("pipelined by 8", in, out, cyc);
  runp<16, 0>("pipelined by 16", in, out, cyc);
  runp<32, 0>("pipelined by 32", in, out, cyc);
  runp<16, 4>("pipelined by 16, every 4th polynomial", in, out, cyc);
  runp<16, 3>("pipelined by 16, every 3rd polynomial", in, out, cyc);
  runp<32, 4>("pipelined by 32, every 4th polynomial", in, out, cyc);
  runp<16, 2>("pipelined by 16, every 2nd polynomial", in, out, cyc);
  runp<16, 8>("pipelined by 16, every 8th polynomial", in, out, cyc);
  runx<0>("packed f32x2 fma/add, all MUFU (32/visit)", in, out, cyc);
  runx<8>("packed f32x2 fma/add, every 8th polynomial", in, out, cyc);
  runx<4>("packed f32x2 fma/add, every 4th polynomial", in, out, cyc);
  runx<3>("packed f32x2 fma/add, every 3rd polynomial", in, out, cyc);
  runy<0, 0>("r2: ffma2 + ex2 + cvt, no fadd, 0/16 poly", in, out, cyc);
  runy<4, 0>("r2: no fadd, 4/16 pairs packed poly", in, out, cyc);
  runy<5, 0>("r2: no fadd, 5/16 pairs packed poly", in, out, cyc);
  runy<6, 0>("r2: no fadd, 6/16 pairs packed poly", in, out, cyc);
  runy<8, 0>("r2: no fadd, 8/16 pairs packed poly", in, out, cyc);
  runy<0, 1>("r2: no fadd, 0/16 poly, prmt pack", in, out, cyc);
  runy<5, 1>("r2: no fadd, 5/16 poly, prmt pack", in, out, cyc);
  runy<6, 1>("r2: no fadd, 6/16 poly, prmt pack", in, out, cyc);
  runy<8, 1>("r2: no fadd, 8/16 poly, prmt pack", in, out, cyc);
  return 0;
}
