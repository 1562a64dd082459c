// Instruction-throughput microbenchmarks for the softmax inner loop on sm_100a (B200):
// MUFU ex2 (f32 / f16x2 / bf16x2), packed f32x2 FMA/ADD, 3-input max, cvt packs, tanh.
// Prints cycles per warp-instruction per SMSP at 1/2/4 warps per SMSP.   nvcc -arch=sm_100a -O3 microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

template <int OP>
__global__ void k(float* out, unsigned long long* cyc, float seed) {
  float a[CHAINS];
  unsigned u[CHAINS];
  unsigned long long d[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) {
    a[i] = seed * (i + 1 + threadIdx.x * 0.001f);
    u[i] = __float_as_uint(a[i]);
    d[i] = ((unsigned long long)u[i] << 32) | u[i];
  }
  const float c1 = seed * 0.5f, c2 = seed * 0.25f;
  unsigned long long cc = ((unsigned long long)__float_as_uint(c1) << 32) | __float_as_uint(c2);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
      if (OP == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
      if (OP == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(cc));
      if (OP == 4) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(cc));
      if (OP == 5) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c1), "f"(c2));
      if (OP == 6) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(__uint_as_float(u[i])), "f"(c1));
      if (OP == 7) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(__uint_as_float(u[i])), "f"(c1));
      if (OP == 16) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(u[i]) : "r"(__float_as_uint(c1)));
      if (OP == 17) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(__float_as_uint(c1)), "r"(__float_as_uint(c2)));
      if (OP == 18) {  // ex2 + cvt pack interleaved: do they share a pipe?
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(__uint_as_float(u[i])), "f"(c1));
      }
      if (OP == 19) {  // ex2 + (iadd, prmt) interleaved
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("add.s32 %0, %0, 0x8000;" : "+r"(u[i]));
        asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(u[i]) : "r"(__float_as_uint(c1)));
      }
      if (OP == 20) asm volatile("cvt.rn.bf16.f32 %0, %1;" : "=h"(*(unsigned short*)&u[i]) : "f"(__uint_as_float(u[i])));
      if (OP == 8) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 9) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c1), "f"(c2));
      if (OP == 10) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c1));
      if (OP == 11) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(cc));
      if (OP == 12) {  // mixed: 3 ex2.f32 + 1 fma chain of 6 (poly-like) to see co-issue of MUFU and FMA pipes
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(u[i]) : "f"(c1), "f"(c2));
      }
      if (OP == 13) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c1));
      if (OP == 14) asm volatile("shl.b32 %0, %0, 23;" : "+r"(u[i]));
      if (OP == 15) asm volatile("add.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(__float_as_uint(c1)));
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += a[i] + __uint_as_float(u[i]) + (float)(d[i] & 0xff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name) {
  float* out;
  unsigned long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 148 * sizeof(unsigned long long));
  printf("%-28s", name);
  for (int threads : {128, 256, 512}) {
    k<OP><<<148, threads>>>(out, cyc, 0.5f);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf(" ERR %s", cudaGetErrorString(e)); continue; }
    unsigned long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    int warps_per_smsp = threads / 128;
    double instr_per_smsp = (double)ITERS * CHAINS * warps_per_smsp * (OP == 12 || OP == 18 ? 2 : (OP == 19 ? 3 : 1));
    printf("  %dw/smsp: %6.2f cyc/winstr", warps_per_smsp, avg / instr_per_smsp);
  }
  printf("\n");
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("ex2.approx.ftz.f32");
  run<1>("ex2.approx.f16x2");
  run<2>("ex2.approx.ftz.bf16x2");
  run<8>("tanh.approx.f32");
  run<9>("fma.rn.f32");
  run<3>("fma.rn.f32x2");
  run<10>("add.rn.f32");
  run<4>("add.rn.f32x2");
  run<11>("mul.rn.f32x2");
  run<13>("max.f32 (2-input)");
  run<5>("max.f32 (3-input)");
  run<6>("cvt.rn.bf16x2.f32");
  run<7>("cvt.rn.f16x2.f32");
  run<14>("shl.b32");
  run<15>("add.s32");
  run<12>("ex2.f32 + fma.f32 interleaved");
  run<16>("prmt.b32");
  run<17>("lop3.b32");
  run<18>("ex2.f32 + cvt.bf16x2 interleaved");
  run<19>("ex2.f32 + add.s32 + prmt");
  return 0;
}
