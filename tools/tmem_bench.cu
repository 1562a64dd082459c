// TMEM load/store bandwidth microbenchmark (sm_100a): NW warps loop tcgen05.ld (or st) 32x32b.x32 over their lane
// quadrant; reports bytes per clock per SM.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I csrc ...
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace ld;

template <int MODE>   // 0: ld x32 + wait each   1: 4 x (ld x32) then wait   2: st x32   3: ld x32 while an MMA stream runs
__global__ void __launch_bounds__(640, 1) k(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t base = tm + (uint32_t((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  if (MODE == 3 && warp == 0) {
    // background MMA stream: TS N=64 + SS N=64 alternating, like the attention kernel
    const bool leader = elect_one();
    constexpr uint32_t idesc_ts = make_idesc_bf16(128, 64, false, true);
    constexpr uint32_t idesc_ss = make_idesc_bf16(128, 64, false, false);
    const uint64_t adesc = make_sdesc_sw128(smem_u32(smem));
    const uint64_t bdesc = make_sdesc_sw128(smem_u32(smem + 32768));
    long long t0 = clock64();
    for (int i = 0; i < iters * 2; ++i) {
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_ts(tm + 256, tm + 448 + ks * 8, bdesc + 128 * ks, idesc_ts, 1);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_ss(tm + 320, adesc + 2 * ks, bdesc + 2 * ks, idesc_ss, 1);
      }
    }
    if (leader) umma_commit(&bar[0]);
    mbar_wait(&bar[0], 0);
    long long t1 = clock64();
    if (leader) out[blockIdx.x * 32 + 31] = t1 - t0;
  } else if (warp >= 4) {
    uint32_t r[32], r2[32], r3[32], r4[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t col = (it & 1) * 128;
      if (MODE == 0 || MODE == 3) {
        LD_TMEM_LD32(base + col, r);
        tmem_ld_wait();
        acc ^= r[0] ^ r[31];
      } else if (MODE == 1) {
        LD_TMEM_LD32(base + col, r);
        LD_TMEM_LD32(base + col + 32, r2);
        LD_TMEM_LD32(base + col + 64, r3);
        LD_TMEM_LD32(base + col + 96, r4);
        tmem_ld_wait();
        acc ^= r[0] ^ r2[1] ^ r3[2] ^ r4[3];
      } else {
        LD_TMEM_ST32(base + col, r);
        tmem_st_wait();
      }
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 32 + warp] = t1 - t0;
    if (acc == 0x12345) out[0] = acc;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}

template <int MODE>
void run(const char* name, long long* out) {
  const int iters = 4096;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  printf("%-44s", name);
  for (int warps : {8, 12, 20}) {   // 4, 8, 16 data warps
    k<MODE><<<148, warps * 32, 100 * 1024>>>(out, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf(" ERR %s\n", cudaGetErrorString(e)); exit(1); }
    long long h[148 * 32];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int w = 4; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
    const int nw = warps - 4;
    const double bytes = (double)nw * iters * 32 * 32 * 4 * (MODE == 1 ? 4 : 1);
    printf("  %2d warps: %6.1f B/clk/SM", nw, bytes / mx);
    if (MODE == 3) printf(" (MMA %5.1f clk each)", (double)h[31] / (iters * 2 * 8));
  }
  printf("\n");
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * 32 * sizeof(long long));
  cudaMemset(out, 0, 148 * 32 * sizeof(long long));
  run<0>("tcgen05.ld 32x32b.x32, wait each", out);
  run<1>("4 x tcgen05.ld 32x32b.x32, one wait", out);
  run<2>("tcgen05.st 32x32b.x32, wait each", out);
  run<3>("tcgen05.ld x32 + concurrent TS/SS N=64 MMAs", out);
  return 0;
}
