// tcgen05.mma dispatch-rate microbenchmark (sm_100a): cycles per MMA instruction as a function of N, operand
// source (SS / TS), accumulator dependence and commit frequency.  Operand contents are garbage.  The issuing warp
// runs warp-uniform code with an elect.sync leader (the pattern ptxas turns into back-to-back UTCHMMA).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I csrc -o tools/mma_bench tools/mma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace ld;

// MODE 0: SS only; 1: TS only (B N-major); 2: attention pattern: 4 x SS(N=64) then 4 x TS(N=64) alternating
// MODE 3: TS with K-major B; 4: attn4 pattern: 4 x TS(K-major B) then 4 x TS(N-major B)
// MODE 5 (round 2, attn5 pattern): per group of 12: 4 x TS(K-major B, N) [S], 4 x TS(N-major B, N) [PV], 4 x TS(K-major B,
//         N = 16 against a ones tile) [row sums]; MODE 6: the same with the row-sum MMAs folded into PV as N + 16 (N-major)
template <int N, int MODE, int NACC, int COMMIT_EVERY>
__global__ void __launch_bounds__(128, 1) k(long long* out, int total) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    constexpr uint32_t idesc_ss = make_idesc_bf16(128, N, false, false);
    constexpr uint32_t idesc_ts = make_idesc_bf16(128, N, false, true);
    const uint64_t adesc = make_sdesc_sw128(smem_u32(smem));
    const uint64_t bdesc = make_sdesc_sw128(smem_u32(smem + 32768));
    const bool leader = elect_one();
    long long t0 = clock64();
    for (int i = 0; i < total; i += 4) {
      const int g = i / 4;
      const uint32_t d = tm + (g % NACC) * N;
      const bool ts = MODE == 1 || (MODE == 2 && (g & 1)) || (MODE == 4 && (g & 1));
      const bool tsk = MODE == 3 || (MODE == 4 && !(g & 1));
      if (MODE == 5 || MODE == 6) {
        constexpr uint32_t idesc_l = make_idesc_bf16(128, 16, false, false);
        constexpr uint32_t idesc_pvl = make_idesc_bf16(128, N + 16, false, true);
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_ts(tm + (g & 1) * 64, tm + 416 + ks * 8, bdesc + 2 * ks, idesc_ss, 1);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_ts(tm + 256, tm + 128 + (g & 1) * 64 + ks * 8, bdesc + 128 * ks, MODE == 6 ? idesc_pvl : idesc_ts, 1);
          if (MODE == 5) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_ts(tm + 384, tm + 128 + (g & 1) * 64 + ks * 8, adesc, idesc_l, 1);
          }
          if (COMMIT_EVERY) umma_commit(&bar[1]);
        }
        continue;
      }
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if (tsk) umma_ts(d, tm + 416 + ks * 8, bdesc + 2 * ks, idesc_ss, 1);
          else if (ts) umma_ts(d, tm + 448 + ks * 8, bdesc + 128 * ks, idesc_ts, 1);
          else umma_ss(d, adesc + 2 * ks, bdesc + 2 * ks, idesc_ss, 1);
        }
        if (COMMIT_EVERY && (g % COMMIT_EVERY) == COMMIT_EVERY - 1) umma_commit(&bar[1]);
      }
    }
    long long t1 = clock64();
    if (leader) umma_commit(&bar[0]);
    mbar_wait(&bar[0], 0);
    long long t2 = clock64();
    if (leader) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}

template <int N, int MODE, int NACC, int COMMIT_EVERY>
void run(long long* out) {
  const int total = 1024;
  auto kern = k<N, MODE, NACC, COMMIT_EVERY>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  printf("N=%3d %-4s acc=%d commit/%d groups |", N, MODE == 0 ? "SS" : MODE == 1 ? "TS" : MODE == 2 ? "S+TS" : MODE == 3 ? "TSk" : MODE == 4 ? "TSk+TS" : MODE == 5 ? "S+PV+L16 (x/4: per group of 4 dispatched)" : "S+PV(N+16)", NACC, COMMIT_EVERY);
  for (int grid : {148, 1}) {
    kern<<<grid, 128, 100 * 1024>>>(out, total);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("ERR %s\n", cudaGetErrorString(e)); exit(1); }
    long long h[296];
    cudaMemcpy(h, out, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
    double a = 0, b = 0;
    for (int i = 0; i < grid; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
    printf("  grid %3d: issue %6.1f  complete %6.1f cyc/MMA (ideal %d)", grid, a / grid / total, b / grid / total, N / 2);
  }
  printf("\n");
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * 2 * sizeof(long long));
  run<64, 0, 1, 0>(out);  run<64, 0, 2, 0>(out);  run<64, 0, 4, 0>(out);
  run<128, 0, 1, 0>(out); run<128, 0, 2, 0>(out);
  run<192, 0, 1, 0>(out); run<192, 0, 2, 0>(out);
  run<256, 0, 1, 0>(out); run<256, 0, 2, 0>(out);
  run<64, 1, 1, 0>(out);  run<64, 1, 2, 0>(out);
  run<128, 1, 1, 0>(out); run<128, 1, 2, 0>(out);
  run<64, 2, 1, 0>(out);  run<64, 2, 2, 0>(out);  run<64, 2, 4, 0>(out);
  run<64, 0, 2, 1>(out);  run<64, 2, 4, 1>(out);  run<128, 0, 2, 1>(out); run<192, 0, 2, 4>(out);
  run<32, 0, 2, 0>(out);  run<16, 0, 2, 0>(out);
  run<64, 3, 1, 0>(out);  run<64, 3, 2, 0>(out);  run<128, 3, 2, 0>(out);  run<64, 4, 4, 1>(out);
  run<16, 3, 2, 0>(out);  run<32, 3, 2, 0>(out);   // TS-mode dispatch floor at small N (row-sum MMA against a ones tile)
  run<64, 5, 1, 1>(out);  run<64, 6, 1, 1>(out);   // cycles are per counted MMA (4 per group): x4 = per 64-key sub-block
  return 0;
}
