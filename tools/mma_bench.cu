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
  printf("N=%3d %-4s acc=%d commit/%d groups |", N, MODE == 0 ? "SS" : MODE == 1 ? "TS" : MODE == 2 ? "S+TS" : MODE == 3 ? "TSk" : "TSk+TS", NACC, COMMIT_EVERY);
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
  return 0;
}
