// Host-side helpers shared by the C-ABI translation units: thread-local error string, CUDA error checks and
// the TMA tensor-map encoder (driver entry point fetched through the runtime, so libcuda is not linked).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../include/landiff_b200.h"

namespace ld {

void set_error(const char* fmt, ...);
int check_device();  // LD_OK or LD_ERR_DEVICE (cached per device)
int sm_count();

// rank-2/3 bf16 tensor map, SWIZZLE_128B, inner box extent 64 elements (128 B).
// dims/strides innermost first; strides in BYTES for dims 1.. (dim 0 is contiguous).
int make_tmap_bf16(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);
// im2col-mode map over channels-last bf16 [F, H, W, C] for a 3x3 / stride 1 / padding 1 convolution (abi.cu)
int make_tmap_im2col3x3_bf16(CUtensorMap* out, const void* gptr, int C, int W, int H, int F, int pixels);

#define LD_CHECK_ARG(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      ld::set_error(__VA_ARGS__);    \
      return LD_ERR_ARG;             \
    }                                \
  } while (0)

#define LD_CHECK_CUDA(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ld::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return LD_ERR_CUDA;                                                                    \
    }                                                                                        \
  } while (0)

}  // namespace ld
