// C-ABI plumbing: error string, device check, TMA tensor-map encoder.
#include "host_util.h"

#include <cstring>
#include <mutex>

namespace ld {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_dev_ok[64];       // 0 unknown, 1 ok, -1 bad
static int g_dev_sms[64];

int check_device() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no CUDA device: %s (landiff_b200 has no CPU fallback)", cudaGetErrorString(e));
    return LD_ERR_DEVICE;
  }
  if (dev < 0 || dev >= 64) dev = 0;
  if (g_dev_ok[dev] == 0) {
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) {
      set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
      return LD_ERR_DEVICE;
    }
    g_dev_sms[dev] = prop.multiProcessorCount;
    g_dev_ok[dev] = (prop.major == 10) ? 1 : -1;
    if (g_dev_ok[dev] < 0)
      set_error("device %d is sm_%d%d; landiff_b200 kernels are built for sm_100a only", dev, prop.major, prop.minor);
  }
  if (g_dev_ok[dev] < 0) {
    set_error("current device is not sm_100 (landiff_b200 has no fallback path)");
    return LD_ERR_DEVICE;
  }
  return LD_OK;
}

int sm_count() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  return g_dev_sms[dev] > 0 ? g_dev_sms[dev] : 148;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled driver entry point unavailable");
    return LD_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(gptr) & 15) != 0) {
    set_error("TMA source pointer %p not 16-byte aligned", gptr);
    return LD_ERR_ARG;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i - 1];
      if (gstr[i - 1] % 16 != 0) {
        set_error("TMA stride %llu not a multiple of 16 bytes", (unsigned long long)gstr[i - 1]);
        return LD_ERR_ARG;
      }
    }
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(gptr), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu x %llu, box %u x %u)", (int)r, rank,
              (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 1), bx[0], rank > 1 ? bx[1] : 1);
    return LD_ERR_CUDA;
  }
  return LD_OK;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeIm2colFn get_encode_im2col() {
  static EncodeIm2colFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeIm2colFn>(p);
  });
  return fn;
}

// im2col-mode tensor map over channels-last bf16 activations [F, H, W, C] for a 3x3 / stride 1 / padding 1 convolution:
// one load brings `pixels` consecutive output positions (raster order over x, y, frame) x 64 channels of ONE filter tap into
// a SWIZZLE_128B tile; the tap is given per load as the (kx, ky) offsets, out-of-image taps are zero-filled (the padding).
// Bounding box: lower corner = -padding = -1, upper corner = padding - (3 - 1) = -1 in x and y.
int make_tmap_im2col3x3_bf16(CUtensorMap* out, const void* gptr, int C, int W, int H, int F, int pixels) {
  EncodeIm2colFn enc = get_encode_im2col();
  if (!enc) {
    set_error("cuTensorMapEncodeIm2col driver entry point unavailable");
    return LD_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(gptr) & 15) != 0 || C % 64 != 0) {
    set_error("im2col TMA source %p must be 16-byte aligned with C=%d a multiple of 64", gptr, C);
    return LD_ERR_ARG;
  }
  const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)F};
  const cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const int lower[2] = {-1, -1}, upper[2] = {-1, -1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(gptr), gdim, gstr, lower, upper, 64u,
                   (cuuint32_t)pixels, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeIm2col failed with CUresult %d (C %d W %d H %d F %d)", (int)r, C, W, H, F);
    return LD_ERR_CUDA;
  }
  // drivers up to 13.1 set a descriptor bit that breaks im2col loads from tensors smaller than 128 KB (the same fix-up is
  // applied by CUTLASS, cute/atom/copy_traits_sm90_im2col.hpp)
  int drv = 0;
  if (cudaDriverGetVersion(&drv) == cudaSuccess && drv <= 13010 && (uint64_t)F * H * W * C * 2 < 131072)
    reinterpret_cast<uint64_t*>(out)[1] &= ~(1llu << 21);
  return LD_OK;
}

}  // namespace ld

extern "C" {

const char* ld_last_error(void) { return ld::g_err; }
int ld_abi_version(void) { return 7; }

/* sizes of the structs that cross the boundary, so that a binding can verify its mirror of them at load time */
int ld_struct_sizes(int* gemm_args, int* kv_shard, int* token_blocks) {
  if (gemm_args) *gemm_args = (int)sizeof(ld_gemm_args);
  if (kv_shard) *kv_shard = (int)sizeof(ld_kv_shard);
  if (token_blocks) *token_blocks = (int)sizeof(ld_token_blocks);
  return LD_OK;
}
int ld_device_check(int* sms) {
  int rc = ld::check_device();
  if (rc == LD_OK && sms) *sms = ld::sm_count();
  return rc;
}

}  // extern "C"
