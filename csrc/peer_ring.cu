// Copy-engine ring transport for sequence-parallel attention: K|V shards move between the GPUs of one NVLink domain
// with cudaMemcpyAsync peer copies (DMA engines, no SMs — the attention kernel owns every SM) into buffers that the
// receiving process exports with CUDA IPC, ordered across processes by 32-bit stream memory operations
// (cuStreamWriteValue32 / cuStreamWaitValue32) instead of host synchronisation.  Plain C-ABI, no torch types.
// Replaces the NCCL send/recv hop of landiff_b200/parallel.py (the reference itself has no sequence parallelism).
#include "host_util.h"

#include <cstring>
#include <mutex>

namespace ld {

typedef CUresult (*WriteValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*WaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

static WriteValue32Fn g_write32 = nullptr;
static WaitValue32Fn g_wait32 = nullptr;

static void load_memops() {
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      g_write32 = reinterpret_cast<WriteValue32Fn>(p);
    p = nullptr;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      g_wait32 = reinterpret_cast<WaitValue32Fn>(p);
  });
}

}  // namespace ld

using namespace ld;

extern "C" int ld_ipc_alloc(size_t bytes, void** dev_ptr, unsigned char handle_out[64]) {
  int rc = check_device();
  if (rc != LD_OK) return rc;
  LD_CHECK_ARG(bytes > 0 && dev_ptr && handle_out, "ld_ipc_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  void* p = nullptr;
  LD_CHECK_CUDA(cudaMalloc(&p, bytes));
  // Zero the buffer (the flag words start at 0) and make sure the zeroing HAS HAPPENED before the handle leaves this
  // function: cudaMemset on device memory is asynchronous to the host and runs on the legacy default stream — torch's
  // compute stream — i.e. behind every kernel already queued there.  A peer that maps the buffer and raises a flag in the
  // meantime would have its flag wiped when the memset finally executes (seen as a lost arrival under load).  A private
  // non-blocking stream keeps the wait independent of whatever the compute stream is doing (e.g. a kernel polling a flag).
  cudaStream_t zs = nullptr;
  cudaError_t e = cudaStreamCreateWithFlags(&zs, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMemsetAsync(p, 0, bytes, zs);
  if (e == cudaSuccess) e = cudaStreamSynchronize(zs);
  if (zs != nullptr) cudaStreamDestroy(zs);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("ld_ipc_alloc: %s", cudaGetErrorString(e));
    return LD_ERR_CUDA;
  }
  memcpy(handle_out, &h, 64);
  *dev_ptr = p;
  return LD_OK;
}

extern "C" int ld_ipc_open(const unsigned char handle[64], void** dev_ptr) {
  int rc = check_device();
  if (rc != LD_OK) return rc;
  LD_CHECK_ARG(handle && dev_ptr, "ld_ipc_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  LD_CHECK_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return LD_OK;
}

extern "C" int ld_ipc_close(void* dev_ptr) {
  LD_CHECK_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return LD_OK;
}

extern "C" int ld_ipc_free(void* dev_ptr) {
  LD_CHECK_CUDA(cudaFree(dev_ptr));
  return LD_OK;
}

extern "C" int ld_copy_async(void* dst, const void* src, size_t bytes, void* stream) {
  LD_CHECK_ARG(dst && src, "ld_copy_async: null pointer");
  // unified addressing resolves local / peer-mapped pointers; peer copies run on the copy engines over NVLink
  LD_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return LD_OK;
}

extern "C" int ld_stream_write_u32(void* dev_addr, unsigned int value, void* stream) {
  load_memops();
  LD_CHECK_ARG(dev_addr && (reinterpret_cast<uintptr_t>(dev_addr) & 3) == 0, "ld_stream_write_u32: bad address");
  if (!g_write32) {
    set_error("cuStreamWriteValue32 driver entry point unavailable");
    return LD_ERR_CUDA;
  }
  CUresult r = g_write32((CUstream)stream, (CUdeviceptr)dev_addr, value, CU_STREAM_WRITE_VALUE_DEFAULT);
  if (r != CUDA_SUCCESS) {
    set_error("cuStreamWriteValue32 failed with CUresult %d", (int)r);
    return LD_ERR_CUDA;
  }
  return LD_OK;
}

extern "C" int ld_stream_wait_geq_u32(void* dev_addr, unsigned int value, void* stream) {
  load_memops();
  LD_CHECK_ARG(dev_addr && (reinterpret_cast<uintptr_t>(dev_addr) & 3) == 0, "ld_stream_wait_geq_u32: bad address");
  if (!g_wait32) {
    set_error("cuStreamWaitValue32 driver entry point unavailable");
    return LD_ERR_CUDA;
  }
  // GEQ compares (int32)(*addr - value) >= 0: transfer ids are monotonically increasing and wrap safely
  CUresult r = g_wait32((CUstream)stream, (CUdeviceptr)dev_addr, value, CU_STREAM_WAIT_VALUE_GEQ);
  if (r != CUDA_SUCCESS) {
    set_error("cuStreamWaitValue32 failed with CUresult %d", (int)r);
    return LD_ERR_CUDA;
  }
  return LD_OK;
}
