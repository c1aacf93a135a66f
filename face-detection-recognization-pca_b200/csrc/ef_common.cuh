// Shared helpers for libeigenfaces_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "../../include/eigenfaces_b200.h"

namespace ef {

extern std::atomic<int64_t> g_launches;
void set_error_detail(const char* what, cudaError_t e);

#define EF_CUDA(call)                                   \
  do {                                                  \
    cudaError_t e__ = (call);                           \
    if (e__ != cudaSuccess) {                           \
      ::ef::set_error_detail(#call, e__);               \
      return EF_ERR_CUDA;                               \
    }                                                   \
  } while (0)

#define EF_TRY(call)              \
  do {                            \
    int s__ = (call);             \
    if (s__ != EF_OK) return s__; \
  } while (0)

// Every kernel launch of the library goes through this macro so that ef_launch_count() is exact.
#define EF_LAUNCH(kernel, grid, block, smem, stream, ...)                 \
  do {                                                                    \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);           \
    ::ef::g_launches.fetch_add(1, std::memory_order_relaxed);             \
    cudaError_t e__ = cudaPeekAtLastError();                              \
    if (e__ != cudaSuccess) {                                             \
      ::ef::set_error_detail(#kernel, e__);                               \
      return EF_ERR_CUDA;                                                 \
    }                                                                     \
  } while (0)

// The same launch with programmatic stream serialization: the grid may be SCHEDULED while the previous kernel of the
// stream is still running (once all of that kernel's CTAs have executed griddepcontrol.launch_dependents or exited); the
// kernel must execute griddepcontrol.wait before it touches anything an earlier kernel of the stream writes.
// EF_NO_PDL=1 launches normally.
#define EF_LAUNCH_PDL(kernel, grid, block, smem, stream, ...)                                   \
  do {                                                                                          \
    cudaLaunchConfig_t cfg__{};                                                                 \
    cfg__.gridDim = dim3(grid);                                                                 \
    cfg__.blockDim = dim3(block);                                                               \
    cfg__.dynamicSmemBytes = (smem);                                                            \
    cfg__.stream = (stream);                                                                    \
    cudaLaunchAttribute attr__[1];                                                              \
    attr__[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                          \
    attr__[0].val.programmaticStreamSerializationAllowed = 1;                                   \
    cfg__.attrs = attr__;                                                                       \
    cfg__.numAttrs = getenv("EF_NO_PDL") ? 0 : 1;                                               \
    cudaError_t e__ = cudaLaunchKernelEx(&cfg__, kernel, __VA_ARGS__);                          \
    ::ef::g_launches.fetch_add(1, std::memory_order_relaxed);                                   \
    if (e__ != cudaSuccess) {                                                                   \
      ::ef::set_error_detail(#kernel, e__);                                                     \
      return EF_ERR_CUDA;                                                                       \
    }                                                                                           \
  } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per DEVICE: one process may drive several GPUs, so the "already
// raised to N bytes" bookkeeping is kept per device ordinal (one static table per call site / template instantiation).
#define EF_ENSURE_SMEM(kernel, bytes)                                                                           \
  do {                                                                                                          \
    static size_t attr__[64] = {0};                                                                             \
    int dev__ = 0;                                                                                              \
    EF_CUDA(cudaGetDevice(&dev__));                                                                             \
    if (dev__ < 0 || dev__ >= 64) return EF_ERR_UNSUPPORTED;                                                    \
    if ((size_t)(bytes) > attr__[dev__]) {                                                                      \
      EF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));         \
      attr__[dev__] = (size_t)(bytes);                                                                          \
    }                                                                                                           \
  } while (0)

// One-time per-device setup block: `if (EF_FIRST_ON_DEVICE()) { ... }`
#define EF_FIRST_ON_DEVICE()                                                \
  ([]() -> bool {                                                           \
    static bool done__[64] = {false};                                       \
    int dev__ = 0;                                                          \
    if (cudaGetDevice(&dev__) != cudaSuccess || dev__ < 0 || dev__ >= 64) return true; \
    if (done__[dev__]) return false;                                        \
    done__[dev__] = true;                                                   \
    return true;                                                            \
  }())

static inline cudaStream_t as_stream(ef_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

int sm_count();  // cached SM count of the current device (148 on B200)

// Simple RAII device buffer for the host-side orchestration code.
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  int ensure(size_t n) {
    if (n <= bytes) return EF_OK;
    release();
    cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess) {
      set_error_detail("cudaMalloc", e);
      p = nullptr;
      return EF_ERR_NOMEM;
    }
    bytes = n;
    return EF_OK;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

// ---------------------------------------------------------------------------------------------- device
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace ef
