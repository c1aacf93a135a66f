// Library-wide state: ABI version, error text, launch counter.
#include "ef_common.cuh"

namespace ef {

std::atomic<int64_t> g_launches{0};
static thread_local char g_detail[512] = "";

void set_error_detail(const char* what, cudaError_t e) {
  snprintf(g_detail, sizeof(g_detail), "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  cudaGetLastError();  // clear the sticky-less error so later calls report their own
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace ef

extern "C" {

int ef_version(void) { return EF_ABI_VERSION; }

const char* ef_error_string(int status) {
  switch (status) {
    case EF_OK: return "ok";
    case EF_ERR_INVALID: return "invalid argument";
    case EF_ERR_UNSUPPORTED: return "unsupported shape";
    case EF_ERR_CUDA: return "CUDA error (no device, or a runtime failure; see ef_last_error_detail)";
    case EF_ERR_NOMEM: return "out of memory";
    case EF_ERR_NOCONVERGE: return "Jacobi eigensolver did not converge";
    default: return "unknown status";
  }
}

const char* ef_last_error_detail(void) { return ef::g_detail; }

int64_t ef_launch_count(void) { return ef::g_launches.load(std::memory_order_relaxed); }

int ef_device_sm_count(int* out) {
  if (!out) return EF_ERR_INVALID;
  int dev = 0;
  EF_CUDA(cudaGetDevice(&dev));
  int n = 0;
  EF_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  *out = n;
  return EF_OK;
}

}  // extern "C"
