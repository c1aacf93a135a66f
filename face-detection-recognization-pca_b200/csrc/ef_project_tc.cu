// K2a on tensor cores: tcgen05 kind::i8 (uint8 crops x int8 digit planes -> int32 in TMEM).
// Placeholder until the tcgen05 kernel lands: reports "unsupported" so the dp4a path (same integers) is used.
#include "ef_common.cuh"
#include "ef_internal.cuh"

namespace ef {

bool project_tc_supported(int, int, int) { return false; }

int project_tc(const uint8_t*, int64_t, int, int, const int8_t*, int64_t, int, int32_t*, int, cudaStream_t) {
  return EF_ERR_UNSUPPORTED;
}

}  // namespace ef
