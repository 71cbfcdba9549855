// tcgen05 path of the dense projections (placeholder until the UMMA kernels land).
#include "common.cuh"

namespace magat {
bool tc_supported(const magat_gat_fwd_args*) { return false; }
int forward_tc(const magat_gat_fwd_args*, cudaStream_t) {
  set_error("tcgen05 path not built");
  return MAGAT_E_UNSUPPORTED;
}
}  // namespace magat
