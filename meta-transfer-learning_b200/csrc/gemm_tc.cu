// tcgen05 / TMA / TMEM GEMM (placeholder until the tensor-core path lands; never eligible).
#include "kernels.h"
bool k_gemm_tc_eligible(const GemmArgs&) { return false; }
int k_gemm_tc(const GemmArgs&, int, cudaStream_t) {
  mtl_set_error("tcgen05 GEMM path not built");
  return MTL_ERR_ARG;
}
