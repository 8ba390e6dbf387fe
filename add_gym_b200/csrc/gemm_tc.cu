// Tensor-core (tcgen05) dense layer modes; see gemm.cu for the shared argument struct.
#include "common.cuh"
#include "addk.h"

int addk_gemm_tc(cudaStream_t st, const addk_gemm_args& a, int precision) {
  (void)st; (void)a; (void)precision;
  addk_set_error("tensor-core GEMM mode not built");
  return ADDK_ERR_UNSUPPORTED;
}
