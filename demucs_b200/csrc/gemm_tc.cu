// K3/K7 (tensor-core arm): tcgen05 / TMEM / TMA GEMM.  Placeholder dispatcher until the kernel lands:
// reports "not handled" so that bd_conv_gemm falls through to the exact fp32 arm.
#include "common.cuh"
#include "../../include/demucs_b200.h"

int bd_conv_gemm_tc(const bd_gemm_desc* d, void* stream, int* handled) {
  (void)d;
  (void)stream;
  *handled = 0;
  return BD_OK;
}
