// K3/K4/K7 (tensor-core arm): the BD_TC_BF16 instantiations of the persistent tcgen05 implicit-GEMM kernel
// (gemm_tc_impl.cuh): bf16 operands, fp32 accumulation.
#include "gemm_tc_impl.cuh"

// bf16 A tensor in HBM: 64-element k-blocks, no splitter
int bd_tc_launch_bf16d(int tbk, int tbn, const bd_gemm_desc& d, const TileGeom& g, int items, cudaStream_t st) {
  (void)tbk;
  return tbn == 256 ? launch_tc_persist<64, 256, BD_TC_BF16D>(d, g, items, st) : launch_tc_width<64, BD_TC_BF16D>(tbn, d, g, items, st);
}

int bd_tc_launch_bf16(int tbk, int tbn, const bd_gemm_desc& d, const TileGeom& g, int items, cudaStream_t st) {
  if (tbk == 32)
    return tbn == 256 ? launch_tc_persist<32, 256, BD_TC_BF16>(d, g, items, st) : launch_tc_width<32, BD_TC_BF16>(tbn, d, g, items, st);
  return tbn == 256 ? launch_tc_persist<16, 256, BD_TC_BF16>(d, g, items, st) : launch_tc_width<16, BD_TC_BF16>(tbn, d, g, items, st);
}
