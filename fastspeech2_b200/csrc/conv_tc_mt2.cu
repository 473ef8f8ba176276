// conv_tc_kernel<2, *> instantiations (see conv_tc_kernel.cuh); a separate translation unit per MT keeps the build parallel.
#include "conv_tc_kernel.cuh"

namespace fs2 {

cudaError_t conv_tc_prepare_mt2(int smem_bytes) {
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  return e;
}

void conv_tc_launch_mt2(const TcP& p, unsigned grid, size_t smem, cudaStream_t s) {
  void (*kern)(const TcP) = conv_tc_kernel<2, 1>;
  if (p.TG == 2) kern = conv_tc_kernel<2, 2>;
  conv_tc_launch(kern, p, grid, smem, s);
}

}  // namespace fs2
