// conv_tc_kernel<1, *> instantiations (see conv_tc_kernel.cuh); a separate translation unit per MT keeps the build parallel.
#include "conv_tc_kernel.cuh"

namespace fs2 {

cudaError_t conv_tc_prepare_mt1(int smem_bytes) {
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  return e;
}

void conv_tc_launch_mt1(const TcP& p, unsigned grid, size_t smem, cudaStream_t s) {
  if (p.TG == 3) conv_tc_kernel<1, 3><<<grid, TC_THREADS, smem, s>>>(p);
  else if (p.TG == 2) conv_tc_kernel<1, 2><<<grid, TC_THREADS, smem, s>>>(p);
  else conv_tc_kernel<1, 1><<<grid, TC_THREADS, smem, s>>>(p);
}

}  // namespace fs2
