// instantiation unit: __nv_bfloat16, 8 element(s) per vector
#include "spmm_kernels.cuh"
namespace rgbmp {
int spmm_dispatch_bf16(const SpmmParams& p, int G, int V, int U, cudaStream_t st) {
  return dispatch_g<__nv_bfloat16, 8>(p, G, V, U, st);
}
}  // namespace rgbmp
