// instantiation unit: float, 1 element(s) per vector
#include "spmm_kernels.cuh"
namespace rgbmp {
int spmm_dispatch_f32s(const SpmmParams& p, int G, int V, int U, cudaStream_t st) {
  return dispatch_g<float, 1>(p, G, V, U, st);
}
}  // namespace rgbmp
