// instantiation unit: float, 4 element(s) per vector
#include "spmm_kernels.cuh"
namespace rgbmp {
int spmm_dispatch_f32v(const SpmmParams& p, int G, int V, int U, cudaStream_t st) {
  return dispatch_g<float, 4>(p, G, V, U, st);
}
}  // namespace rgbmp
