// Micro-benchmark: random ROW GATHER peak (sm_100a).  The roofline denominator for aggregation kernels whose
// feature matrix is L2-resident (C2 / C3 of BASELINE.json: arxiv-shaped SAGE, Reddit-shaped GAT) -- an HBM copy
// peak says nothing about them.  Every group of `row_bytes / 16` lanes reads random rows of a table with 16-byte
// loads, 8 independent rows in flight per lane, ids from a counter hash (no index array is read, nothing but a
// 16-byte checksum per group is written): the rate at which this device can deliver gathered rows to registers,
// from L2 when the table fits it, from HBM when it does not.
#include "common.cuh"

namespace rgbmp {

__device__ __forceinline__ uint32_t mb_hash(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

template <int G>
__global__ void __launch_bounds__(256, 4)
gather_peak_kernel(const float4* __restrict__ table, uint32_t n_rows, int64_t gathers_per_group, uint32_t seed,
                   float4* __restrict__ out) {
  constexpr int UU = 8;
  const int gl = threadIdx.x % G;
  const uint32_t group = (blockIdx.x * 256u + threadIdx.x) / G;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t ctr = seed + group * 0x9E3779B9u;
  for (int64_t it = 0; it < gathers_per_group; it += UU) {
    float4 v[UU];
#pragma unroll
    for (int u = 0; u < UU; ++u) {
      const uint32_t r = mb_hash(ctr + (uint32_t)(it + u)) % n_rows;     // same id in all lanes of the group
      v[u] = __ldg(table + (size_t)r * G + gl);
    }
#pragma unroll
    for (int u = 0; u < UU; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  }
  if (gl == 0) out[group] = acc;
}

}  // namespace rgbmp

using namespace rgbmp;

extern "C" {

int rgbmp_microbench_gather(const void* table, int64_t n_rows, int row_bytes, int64_t gathers_per_group, uint32_t seed,
                            void* out, int64_t n_groups, int device, void* stream) {
  if (!table || !out || n_rows <= 0 || n_rows >= (1ll << 32) || gathers_per_group <= 0 || n_groups <= 0)
    return fail(RGBMP_EINVAL, "rgbmp_microbench_gather: bad argument");
  const int G = row_bytes / 16;
  if (row_bytes % 16 != 0 || (G != 4 && G != 8 && G != 16 && G != 32))
    return fail(RGBMP_ERANGE, "rgbmp_microbench_gather: row_bytes must be 64, 128, 256 or 512");
  if ((n_groups * G) % 256 != 0) return fail(RGBMP_EINVAL, "rgbmp_microbench_gather: n_groups * lanes must fill whole CTAs");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RGBMP_EINVAL, "rgbmp_microbench_gather: bad device");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)(n_groups * G / 256);
  const float4* t = (const float4*)table;
  float4* o = (float4*)out;
  switch (G) {
    case 4: gather_peak_kernel<4><<<grid, 256, 0, st>>>(t, (uint32_t)n_rows, gathers_per_group, seed, o); break;
    case 8: gather_peak_kernel<8><<<grid, 256, 0, st>>>(t, (uint32_t)n_rows, gathers_per_group, seed, o); break;
    case 16: gather_peak_kernel<16><<<grid, 256, 0, st>>>(t, (uint32_t)n_rows, gathers_per_group, seed, o); break;
    default: gather_peak_kernel<32><<<grid, 256, 0, st>>>(t, (uint32_t)n_rows, gathers_per_group, seed, o); break;
  }
  RGBMP_LAUNCH_CHECK("gather_peak_kernel");
  return 0;
}

}  // extern "C"
