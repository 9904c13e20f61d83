// CSR SpMM kernels with fused epilogue (sm_100a) -- templates shared by the per-dtype
// instantiation units spmm_inst_*.cu.
//
// HBM-bound gather/segment-reduce (0.25-0.5 flop/B): no tensor cores.  Layout of the work:
//   * a GROUP of G lanes owns one output row; lane l of the group owns V 16-byte vectors of the
//     row (vector index l + v*G), so every gathered feature row is read with fully coalesced
//     16-byte loads and accumulated in registers with no cross-lane traffic;
//   * U edges are kept in flight per group (U*V independent 16-byte loads per lane) and the next
//     U column indices are prefetched while the current features are consumed;
//   * edges of a row are accumulated sequentially in CSR (= stable edge) order, with separate
//     multiply and add (no FMA contraction), which reproduces PyG-CPU scatter_add_ bit for bit;
//   * rows longer than `chunk` edges are split into CTA-sized work items (spmm_long_kernel) whose
//     partial sums are combined in a fixed order (spmm_combine_kernel) -- deterministic, no atomics;
//   * the epilogue (row scale, teleport, clamp, reset rows, pre-scaled second output) is applied
//     in registers before the single store of the row.
#pragma once
#include "common.cuh"

namespace rgbmp {

struct SpmmParams {
  // graph
  const int64_t* rowptr;
  const int32_t* col;
  const float* val;
  int64_t n_rows;
  int32_t chunk;
  // long rows
  int32_t long_chunk;
  const int32_t* long_rows;
  const int32_t* long_item_ptr;
  const int32_t* item_long;
  const int64_t* item_start;
  int64_t n_long, n_items;
  float* partial;  // [n_items, ldpart]
  int64_t ldpart;
  // features
  const void* X;
  int64_t ldx;
  void* Y;
  int64_t ldy;
  int F;
  // epilogue
  rgbmp_epilogue_t ep;
};

// ------------------------------------------------------------------------------------------
// vector load / store helpers.  EPV = elements per 16-byte vector (4 fp32, 8 bf16) or 1 (scalar).
// ------------------------------------------------------------------------------------------
template <typename T, int EPV>
struct Raw;
template <>
struct Raw<float, 4> {
  float4 r;
  __device__ __forceinline__ void load(const float* p) { r = __ldg(reinterpret_cast<const float4*>(p)); }
  __device__ __forceinline__ void zero() { r = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void unpack(float (&f)[4]) const { f[0] = r.x; f[1] = r.y; f[2] = r.z; f[3] = r.w; }
  static __device__ __forceinline__ void store(float* p, const float (&f)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};
template <>
struct Raw<float, 1> {
  float r;
  __device__ __forceinline__ void load(const float* p) { r = __ldg(p); }
  __device__ __forceinline__ void zero() { r = 0.f; }
  __device__ __forceinline__ void unpack(float (&f)[1]) const { f[0] = r; }
  static __device__ __forceinline__ void store(float* p, const float (&f)[1]) { *p = f[0]; }
};
template <>
struct Raw<__nv_bfloat16, 8> {
  uint4 r;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { r = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void zero() { r = make_uint4(0, 0, 0, 0); }
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <>
struct Raw<__nv_bfloat16, 1> {
  __nv_bfloat16 r;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { r = *p; }
  __device__ __forceinline__ void zero() { r = __float2bfloat16(0.f); }
  __device__ __forceinline__ void unpack(float (&f)[1]) const { f[0] = __bfloat162float(r); }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&f)[1]) { *p = __float2bfloat16(f[0]); }
};

__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }

// ------------------------------------------------------------------------------------------
// accumulate edges [k0, k1) of one row into acc (sequential, stable order)
// ------------------------------------------------------------------------------------------
template <typename T, int EPV, int G, int V, int U>
__device__ __forceinline__ void accumulate_range(const SpmmParams& p, int64_t k0, int64_t k1, int f_lane0,
                                                 const bool (&active)[V], float (&acc)[V][EPV]) {
  const T* __restrict__ X = reinterpret_cast<const T*>(p.X);
  const int32_t* __restrict__ col = p.col;
  const float* __restrict__ val = p.val;
  int32_t c[U];
  float w[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    c[u] = (k0 + u < k1) ? __ldg(col + k0 + u) : -1;
    w[u] = (val != nullptr && k0 + u < k1) ? __ldg(val + k0 + u) : 1.0f;
  }
  for (int64_t k = k0; k < k1; k += U) {
    Raw<T, EPV> raw[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const T* rowp = X + (int64_t)c[u] * p.ldx + f_lane0;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        if (c[u] >= 0 && active[v]) raw[u][v].load(rowp + v * G * EPV);
        else raw[u][v].zero();
      }
    }
    // prefetch the next U column ids / weights while the feature loads are in flight
    int32_t cn[U];
    float wn[U];
    const int64_t kn = k + U;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      cn[u] = (kn + u < k1) ? __ldg(col + kn + u) : -1;
      wn[u] = (val != nullptr && kn + u < k1) ? __ldg(val + kn + u) : 1.0f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (c[u] >= 0) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          float f[EPV];
          raw[u][v].unpack(f);
#pragma unroll
          for (int i = 0; i < EPV; ++i) {
            const float m = (val != nullptr) ? __fmul_rn(w[u], f[i]) : f[i];
            acc[v][i] = __fadd_rn(acc[v][i], m);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      c[u] = cn[u];
      w[u] = wn[u];
    }
  }
}

// epilogue for EPV consecutive features [f, f+EPV) of row `row`
template <typename T, int EPV>
__device__ __forceinline__ void epilogue_store(const SpmmParams& p, int64_t row, int f, float (&s)[EPV]) {
  const rgbmp_epilogue_t& ep = p.ep;
  const float rs = ep.row_scale ? __ldg(ep.row_scale + row) : 1.0f;
  const bool reset = (ep.reset_when != 0) && ep.reset_mask[row];
  float rv[EPV];
  if (reset) {
#pragma unroll
    for (int i = 0; i < EPV; ++i) rv[i] = (f + i < p.F) ? ep.reset_val[row * ep.ld_reset + f + i] : 0.f;
  }
  float t[EPV];
  if (ep.T) {
    Raw<T, EPV> tr;
    tr.load(reinterpret_cast<const T*>(ep.T) + row * ep.ldt + f);
    tr.unpack(t);
  }
#pragma unroll
  for (int i = 0; i < EPV; ++i) {
    float v = ep.row_scale ? (ep.row_div ? __fdiv_rn(s[i], rs) : __fmul_rn(rs, s[i])) : s[i];
    if (reset && ep.reset_when == 1) v = rv[i];
    v = __fmul_rn(ep.a, v);
    if (ep.T) v = __fadd_rn(v, __fmul_rn(ep.b, t[i]));
    if (ep.clamp) v = fminf(fmaxf(v, ep.lo), ep.hi);
    if (reset && ep.reset_when == 2) v = rv[i];
    s[i] = v;
  }
  if (p.Y) Raw<T, EPV>::store(reinterpret_cast<T*>(p.Y) + row * p.ldy + f, s);
  if (ep.Y2) {
    const float s2 = __ldg(ep.out2_scale + row);
    float o[EPV];
#pragma unroll
    for (int i = 0; i < EPV; ++i) o[i] = __fmul_rn(s2, s[i]);
    Raw<T, EPV>::store(reinterpret_cast<T*>(ep.Y2) + row * ep.ldy2 + f, o);
  }
}

constexpr int SPMM_THREADS = 256;

// short rows: one group of G lanes per row.  blockIdx.y = feature tile of G*V*EPV elements.
template <typename T, int EPV, int G, int V, int U>
__global__ void __launch_bounds__(SPMM_THREADS) spmm_rows_kernel(const SpmmParams p) {
  constexpr int GPB = SPMM_THREADS / G;
  const int gl = threadIdx.x % G;
  const int64_t row = (int64_t)blockIdx.x * GPB + threadIdx.x / G;
  if (row >= p.n_rows) return;
  const int64_t k0 = __ldg(p.rowptr + row), k1 = __ldg(p.rowptr + row + 1);
  if (p.chunk > 0 && k1 - k0 > p.chunk) return;  // long row: handled by spmm_long_kernel
  const int f0 = blockIdx.y * (G * V * EPV) + gl * EPV;
  bool active[V];
  float acc[V][EPV];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    active[v] = (f0 + v * G * EPV) < p.F;
#pragma unroll
    for (int i = 0; i < EPV; ++i) acc[v][i] = 0.f;
  }
  accumulate_range<T, EPV, G, V, U>(p, k0, k1, f0, active, acc);
#pragma unroll
  for (int v = 0; v < V; ++v)
    if (active[v]) epilogue_store<T, EPV>(p, row, f0 + v * G * EPV, acc[v]);
}

// long rows: one CTA per work item (<= long_chunk edges of one row); the CTA's groups take
// contiguous sub-ranges, partial sums are reduced through shared memory in a fixed order.
template <typename T, int EPV, int G, int V, int U>
__global__ void __launch_bounds__(SPMM_THREADS) spmm_long_kernel(const SpmmParams p) {
  constexpr int Q = SPMM_THREADS / G;      // groups per CTA
  constexpr int W = G * V * EPV;           // feature tile width
  extern __shared__ float sm[];            // [Q][W]
  const int gl = threadIdx.x % G, q = threadIdx.x / G;
  const int64_t item = blockIdx.x;
  const int32_t slot = p.item_long[item];
  const int64_t row = p.long_rows[slot];
  const int64_t rs = p.item_start[item];
  const int64_t rend = __ldg(p.rowptr + row + 1);
  const int64_t re = (rs + p.long_chunk < rend) ? rs + p.long_chunk : rend;
  const int64_t len = re - rs;
  const int64_t per = (len + Q - 1) / Q;
  int64_t k0 = rs + (int64_t)q * per, k1 = k0 + per;
  if (k0 > re) k0 = re;
  if (k1 > re) k1 = re;
  const int ftile = blockIdx.y * W;
  const int f0 = ftile + gl * EPV;
  bool active[V];
  float acc[V][EPV];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    active[v] = (f0 + v * G * EPV) < p.F;
#pragma unroll
    for (int i = 0; i < EPV; ++i) acc[v][i] = 0.f;
  }
  accumulate_range<T, EPV, G, V, U>(p, k0, k1, f0, active, acc);
#pragma unroll
  for (int v = 0; v < V; ++v)
#pragma unroll
    for (int i = 0; i < EPV; ++i) sm[q * W + (gl + v * G) * EPV + i] = acc[v][i];
  __syncthreads();
  for (int t = threadIdx.x; t < W; t += SPMM_THREADS) {
    float s = 0.f;
    for (int qq = 0; qq < Q; ++qq) s = __fadd_rn(s, sm[qq * W + t]);
    if (ftile + t < p.ldpart) p.partial[item * p.ldpart + ftile + t] = s;
  }
}

// combine the items of each long row in order and apply the epilogue (one thread per element)
template <typename T>
__global__ void __launch_bounds__(256) spmm_combine_kernel(const SpmmParams p) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t slot = t / p.F;
  const int f = (int)(t - slot * p.F);
  if (slot >= p.n_long) return;
  const int64_t row = p.long_rows[slot];
  float s[1] = {0.f};
  for (int32_t it = p.long_item_ptr[slot]; it < p.long_item_ptr[slot + 1]; ++it)
    s[0] = __fadd_rn(s[0], p.partial[(int64_t)it * p.ldpart + f]);
  epilogue_store<T, 1>(p, row, f, s);
}

template <typename T>
__global__ void __launch_bounds__(256)
row_scale_kernel(const T* __restrict__ X, int64_t ldx, const float* __restrict__ scale, int divide, T* __restrict__ Y,
                 int64_t ldy, int64_t n_rows, int F) {
  const int64_t total = n_rows * F;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t r = t / F;
    const int f = (int)(t - r * F);
    const float xv = to_f(X[r * ldx + f]);
    float o[1] = {divide ? __fdiv_rn(xv, scale[r]) : __fmul_rn(scale[r], xv)};
    Raw<T, 1>::store(Y + r * ldy + f, o);
  }
}

// ------------------------------------------------------------------------------------------
// launch dispatch
// ------------------------------------------------------------------------------------------
template <typename T, int EPV, int G, int V, int U>
int launch_cfg(const SpmmParams& p, cudaStream_t st) {
  constexpr int W = G * V * EPV;
  const unsigned ytiles = (unsigned)ceil_div(p.F, W);
  constexpr int GPB = SPMM_THREADS / G;
  if (p.n_rows > 0) {
    dim3 grid((unsigned)ceil_div(p.n_rows, GPB), ytiles);
    spmm_rows_kernel<T, EPV, G, V, U><<<grid, SPMM_THREADS, 0, st>>>(p);
    RGBMP_LAUNCH_CHECK("spmm_rows_kernel");
  }
  if (p.n_items > 0) {
    constexpr int Q = SPMM_THREADS / G;
    dim3 grid((unsigned)p.n_items, ytiles);
    spmm_long_kernel<T, EPV, G, V, U><<<grid, SPMM_THREADS, Q * W * sizeof(float), st>>>(p);
    RGBMP_LAUNCH_CHECK("spmm_long_kernel");
    spmm_combine_kernel<T><<<(unsigned)ceil_div(p.n_long * p.F, 256), 256, 0, st>>>(p);
    RGBMP_LAUNCH_CHECK("spmm_combine_kernel");
  }
  return 0;
}

template <typename T, int EPV, int G, int V>
int dispatch_u(const SpmmParams& p, int U, cudaStream_t st) {
  switch (U) {
    case 2: return launch_cfg<T, EPV, G, V, 2>(p, st);
    case 8: return launch_cfg<T, EPV, G, V, 8>(p, st);
    default: return launch_cfg<T, EPV, G, V, 4>(p, st);
  }
}

template <typename T, int EPV, int G>
int dispatch_v(const SpmmParams& p, int V, int U, cudaStream_t st) {
  switch (V) {
    case 1: return dispatch_u<T, EPV, G, 1>(p, U, st);
    case 2: return dispatch_u<T, EPV, G, 2>(p, U, st);
    case 3: return dispatch_u<T, EPV, G, 3>(p, U, st);
    default: return dispatch_u<T, EPV, G, 4>(p, U, st);
  }
}

template <typename T, int EPV>
int dispatch_g(const SpmmParams& p, int G, int V, int U, cudaStream_t st) {
  switch (G) {
    case 1: return dispatch_v<T, EPV, 1>(p, V, U, st);
    case 2: return dispatch_v<T, EPV, 2>(p, V, U, st);
    case 4: return dispatch_v<T, EPV, 4>(p, V, U, st);
    case 8: return dispatch_v<T, EPV, 8>(p, V, U, st);
    case 16: return dispatch_v<T, EPV, 16>(p, V, U, st);
    default: return dispatch_v<T, EPV, 32>(p, V, U, st);
  }
}

// per-dtype instantiation units
int spmm_dispatch_f32v(const SpmmParams& p, int G, int V, int U, cudaStream_t st);
int spmm_dispatch_f32s(const SpmmParams& p, int G, int V, int U, cudaStream_t st);
int spmm_dispatch_bf16(const SpmmParams& p, int G, int V, int U, cudaStream_t st);

}  // namespace rgbmp
