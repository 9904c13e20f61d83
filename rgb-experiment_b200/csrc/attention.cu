// Attention kernels (sm_100a): fused GATConv edge-softmax + aggregate (forward / backward),
// SDDMM, u_add_v, segmented softmax (+backward), segmented sum, multi-head weighted SpMM.
//
// All are gather / segment-reduce kernels over CSR -- HBM/L2-bound, no tensor cores.  A row of the
// CSR (all in-edges of one target, or all out-edges of one source for the transpose) is owned by a
// group of G lanes (fused GAT) or a warp (edge-score kernels); reductions over a row's edges are
// sequential per lane + warp shuffles, so results are deterministic.  The only atomics are the
// da_dst accumulation of the GAT backward (one float per edge and head).
#include "common.cuh"

namespace rgbmp {

__device__ __forceinline__ float leaky(float x, float slope) { return x > 0.f ? x : x * slope; }

// ------------------------------------------------------------------------------------------
// fused GAT forward.  G lanes per target row, lane l owns float4 #l of the H*C row
// (requires H == 1 or C % 4 == 0 so that a float4 never straddles two heads).
// Pass 1: per-head max of e = leaky(a_src[j] + a_dst[i]).  Pass 2: p = exp(e - max), s += p,
// acc += p * Xp[j].  out = acc / (s + 1e-16)  (softmax denominator applied once per row).
// ------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(256)
gat_fwd_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_rows,
               const float* __restrict__ Xp, int64_t ldx, const float* __restrict__ a_src,
               const float* __restrict__ a_dst, int H, int C, float slope, const float* __restrict__ drop,
               float* __restrict__ out, int64_t ldo, float* __restrict__ rowmax, float* __restrict__ rowsum) {
  constexpr int GPB = 256 / G;
  const int gl = threadIdx.x % G;
  const int64_t row = (int64_t)blockIdx.x * GPB + threadIdx.x / G;
  if (row >= n_rows) return;
  const int HC = H * C;
  const int f = gl * 4;
  const bool active = f < HC;
  const int h = active ? (H == 1 ? 0 : f / C) : 0;
  const int64_t k0 = __ldg(rowptr + row), k1 = __ldg(rowptr + row + 1);
  const float ad = a_dst[row * H + h];
  // pass 1: max
  float m = -INFINITY;
  for (int64_t k = k0; k < k1; ++k) {
    const int32_t j = __ldg(col + k);
    m = fmaxf(m, leaky(__ldg(a_src + (int64_t)j * H + h) + ad, slope));
  }
  // pass 2
  float s = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr int U = 4;
  for (int64_t k = k0; k < k1; k += U) {
    int32_t j[U];
    float p[U];
    float4 x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) j[u] = (k + u < k1) ? __ldg(col + k + u) : -1;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      p[u] = 0.f;
      if (j[u] >= 0) {
        if (active) x[u] = __ldg(reinterpret_cast<const float4*>(Xp + (int64_t)j[u] * ldx + f));
        p[u] = __ldg(a_src + (int64_t)j[u] * H + h);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (j[u] >= 0) {
        const float e = __expf(leaky(p[u] + ad, slope) - m);
        s += e;
        const float w = drop ? e * __ldg(drop + (k + u) * H + h) : e;
        acc.x += w * x[u].x;
        acc.y += w * x[u].y;
        acc.z += w * x[u].z;
        acc.w += w * x[u].w;
      }
    }
  }
  const float inv = (k1 > k0) ? 1.0f / (s + 1e-16f) : 0.f;
  if (active) {
    *reinterpret_cast<float4*>(out + row * ldo + f) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
    // one lane per head records the softmax statistics
    if (H == 1 ? (gl == 0) : (f % C == 0)) {
      rowmax[row * H + h] = (k1 > k0) ? m : 0.f;
      rowsum[row * H + h] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------
// fused GAT backward on the transpose CSR (rows = sources j, col = targets i).
//   alpha_ij = exp(leaky(a_s[j]+a_d[i]) - max_i) / (sum_i + 1e-16)
//   dXp[j]   = sum_i (alpha_ij*drop) * dout[i]
//   dalpha   = drop * <dout[i,h,:], Xp[j,h,:]> ;  de = alpha * (dalpha - S[i,h]) ; dlogit = de * leaky'
//   da_src[j,h] = sum_i dlogit  (row-local) ;  da_dst[i,h] += dlogit (atomic)
// LPH = lanes per head (power of two); the per-head dot product is a shuffle reduction.
// ------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(256)
gat_bwd_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_rows,
               const float* __restrict__ Xp, int64_t ldx, const float* __restrict__ a_src,
               const float* __restrict__ a_dst, int H, int C, int LPH, float slope,
               const float* __restrict__ drop, const int32_t* __restrict__ tpos,
               const float* __restrict__ rowmax, const float* __restrict__ rowsum, const float* __restrict__ S,
               const float* __restrict__ dout, int64_t ldd, float* __restrict__ dXp, int64_t lddx,
               float* __restrict__ da_src, float* __restrict__ da_dst) {
  constexpr int GPB = 256 / G;
  const int gl = threadIdx.x % G;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * GPB + threadIdx.x / G;
  if (row >= n_rows) return;   // whole groups exit together; shuffles below use the group mask
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane / G * G));
  const int HC = H * C;
  const int f = gl * 4;
  const bool active = f < HC;
  const int h = active ? (H == 1 ? 0 : f / C) : 0;
  const int64_t k0 = __ldg(rowptr + row), k1 = __ldg(rowptr + row + 1);
  const float as = a_src[row * H + h];
  float4 xj = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) xj = __ldg(reinterpret_cast<const float4*>(Xp + row * ldx + f));
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float das = 0.f;
  for (int64_t k = k0; k < k1; ++k) {
    const int64_t i = __ldg(col + k);
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) d = __ldg(reinterpret_cast<const float4*>(dout + i * ldd + f));
    const float raw = as + __ldg(a_dst + i * H + h);
    const float alpha = __expf(leaky(raw, slope) - __ldg(rowmax + i * H + h)) / (__ldg(rowsum + i * H + h) + 1e-16f);
    const float dr = drop ? __ldg(drop + (int64_t)__ldg(tpos + k) * H + h) : 1.0f;
    float dot = d.x * xj.x + d.y * xj.y + d.z * xj.z + d.w * xj.w;
    for (int o = LPH >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(gmask, dot, o, G);
    const float w = alpha * dr;
    acc.x += w * d.x;
    acc.y += w * d.y;
    acc.z += w * d.z;
    acc.w += w * d.w;
    const float de = alpha * (dr * dot - __ldg(S + i * H + h));
    const float dl = de * (raw > 0.f ? 1.0f : slope);
    if (active && (gl % LPH) == 0) {
      das += dl;
      atomicAdd(da_dst + i * H + h, dl);
    }
  }
  if (active) {
    *reinterpret_cast<float4*>(dXp + row * lddx + f) = acc;
    if ((gl % LPH) == 0) da_src[row * H + h] = das;
  }
}

__global__ void __launch_bounds__(256)
rowdot_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb, int64_t n_rows, int H,
              int C, float* __restrict__ S) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_rows * H) return;
  const int64_t i = t / H;
  const int h = (int)(t - i * H);
  const float* a = A + i * lda + h * C;
  const float* b = B + i * ldb + h * C;
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += a[c] * b[c];
  S[t] = s;
}

// ------------------------------------------------------------------------------------------
// warp-per-row edge-score kernels
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sddmm_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_rows,
             const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb, int H, int C,
             float* __restrict__ out) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const int64_t k0 = rowptr[row], k1 = rowptr[row + 1];
  const int64_t total = (k1 - k0) * H;
  const float* a_row = A + row * lda;
  for (int64_t idx = lane; idx < total; idx += 32) {
    const int64_t k = k0 + idx / H;
    const int h = (int)(idx % H);
    const float* a = a_row + h * C;
    const float* b = B + (int64_t)col[k] * ldb + h * C;
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += __ldg(a + c) * __ldg(b + c);
    out[k * H + h] = s;
  }
}

__global__ void __launch_bounds__(256)
u_add_v_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_rows,
               const float* __restrict__ u, const float* __restrict__ v, int H, float* __restrict__ out) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const int64_t k0 = rowptr[row], k1 = rowptr[row + 1];
  const int64_t total = (k1 - k0) * H;
  for (int64_t idx = lane; idx < total; idx += 32) {
    const int64_t k = k0 + idx / H;
    const int h = (int)(idx % H);
    out[k * H + h] = __ldg(u + (int64_t)col[k] * H + h) + __ldg(v + row * H + h);
  }
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// out[k,h] = exp(in[k,h]-max_h) / (sum_h + 1e-16)   
__global__ void __launch_bounds__(256)
seg_softmax_kernel(const int64_t* __restrict__ rowptr, int64_t n_rows, const float* __restrict__ in, int H,
                   float* __restrict__ out) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const int64_t k0 = rowptr[row], k1 = rowptr[row + 1];
  for (int h = 0; h < H; ++h) {
    float m = -INFINITY;
    for (int64_t k = k0 + lane; k < k1; k += 32) m = fmaxf(m, in[k * H + h]);
    m = warp_max(m);
    float s = 0.f;
    for (int64_t k = k0 + lane; k < k1; k += 32) s += expf(in[k * H + h] - m);
    s = warp_sum(s);
    const float den = s + 1e-16f;
    for (int64_t k = k0 + lane; k < k1; k += 32) out[k * H + h] = expf(in[k * H + h] - m) / den;
  }
}

// dlogit[k,h] = alpha[k,h] * (dalpha[k,h] - sum_k' alpha[k',h]*dalpha[k',h])
__global__ void __launch_bounds__(256)
seg_softmax_bwd_kernel(const int64_t* __restrict__ rowptr, int64_t n_rows, const float* __restrict__ alpha,
                       const float* __restrict__ dalpha, int H, float* __restrict__ dlogit) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const int64_t k0 = rowptr[row], k1 = rowptr[row + 1];
  for (int h = 0; h < H; ++h) {
    float s = 0.f;
    for (int64_t k = k0 + lane; k < k1; k += 32) s += alpha[k * H + h] * dalpha[k * H + h];
    s = warp_sum(s);
    for (int64_t k = k0 + lane; k < k1; k += 32) dlogit[k * H + h] = alpha[k * H + h] * (dalpha[k * H + h] - s);
  }
}

__global__ void __launch_bounds__(256)
seg_sum_kernel(const int64_t* __restrict__ rowptr, int64_t n_rows, const float* __restrict__ in, int H,
               float* __restrict__ out) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const int64_t k0 = rowptr[row], k1 = rowptr[row + 1];
  for (int h = 0; h < H; ++h) {
    float s = 0.f;
    for (int64_t k = k0 + lane; k < k1; k += 32) s += in[k * H + h];
    s = warp_sum(s);
    if (lane == 0) out[row * H + h] = s;
  }
}

// out[i,h,:] = sum_k w[k,h] * X[col[k],h,:]   warp per row, lanes over the H*C features
__global__ void __launch_bounds__(256)
spmm_heads_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_rows,
                  const float* __restrict__ w, const float* __restrict__ X, int64_t ldx, int H, int C,
                  float* __restrict__ out, int64_t ldo) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const int64_t k0 = rowptr[row], k1 = rowptr[row + 1];
  const int HC = H * C;
  for (int fb = 0; fb < HC; fb += 32) {
    const int f = fb + lane;
    const bool act = f < HC;
    const int h = act ? f / C : 0;
    float acc = 0.f;
    for (int64_t k = k0; k < k1; ++k) {
      const int64_t j = col[k];
      if (act) acc = __fadd_rn(acc, __fmul_rn(__ldg(w + k * H + h), __ldg(X + j * ldx + f)));
    }
    if (act) out[row * ldo + f] = acc;
  }
}

static int check_g(const rgbmp_graph_t* g, const char* fn) {
  if (!g || !g->rowptr || g->n_rows < 0 || g->nnz < 0 || (g->nnz > 0 && !g->col))
    return fail(RGBMP_EINVAL, "%s: bad graph descriptor", fn);
  return 0;
}

static inline bool al16(const void* p, int64_t ld) { return (((uintptr_t)p) & 15) == 0 && (ld % 4) == 0; }

}  // namespace rgbmp

using namespace rgbmp;

extern "C" {

size_t rgbmp_gat_workspace_bytes(const rgbmp_graph_t* g, int H, int C) {
  (void)g; (void)H; (void)C;
  return 256;
}

#define GAT_DISPATCH(KERNEL, ...)                                                              \
  switch (G) {                                                                                 \
    case 1: KERNEL<1><<<(unsigned)ceil_div(n, 256 / 1), 256, 0, st>>>(__VA_ARGS__); break;     \
    case 2: KERNEL<2><<<(unsigned)ceil_div(n, 256 / 2), 256, 0, st>>>(__VA_ARGS__); break;     \
    case 4: KERNEL<4><<<(unsigned)ceil_div(n, 256 / 4), 256, 0, st>>>(__VA_ARGS__); break;     \
    case 8: KERNEL<8><<<(unsigned)ceil_div(n, 256 / 8), 256, 0, st>>>(__VA_ARGS__); break;     \
    case 16: KERNEL<16><<<(unsigned)ceil_div(n, 256 / 16), 256, 0, st>>>(__VA_ARGS__); break;  \
    default: KERNEL<32><<<(unsigned)ceil_div(n, 256 / 32), 256, 0, st>>>(__VA_ARGS__); break;  \
  }

static int gat_group(int HC) {
  const int nvec = (HC + 3) / 4;
  int G = 1;
  while (G < nvec) G <<= 1;
  return G;
}

int rgbmp_gat_forward(const rgbmp_graph_t* g, const float* Xp, int64_t ldx, const float* a_src, const float* a_dst,
                      int H, int C, float slope, const float* drop, float* out, int64_t ldo, float* rowmax,
                      float* rowsum, void* ws, size_t ws_bytes, int device, void* stream) {
  (void)ws; (void)ws_bytes;
  int rc = check_g(g, "rgbmp_gat_forward");
  if (rc) return rc;
  if (!Xp || !a_src || !a_dst || !out || !rowmax || !rowsum || H <= 0 || C <= 0)
    return fail(RGBMP_EINVAL, "rgbmp_gat_forward: null pointer / bad H,C");
  const int HC = H * C;
  if (HC > 128) return fail(RGBMP_ERANGE, "rgbmp_gat_forward: H*C = %d > 128 (use the unfused kernels)", HC);
  if (H != 1 && (C % 4) != 0) return fail(RGBMP_ERANGE, "rgbmp_gat_forward: needs H == 1 or C %% 4 == 0");
  if (!al16(Xp, ldx) || !al16(out, ldo) || ldx < (int64_t)align_up(HC, 4) || ldo < (int64_t)align_up(HC, 4))
    return fail(RGBMP_EALIGN, "rgbmp_gat_forward: Xp/out need 16-byte aligned rows with ld >= roundup(H*C,4)");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RGBMP_EINVAL, "rgbmp_gat_forward: bad device");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = g->n_rows;
  if (n == 0) return 0;
  const int G = gat_group(HC);
  GAT_DISPATCH(gat_fwd_kernel, g->rowptr, g->col, n, Xp, ldx, a_src, a_dst, H, C, slope, drop, out, ldo, rowmax, rowsum)
  RGBMP_LAUNCH_CHECK("gat_fwd_kernel");
  return 0;
}

int rgbmp_gat_backward(const rgbmp_graph_t* gT, const float* Xp, int64_t ldx, const float* a_src, const float* a_dst,
                       int H, int C, float slope, const float* drop, const int32_t* tpos, const float* rowmax,
                       const float* rowsum, const float* S, const float* dout, int64_t ldd, float* dXp, int64_t lddx,
                       float* da_src, float* da_dst, int device, void* stream) {
  int rc = check_g(gT, "rgbmp_gat_backward");
  if (rc) return rc;
  if (!Xp || !a_src || !a_dst || !rowmax || !rowsum || !S || !dout || !dXp || !da_src || !da_dst || H <= 0 || C <= 0 ||
      (drop && !tpos))
    return fail(RGBMP_EINVAL, "rgbmp_gat_backward: null pointer / bad H,C");
  const int HC = H * C;
  if (HC > 128) return fail(RGBMP_ERANGE, "rgbmp_gat_backward: H*C = %d > 128", HC);
  const int G = gat_group(HC);
  int LPH = G;
  if (H != 1) {
    if ((C % 4) != 0 || ((C / 4) & (C / 4 - 1)) != 0)
      return fail(RGBMP_ERANGE, "rgbmp_gat_backward: needs H == 1 or C in {4,8,16,32,64,128}");
    LPH = C / 4;
  }
  if (!al16(Xp, ldx) || !al16(dout, ldd) || !al16(dXp, lddx))
    return fail(RGBMP_EALIGN, "rgbmp_gat_backward: Xp/dout/dXp need 16-byte aligned rows");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RGBMP_EINVAL, "rgbmp_gat_backward: bad device");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = gT->n_rows;
  if (n == 0) return 0;
  GAT_DISPATCH(gat_bwd_kernel, gT->rowptr, gT->col, n, Xp, ldx, a_src, a_dst, H, C, LPH, slope, drop, tpos, rowmax,
               rowsum, S, dout, ldd, dXp, lddx, da_src, da_dst)
  RGBMP_LAUNCH_CHECK("gat_bwd_kernel");
  return 0;
}

int rgbmp_rowdot(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t n_rows, int H, int C, float* S,
                 int device, void* stream) {
  if (!A || !B || !S || n_rows < 0 || H <= 0 || C <= 0) return fail(RGBMP_EINVAL, "rgbmp_rowdot: bad argument");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RGBMP_EINVAL, "rgbmp_rowdot: bad device");
  if (n_rows == 0) return 0;
  rowdot_kernel<<<(unsigned)ceil_div(n_rows * H, 256), 256, 0, (cudaStream_t)stream>>>(A, lda, B, ldb, n_rows, H, C, S);
  RGBMP_LAUNCH_CHECK("rowdot_kernel");
  return 0;
}

#define WARP_ROW_GRID(n) (unsigned)ceil_div((n) * 32, 256), 256, 0, (cudaStream_t)stream

int rgbmp_sddmm(const rgbmp_graph_t* g, const float* A, int64_t lda, const float* B, int64_t ldb, int H, int C,
                float* out, int device, void* stream) {
  int rc = check_g(g, "rgbmp_sddmm");
  if (rc) return rc;
  if (!A || !B || (g->nnz > 0 && !out) || H <= 0 || C <= 0) return fail(RGBMP_EINVAL, "rgbmp_sddmm: bad argument");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RGBMP_EINVAL, "rgbmp_sddmm: bad device");
  if (g->n_rows == 0 || g->nnz == 0) return 0;
  sddmm_kernel<<<WARP_ROW_GRID(g->n_rows)>>>(g->rowptr, g->col, g->n_rows, A, lda, B, ldb, H, C, out);
  RGBMP_LAUNCH_CHECK("sddmm_kernel");
  return 0;
}

int rgbmp_u_add_v(const rgbmp_graph_t* g, const float* u, const float* v, int H, float* out, int device, void* stream) {
  int rc = check_g(g, "rgbmp_u_add_v");
  if (rc) return rc;
  if (!u || !v || (g->nnz > 0 && !out) || H <= 0) return fail(RGBMP_EINVAL, "rgbmp_u_add_v: bad argument");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RGBMP_EINVAL, "rgbmp_u_add_v: bad device");
  if (g->n_rows == 0 || g->nnz == 0) return 0;
  u_add_v_kernel<<<WARP_ROW_GRID(g->n_rows)>>>(g->rowptr, g->col, g->n_rows, u, v, H, out);
  RGBMP_LAUNCH_CHECK("u_add_v_kernel");
  return 0;
}

int rgbmp_seg_softmax(const rgbmp_graph_t* g, const float* in, int H, float* out, int device, void* stream) {
  int rc = check_g(g, "rgbmp_seg_softmax");
  if (rc) return rc;
  if ((g->nnz > 0 && (!in || !out)) || H <= 0) return fail(RGBMP_EINVAL, "rgbmp_seg_softmax: bad argument");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RGBMP_EINVAL, "rgbmp_seg_softmax: bad device");
  if (g->n_rows == 0 || g->nnz == 0) return 0;
  seg_softmax_kernel<<<WARP_ROW_GRID(g->n_rows)>>>(g->rowptr, g->n_rows, in, H, out);
  RGBMP_LAUNCH_CHECK("seg_softmax_kernel");
  return 0;
}

int rgbmp_seg_softmax_backward(const rgbmp_graph_t* g, const float* alpha, const float* dalpha, int H, float* dlogit,
                               int device, void* stream) {
  int rc = check_g(g, "rgbmp_seg_softmax_backward");
  if (rc) return rc;
  if ((g->nnz > 0 && (!alpha || !dalpha || !dlogit)) || H <= 0)
    return fail(RGBMP_EINVAL, "rgbmp_seg_softmax_backward: bad argument");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RGBMP_EINVAL, "rgbmp_seg_softmax_backward: bad device");
  if (g->n_rows == 0 || g->nnz == 0) return 0;
  seg_softmax_bwd_kernel<<<WARP_ROW_GRID(g->n_rows)>>>(g->rowptr, g->n_rows, alpha, dalpha, H, dlogit);
  RGBMP_LAUNCH_CHECK("seg_softmax_bwd_kernel");
  return 0;
}

int rgbmp_seg_sum(const rgbmp_graph_t* g, const float* in, int H, float* out, int device, void* stream) {
  int rc = check_g(g, "rgbmp_seg_sum");
  if (rc) return rc;
  if ((g->nnz > 0 && !in) || !out || H <= 0) return fail(RGBMP_EINVAL, "rgbmp_seg_sum: bad argument");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RGBMP_EINVAL, "rgbmp_seg_sum: bad device");
  if (g->n_rows == 0) return 0;
  seg_sum_kernel<<<WARP_ROW_GRID(g->n_rows)>>>(g->rowptr, g->n_rows, in, H, out);
  RGBMP_LAUNCH_CHECK("seg_sum_kernel");
  return 0;
}

int rgbmp_spmm_heads(const rgbmp_graph_t* g, const float* w, const float* X, int64_t ldx, int H, int C, float* out,
                     int64_t ldo, int device, void* stream) {
  int rc = check_g(g, "rgbmp_spmm_heads");
  if (rc) return rc;
  if ((g->nnz > 0 && !w) || !X || !out || H <= 0 || C <= 0) return fail(RGBMP_EINVAL, "rgbmp_spmm_heads: bad argument");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RGBMP_EINVAL, "rgbmp_spmm_heads: bad device");
  if (g->n_rows == 0) return 0;
  spmm_heads_kernel<<<WARP_ROW_GRID(g->n_rows)>>>(g->rowptr, g->col, g->n_rows, w, X, ldx, H, C, out, ldo);
  RGBMP_LAUNCH_CHECK("spmm_heads_kernel");
  return 0;
}

}  // extern "C"
