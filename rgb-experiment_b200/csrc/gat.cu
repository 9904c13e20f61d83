// Fused GATConv edge-softmax + aggregate, forward and backward (sm_100a), v2.
//
// Gather / segment-reduce over CSR: L2/HBM-bound, no tensor cores.  v1 gave one group of lanes a
// whole row, so the hubs of a Reddit-shaped graph (max in-degree ~5e5 at average 492) serialised
// on 16 lanes and the layer ran at 0.04 of the roofline.  v2 uses the SpMM machinery:
//   * the row maximum of the logits is a separate, lane-dense pre-pass (gat_max_*): every lane
//     handles one (edge, head) pair, so the softmax below is the exact two-pass form of
//     torch_geometric.utils.softmax (max, then exp(e - max));
//   * short rows: one group of G lanes per row in the degree-sorted row schedule, the G column ids
//     of a batch are loaded with one coalesced load and broadcast with shuffles, U edges in flight;
//   * rows longer than `chunk` edges are split into CTA-sized work items whose partial
//     (sum of p, sum of p*x) are combined in item order -- deterministic, no atomics;
//   * backward runs on the transpose CSR with the same split; the per-target statistics it needs
//     (a_dst, max, 1/(sum+1e-16), S = <dout_i, out_i>) are packed into one float4 per (node, head)
//     by gat_bwd_prep so that each edge costs one 16-byte gather per head instead of four.
// The only atomics are the da_dst accumulation of the backward (one float per edge and head).
#include "common.cuh"

namespace rgbmp {

constexpr unsigned FULLMASK = 0xffffffffu;
constexpr int GAT_THREADS = 256;

__device__ __forceinline__ float leaky_relu(float x, float slope) { return x > 0.f ? x : x * slope; }

struct GatParams {
  // graph (forward CSR for the forward pass, transpose CSR for the backward)
  const int64_t* rowptr;
  const int32_t* col;
  const int32_t* row_order;
  int64_t n_rows;
  int32_t chunk, long_chunk;
  const int32_t* long_rows;
  const int32_t* long_item_ptr;
  const int32_t* item_long;
  const int64_t* item_start;
  int64_t n_long, n_items;
  // operands
  const float* Xp;
  int64_t ldx;
  const float* a_src;
  const float* a_dst;
  int H, C, LPH;
  float slope;
  const float* drop;
  const int32_t* tpos;
  // forward outputs
  float* out;
  int64_t ldo;
  float* rowmax;
  float* rowsum;
  // backward operands / outputs
  const float4* stats;   // [n_dst, H]: (a_dst, rowmax, 1/(rowsum+1e-16), S)
  const float* dout;
  int64_t ldd;
  float* dXp;
  int64_t lddx;
  float* da_src;
  float* da_dst;
  // long-row scratch
  float* item_max;   // [n_items, H]
  float* part_acc;   // [n_items, ldp]
  float* part_s;     // [n_items, H]
  int64_t ldp;
};

// ------------------------------------------------------------------------------------------
// pre-pass: rowmax[i,h] = max over in-edges j of leaky(a_src[j,h] + a_dst[i,h])   (0 for empty rows)
// Lane = (edge slot, head): HP = pow2 >= H heads side by side, 32/HP edges per warp step.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float range_max(const GatParams& p, int64_t k0, int64_t k1, int64_t row, int lane, int HP) {
  const int hh = lane % HP, slot = lane / HP, EPW = 32 / HP;
  float m = -INFINITY;
  if (hh < p.H) {
    const float ad = __ldg(p.a_dst + row * p.H + hh);
    for (int64_t k = k0 + slot; k < k1; k += EPW) {
      const int64_t j = __ldcs(p.col + k);
      m = fmaxf(m, leaky_relu(__ldg(p.a_src + j * p.H + hh) + ad, p.slope));
    }
  }
  for (int o = HP; o < 32; o <<= 1) m = fmaxf(m, __shfl_xor_sync(FULLMASK, m, o));
  return m;   // valid in every lane with hh < H
}

__global__ void __launch_bounds__(GAT_THREADS) gat_max_rows_kernel(const GatParams p, int HP) {
  const int64_t row = ((int64_t)blockIdx.x * GAT_THREADS + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= p.n_rows) return;
  const int64_t k0 = __ldg(p.rowptr + row), k1 = __ldg(p.rowptr + row + 1);
  if (p.n_items > 0 && k1 - k0 > p.chunk) return;   // long row: gat_max_long_kernel
  const float m = range_max(p, k0, k1, row, lane, HP);
  if (lane < p.H) p.rowmax[row * p.H + lane] = (k1 > k0) ? m : 0.f;
}

__global__ void __launch_bounds__(GAT_THREADS) gat_max_long_kernel(const GatParams p, int HP) {
  __shared__ float sm[GAT_THREADS / 32][32];
  const int64_t item = blockIdx.x;
  const int64_t row = p.long_rows[p.item_long[item]];
  const int64_t rs = p.item_start[item];
  const int64_t rend = __ldg(p.rowptr + row + 1);
  const int64_t re = (rs + p.long_chunk < rend) ? rs + p.long_chunk : rend;
  constexpr int NW = GAT_THREADS / 32;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t per = (re - rs + NW - 1) / NW;
  int64_t k0 = rs + w * per, k1 = k0 + per;
  if (k0 > re) k0 = re;
  if (k1 > re) k1 = re;
  sm[w][lane] = range_max(p, k0, k1, row, lane, HP);
  __syncthreads();
  if (threadIdx.x < p.H) {
    float m = -INFINITY;
    for (int q = 0; q < NW; ++q) m = fmaxf(m, sm[q][threadIdx.x]);
    p.item_max[item * p.H + threadIdx.x] = m;
  }
}

// row maximum of a long row = max over its items (<= deg/long_chunk values, L2-resident)
__device__ __forceinline__ float long_row_max(const GatParams& p, int32_t slot, int h) {
  float m = -INFINITY;
  for (int32_t it = p.long_item_ptr[slot]; it < p.long_item_ptr[slot + 1]; ++it) m = fmaxf(m, p.item_max[(int64_t)it * p.H + h]);
  return m;
}

// ------------------------------------------------------------------------------------------
// forward accumulation of edges [k0,k1) of target `row`: s += p, acc += p*drop*Xp[j]  with
// p = exp(leaky(a_src[j,h] + ad) - M).  Warp-uniform trip counts (see spmm_kernels.cuh).
// ------------------------------------------------------------------------------------------
template <int G, int U>
__device__ __forceinline__ void gat_fwd_range(const GatParams& p, int64_t k0, int64_t k1, int gl, int f, int h, float ad,
                                              float M, float& s, float4& acc) {
  const int len = (k1 > k0) ? (int)(k1 - k0) : 0;
  const int maxlen = __reduce_max_sync(FULLMASK, len);
  if (maxlen == 0) return;
  const int32_t* __restrict__ col = p.col + k0;
  const char* xbase = reinterpret_cast<const char*>(p.Xp + f);
  const uint32_t row_bytes = (uint32_t)(p.ldx * 4);
  const float* asb = p.a_src + h;
  int32_t cl = (gl < len) ? __ldcs(col + gl) : 0;
  for (int off = 0; off < maxlen; off += G) {
    int nb = len - off;
    nb = nb < 0 ? 0 : (nb > G ? G : nb);
    int32_t cn = 0;
    if (off + G + gl < len) cn = __ldcs(col + off + G + gl);
    if (G >= 2 * U && __all_sync(FULLMASK, nb == G)) {
      // full batch in every group: software-pipelined, unpredicated -- the gathers of step j+U are in
      // flight while step j is consumed (same structure as the SpMM main loop)
      float4 x[U];
      float as[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t c = (uint32_t)__shfl_sync(FULLMASK, cl, u, G);
        x[u] = __ldg(reinterpret_cast<const float4*>(xbase + (size_t)c * row_bytes));
        as[u] = __ldg(asb + (size_t)c * p.H);
      }
#pragma unroll 1
      for (int j = 0; j < G; j += U) {
        float4 xn[U];
        float asn[U];
        const int jn = (j + U < G) ? j + U : j;      // last step re-requests itself (L1 hit, unused)
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint32_t c = (uint32_t)__shfl_sync(FULLMASK, cl, jn + u, G);
          xn[u] = __ldg(reinterpret_cast<const float4*>(xbase + (size_t)c * row_bytes));
          asn[u] = __ldg(asb + (size_t)c * p.H);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float e = __expf(leaky_relu(as[u] + ad, p.slope) - M);
          s += e;
          const float w = p.drop ? e * __ldg(p.drop + (k0 + off + j + u) * p.H + h) : e;
          acc.x += w * x[u].x;
          acc.y += w * x[u].y;
          acc.z += w * x[u].z;
          acc.w += w * x[u].w;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          x[u] = xn[u];
          as[u] = asn[u];
        }
      }
    } else {
      const int nbmax = __reduce_max_sync(FULLMASK, nb);
#pragma unroll 1
      for (int j = 0; j < nbmax; j += U) {
        float4 x[U];
        float as[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint32_t c = (uint32_t)__shfl_sync(FULLMASK, cl, j + u, G);   // lanes past nb hold id 0: a valid row
          x[u] = __ldg(reinterpret_cast<const float4*>(xbase + (size_t)c * row_bytes));
          as[u] = __ldg(asb + (size_t)c * p.H);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (j + u < nb) {
            const float e = __expf(leaky_relu(as[u] + ad, p.slope) - M);
            s += e;
            const float w = p.drop ? e * __ldg(p.drop + (k0 + off + j + u) * p.H + h) : e;
            acc.x += w * x[u].x;
            acc.y += w * x[u].y;
            acc.z += w * x[u].z;
            acc.w += w * x[u].w;
          }
        }
      }
    }
    cl = cn;
  }
}

// lane -> (feature offset, head); inactive lanes (f >= H*C) shadow vector 0 and never store
__device__ __forceinline__ void lane_slot(const GatParams& p, int gl, int& f, int& h, bool& active) {
  f = gl * 4;
  active = f < p.H * p.C;
  if (!active) f = 0;
  h = (p.H == 1) ? 0 : f / p.C;
}

template <int G, int U>
__global__ void __launch_bounds__(GAT_THREADS) gat_fwd_rows_kernel(const GatParams p) {
  constexpr int GPB = GAT_THREADS / G;
  const int gl = threadIdx.x % G;
  const int64_t gid = (int64_t)blockIdx.x * GPB + threadIdx.x / G;
  int64_t row = -1, k0 = 0, k1 = 0;
  if (gid < p.n_rows) {
    row = p.row_order ? (int64_t)__ldg(p.row_order + gid) : gid;
    k0 = __ldg(p.rowptr + row);
    k1 = __ldg(p.rowptr + row + 1);
    if (p.n_items > 0 && k1 - k0 > p.chunk) row = -1;
  }
  if (row < 0) k1 = k0;
  int f, h;
  bool active;
  lane_slot(p, gl, f, h, active);
  float ad = 0.f, M = 0.f;
  if (row >= 0) {
    ad = __ldg(p.a_dst + row * p.H + h);
    M = p.rowmax[row * p.H + h];
  }
  float s = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  gat_fwd_range<G, U>(p, k0, k1, gl, f, h, ad, M, s, acc);
  if (row < 0 || !active) return;
  const float inv = (k1 > k0) ? 1.0f / (s + 1e-16f) : 0.f;
  __stcs(reinterpret_cast<float4*>(p.out + row * p.ldo + f), make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv));
  if (p.H == 1 ? (gl == 0) : (f % p.C == 0)) p.rowsum[row * p.H + h] = s;
}

template <int G, int U>
__global__ void __launch_bounds__(GAT_THREADS) gat_fwd_long_kernel(const GatParams p) {
  constexpr int Q = GAT_THREADS / G;
  __shared__ float sm_acc[Q][G * 4];
  __shared__ float sm_s[Q][32];
  const int gl = threadIdx.x % G, q = threadIdx.x / G;
  const int64_t item = blockIdx.x;
  const int32_t slot = p.item_long[item];
  const int64_t row = p.long_rows[slot];
  const int64_t rs = p.item_start[item];
  const int64_t rend = __ldg(p.rowptr + row + 1);
  const int64_t re = (rs + p.long_chunk < rend) ? rs + p.long_chunk : rend;
  const int64_t per = (re - rs + Q - 1) / Q;
  int64_t k0 = rs + (int64_t)q * per, k1 = k0 + per;
  if (k0 > re) k0 = re;
  if (k1 > re) k1 = re;
  int f, h;
  bool active;
  lane_slot(p, gl, f, h, active);
  const float ad = __ldg(p.a_dst + row * p.H + h);
  const float M = long_row_max(p, slot, h);
  float s = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  gat_fwd_range<G, U>(p, k0, k1, gl, f, h, ad, M, s, acc);
  sm_acc[q][gl * 4 + 0] = acc.x;
  sm_acc[q][gl * 4 + 1] = acc.y;
  sm_acc[q][gl * 4 + 2] = acc.z;
  sm_acc[q][gl * 4 + 3] = acc.w;
  if (active && (p.H == 1 ? (gl == 0) : (f % p.C == 0))) sm_s[q][h] = s;
  __syncthreads();
  const int HC = p.H * p.C;
  for (int t = threadIdx.x; t < G * 4; t += GAT_THREADS) {
    float v = 0.f;
    for (int qq = 0; qq < Q; ++qq) v += sm_acc[qq][t];
    if (t < HC) p.part_acc[item * p.ldp + t] = v;
  }
  if (threadIdx.x < p.H) {
    float v = 0.f;
    for (int qq = 0; qq < Q; ++qq) v += sm_s[qq][threadIdx.x];
    p.part_s[item * p.H + threadIdx.x] = v;
  }
}

// one thread per (long row, feature): sum the items in order, normalise, record the statistics
__global__ void __launch_bounds__(256) gat_fwd_combine_kernel(const GatParams p) {
  const int HC = p.H * p.C;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t slot = t / HC;
  const int f = (int)(t - slot * HC);
  if (slot >= p.n_long) return;
  const int64_t row = p.long_rows[slot];
  const int h = (p.H == 1) ? 0 : f / p.C;
  float s = 0.f, acc = 0.f, m = -INFINITY;
  for (int32_t it = p.long_item_ptr[slot]; it < p.long_item_ptr[slot + 1]; ++it) {
    s += p.part_s[(int64_t)it * p.H + h];
    acc += p.part_acc[(int64_t)it * p.ldp + f];
    m = fmaxf(m, p.item_max[(int64_t)it * p.H + h]);
  }
  p.out[row * p.ldo + f] = acc * (1.0f / (s + 1e-16f));
  if (p.H == 1 ? (f == 0) : (f % p.C == 0)) {
    p.rowsum[row * p.H + h] = s;
    p.rowmax[row * p.H + h] = m;
  }
}

// ------------------------------------------------------------------------------------------
// backward
//   alpha_ij = exp(leaky(a_s[j]+a_d[i]) - max_i) / (sum_i + 1e-16)
//   dXp[j]   = sum_i (alpha_ij*drop) * dout[i]
//   dalpha   = drop * <dout[i,h,:], Xp[j,h,:]> ;  de = alpha * (dalpha - S[i,h]) ; dlogit = de * leaky'
//   da_src[j,h] = sum_i dlogit  (row-local) ;  da_dst[i,h] += dlogit (atomic)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gat_bwd_prep_kernel(const float* __restrict__ dout, int64_t ldd, const float* __restrict__ out, int64_t ldo,
                    const float* __restrict__ a_dst, const float* __restrict__ rowmax, const float* __restrict__ rowsum,
                    int64_t n, int H, int C, float4* __restrict__ stats) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * H) return;
  const int64_t i = t / H;
  const int h = (int)(t - i * H);
  const float* a = dout + i * ldd + h * C;
  const float* b = out + i * ldo + h * C;
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += a[c] * b[c];
  stats[t] = make_float4(a_dst[t], rowmax[t], 1.0f / (rowsum[t] + 1e-16f), s);
}

template <int G, int U>
__device__ __forceinline__ void gat_bwd_range(const GatParams& p, int64_t k0, int64_t k1, int gl, int f, int h, bool active,
                                              float as, const float4& xj, float4& acc, float& das) {
  const int len = (k1 > k0) ? (int)(k1 - k0) : 0;
  const int maxlen = __reduce_max_sync(FULLMASK, len);
  if (maxlen == 0) return;
  const int32_t* __restrict__ col = p.col + k0;
  const char* dbase = reinterpret_cast<const char*>(p.dout + f);
  const uint32_t row_bytes = (uint32_t)(p.ldd * 4);
  const float4* stb = p.stats + h;
  const bool head_lead = active && (gl % p.LPH) == 0;
  int32_t cl = (gl < len) ? __ldcs(col + gl) : 0;
  for (int off = 0; off < maxlen; off += G) {
    int nb = len - off;
    nb = nb < 0 ? 0 : (nb > G ? G : nb);
    int32_t cn = 0;
    if (off + G + gl < len) cn = __ldcs(col + off + G + gl);
    const int nbmax = __reduce_max_sync(FULLMASK, nb);
#pragma unroll 1
    for (int j = 0; j < nbmax; j += U) {
      float4 d[U], st[U];
      uint32_t ci[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        ci[u] = (uint32_t)__shfl_sync(FULLMASK, cl, j + u, G);
        d[u] = __ldg(reinterpret_cast<const float4*>(dbase + (size_t)ci[u] * row_bytes));
        st[u] = __ldg(stb + (size_t)ci[u] * p.H);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        // the dot product is reduced over the LPH lanes of a head: executed by the whole warp
        float dot = d[u].x * xj.x + d[u].y * xj.y + d[u].z * xj.z + d[u].w * xj.w;
        for (int o = p.LPH >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(FULLMASK, dot, o);
        if (j + u < nb) {
          const float raw = as + st[u].x;
          const float alpha = __expf(leaky_relu(raw, p.slope) - st[u].y) * st[u].z;
          const float dr = p.drop ? __ldg(p.drop + (int64_t)__ldg(p.tpos + k0 + off + j + u) * p.H + h) : 1.0f;
          const float w = alpha * dr;
          acc.x += w * d[u].x;
          acc.y += w * d[u].y;
          acc.z += w * d[u].z;
          acc.w += w * d[u].w;
          const float dl = alpha * (dr * dot - st[u].w) * (raw > 0.f ? 1.0f : p.slope);
          if (head_lead) {
            das += dl;
            atomicAdd(p.da_dst + (size_t)ci[u] * p.H + h, dl);
          }
        }
      }
    }
    cl = cn;
  }
}

template <int G, int U>
__global__ void __launch_bounds__(GAT_THREADS) gat_bwd_rows_kernel(const GatParams p) {
  constexpr int GPB = GAT_THREADS / G;
  const int gl = threadIdx.x % G;
  const int64_t gid = (int64_t)blockIdx.x * GPB + threadIdx.x / G;
  int64_t row = -1, k0 = 0, k1 = 0;
  if (gid < p.n_rows) {
    row = p.row_order ? (int64_t)__ldg(p.row_order + gid) : gid;
    k0 = __ldg(p.rowptr + row);
    k1 = __ldg(p.rowptr + row + 1);
    if (p.n_items > 0 && k1 - k0 > p.chunk) row = -1;
  }
  if (row < 0) k1 = k0;
  int f, h;
  bool active;
  lane_slot(p, gl, f, h, active);
  float as = 0.f;
  float4 xj = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row >= 0) {
    as = __ldg(p.a_src + row * p.H + h);
    if (active) xj = __ldg(reinterpret_cast<const float4*>(p.Xp + row * p.ldx + f));
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float das = 0.f;
  gat_bwd_range<G, U>(p, k0, k1, gl, f, h, active, as, xj, acc, das);
  if (row < 0 || !active) return;
  __stcs(reinterpret_cast<float4*>(p.dXp + row * p.lddx + f), acc);
  if ((gl % p.LPH) == 0) p.da_src[row * p.H + h] = das;
}

template <int G, int U>
__global__ void __launch_bounds__(GAT_THREADS) gat_bwd_long_kernel(const GatParams p) {
  constexpr int Q = GAT_THREADS / G;
  __shared__ float sm_acc[Q][G * 4];
  __shared__ float sm_s[Q][32];
  const int gl = threadIdx.x % G, q = threadIdx.x / G;
  const int64_t item = blockIdx.x;
  const int32_t slot = p.item_long[item];
  const int64_t row = p.long_rows[slot];
  const int64_t rs = p.item_start[item];
  const int64_t rend = __ldg(p.rowptr + row + 1);
  const int64_t re = (rs + p.long_chunk < rend) ? rs + p.long_chunk : rend;
  const int64_t per = (re - rs + Q - 1) / Q;
  int64_t k0 = rs + (int64_t)q * per, k1 = k0 + per;
  if (k0 > re) k0 = re;
  if (k1 > re) k1 = re;
  int f, h;
  bool active;
  lane_slot(p, gl, f, h, active);
  const float as = __ldg(p.a_src + row * p.H + h);
  float4 xj = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) xj = __ldg(reinterpret_cast<const float4*>(p.Xp + row * p.ldx + f));
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float das = 0.f;
  gat_bwd_range<G, U>(p, k0, k1, gl, f, h, active, as, xj, acc, das);
  sm_acc[q][gl * 4 + 0] = acc.x;
  sm_acc[q][gl * 4 + 1] = acc.y;
  sm_acc[q][gl * 4 + 2] = acc.z;
  sm_acc[q][gl * 4 + 3] = acc.w;
  if (active && (gl % p.LPH) == 0) sm_s[q][h] = das;
  __syncthreads();
  const int HC = p.H * p.C;
  for (int t = threadIdx.x; t < G * 4; t += GAT_THREADS) {
    float v = 0.f;
    for (int qq = 0; qq < Q; ++qq) v += sm_acc[qq][t];
    if (t < HC) p.part_acc[item * p.ldp + t] = v;
  }
  if (threadIdx.x < p.H) {
    float v = 0.f;
    for (int qq = 0; qq < Q; ++qq) v += sm_s[qq][threadIdx.x];
    p.part_s[item * p.H + threadIdx.x] = v;
  }
}

__global__ void __launch_bounds__(256) gat_bwd_combine_kernel(const GatParams p) {
  const int HC = p.H * p.C;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t slot = t / HC;
  const int f = (int)(t - slot * HC);
  if (slot >= p.n_long) return;
  const int64_t row = p.long_rows[slot];
  const int h = (p.H == 1) ? 0 : f / p.C;
  float acc = 0.f, s = 0.f;
  for (int32_t it = p.long_item_ptr[slot]; it < p.long_item_ptr[slot + 1]; ++it) {
    acc += p.part_acc[(int64_t)it * p.ldp + f];
    s += p.part_s[(int64_t)it * p.H + h];
  }
  p.dXp[row * p.lddx + f] = acc;
  if (p.H == 1 ? (f == 0) : (f % p.C == 0)) p.da_src[row * p.H + h] = s;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int gat_group(int HC) {
  const int nvec = (HC + 3) / 4;
  int G = 1;
  while (G < nvec) G <<= 1;
  return G;
}

static inline bool al16(const void* p, int64_t ld) { return (((uintptr_t)p) & 15) == 0 && (ld % 4) == 0; }

static void fill_graph(GatParams& p, const rgbmp_graph_t* g) {
  p.rowptr = g->rowptr;
  p.col = g->col;
  p.row_order = g->row_order;
  p.n_rows = g->n_rows;
  const bool split = g->n_items > 0 && g->long_rows && g->long_item_ptr && g->item_long && g->item_start &&
                     g->chunk > 0 && g->long_chunk > 0;
  p.chunk = split ? g->chunk : 0;
  p.long_chunk = g->long_chunk;
  p.long_rows = g->long_rows;
  p.long_item_ptr = g->long_item_ptr;
  p.item_long = g->item_long;
  p.item_start = g->item_start;
  p.n_long = split ? g->n_long : 0;
  p.n_items = split ? g->n_items : 0;
}

// scratch of the long-row split: item_max [n_items,H] | part_s [n_items,H] | part_acc [n_items, ldp]
static size_t split_bytes(int64_t n_items, int H, int HC) {
  if (n_items <= 0) return 0;
  return 3 * 256 + align_up((size_t)n_items * H * 4, 256) * 2 + (size_t)n_items * align_up((size_t)HC, 4) * 4;
}

static bool carve_split(GatParams& p, void* ws, size_t ws_bytes, int HC) {
  p.item_max = p.part_s = p.part_acc = nullptr;
  p.ldp = (int64_t)align_up((size_t)HC, 4);
  if (p.n_items <= 0) return true;
  if (!ws) return false;
  Carver cv(ws, ws_bytes);
  p.item_max = cv.take<float>((size_t)p.n_items * p.H);
  p.part_s = cv.take<float>((size_t)p.n_items * p.H);
  p.part_acc = cv.take<float>((size_t)p.n_items * p.ldp);
  return cv.ok();
}

#define GAT_LAUNCH_G(KERNEL, GRID, ...)                                                         \
  switch (G) {                                                                                  \
    case 1: KERNEL<1, 4><<<(unsigned)(GRID(1)), GAT_THREADS, 0, st>>>(__VA_ARGS__); break;      \
    case 2: KERNEL<2, 4><<<(unsigned)(GRID(2)), GAT_THREADS, 0, st>>>(__VA_ARGS__); break;      \
    case 4: KERNEL<4, 4><<<(unsigned)(GRID(4)), GAT_THREADS, 0, st>>>(__VA_ARGS__); break;      \
    case 8: KERNEL<8, 4><<<(unsigned)(GRID(8)), GAT_THREADS, 0, st>>>(__VA_ARGS__); break;      \
    case 16: KERNEL<16, 4><<<(unsigned)(GRID(16)), GAT_THREADS, 0, st>>>(__VA_ARGS__); break;   \
    default: KERNEL<32, 4><<<(unsigned)(GRID(32)), GAT_THREADS, 0, st>>>(__VA_ARGS__); break;   \
  }

}  // namespace rgbmp

using namespace rgbmp;

extern "C" {

size_t rgbmp_gat_workspace_bytes(const rgbmp_graph_t* g, int H, int C) {
  if (!g || H <= 0 || C <= 0) return 256;
  return 256 + split_bytes(g->n_items, H, H * C);
}

size_t rgbmp_gat_backward_workspace_bytes(const rgbmp_graph_t* gT, int64_t n_dst, int H, int C) {
  if (!gT || H <= 0 || C <= 0 || n_dst < 0) return 256;
  return 512 + align_up((size_t)n_dst * H * sizeof(float4), 256) + split_bytes(gT->n_items, H, H * C);
}

int rgbmp_gat_forward(const rgbmp_graph_t* g, const float* Xp, int64_t ldx, const float* a_src, const float* a_dst,
                      int H, int C, float slope, const float* drop, float* out, int64_t ldo, float* rowmax,
                      float* rowsum, void* ws, size_t ws_bytes, int device, void* stream) {
  if (!g || !g->rowptr || g->n_rows < 0 || g->nnz < 0 || (g->nnz > 0 && !g->col))
    return fail(RGBMP_EINVAL, "rgbmp_gat_forward: bad graph descriptor");
  if (!Xp || !a_src || !a_dst || !out || !rowmax || !rowsum || H <= 0 || C <= 0)
    return fail(RGBMP_EINVAL, "rgbmp_gat_forward: null pointer / bad H,C");
  const int HC = H * C;
  if (HC > 128 || H > 32) return fail(RGBMP_ERANGE, "rgbmp_gat_forward: H*C = %d > 128 (use the unfused kernels)", HC);
  if (H != 1 && (C % 4) != 0) return fail(RGBMP_ERANGE, "rgbmp_gat_forward: needs H == 1 or C %% 4 == 0");
  if (!al16(Xp, ldx) || !al16(out, ldo) || ldx < (int64_t)align_up(HC, 4) || ldo < (int64_t)align_up(HC, 4))
    return fail(RGBMP_EALIGN, "rgbmp_gat_forward: Xp/out need 16-byte aligned rows with ld >= roundup(H*C,4)");
  if (g->n_cols >= (1ll << 31) / 4) {
    if ((double)g->n_cols * (double)ldx * 4.0 >= 1.8e19) return fail(RGBMP_ERANGE, "rgbmp_gat_forward: too large");
  }
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RGBMP_EINVAL, "rgbmp_gat_forward: bad device");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = g->n_rows;
  if (n == 0) return 0;
  GatParams p = {};
  fill_graph(p, g);
  p.Xp = Xp; p.ldx = ldx; p.a_src = a_src; p.a_dst = a_dst; p.H = H; p.C = C; p.slope = slope; p.drop = drop;
  p.out = out; p.ldo = ldo; p.rowmax = rowmax; p.rowsum = rowsum;
  if (!carve_split(p, ws, ws_bytes, HC))
    return fail(RGBMP_EWORKSPACE, "rgbmp_gat_forward: workspace %zu < %zu", ws_bytes, rgbmp_gat_workspace_bytes(g, H, C));
  int HP = 1;
  while (HP < H) HP <<= 1;
  gat_max_rows_kernel<<<(unsigned)ceil_div(n * 32, GAT_THREADS), GAT_THREADS, 0, st>>>(p, HP);
  RGBMP_LAUNCH_CHECK("gat_max_rows_kernel");
  if (p.n_items > 0) {
    gat_max_long_kernel<<<(unsigned)p.n_items, GAT_THREADS, 0, st>>>(p, HP);
    RGBMP_LAUNCH_CHECK("gat_max_long_kernel");
  }
  const int G = gat_group(HC);
#define ROWS_GRID(g_) ceil_div(n, GAT_THREADS / (g_))
#define ITEM_GRID(g_) p.n_items
  GAT_LAUNCH_G(gat_fwd_rows_kernel, ROWS_GRID, p)
  RGBMP_LAUNCH_CHECK("gat_fwd_rows_kernel");
  if (p.n_items > 0) {
    GAT_LAUNCH_G(gat_fwd_long_kernel, ITEM_GRID, p)
    RGBMP_LAUNCH_CHECK("gat_fwd_long_kernel");
    gat_fwd_combine_kernel<<<(unsigned)ceil_div(p.n_long * HC, 256), 256, 0, st>>>(p);
    RGBMP_LAUNCH_CHECK("gat_fwd_combine_kernel");
  }
  return 0;
}

int rgbmp_gat_backward(const rgbmp_graph_t* gT, const float* Xp, int64_t ldx, const float* a_src, const float* a_dst,
                       int H, int C, float slope, const float* drop, const int32_t* tpos, const float* rowmax,
                       const float* rowsum, const float* out, int64_t ldo, const float* dout, int64_t ldd, float* dXp,
                       int64_t lddx, float* da_src, float* da_dst, int64_t n_dst, void* ws, size_t ws_bytes, int device,
                       void* stream) {
  if (!gT || !gT->rowptr || gT->n_rows < 0 || gT->nnz < 0 || (gT->nnz > 0 && !gT->col))
    return fail(RGBMP_EINVAL, "rgbmp_gat_backward: bad graph descriptor");
  if (!Xp || !a_src || !a_dst || !rowmax || !rowsum || !out || !dout || !dXp || !da_src || !da_dst || H <= 0 || C <= 0 ||
      (drop && !tpos) || n_dst < 0 || !ws)
    return fail(RGBMP_EINVAL, "rgbmp_gat_backward: null pointer / bad H,C");
  const int HC = H * C;
  if (HC > 128 || H > 32) return fail(RGBMP_ERANGE, "rgbmp_gat_backward: H*C = %d > 128", HC);
  const int G = gat_group(HC);
  int LPH = G;
  if (H != 1) {
    if ((C % 4) != 0 || ((C / 4) & (C / 4 - 1)) != 0)
      return fail(RGBMP_ERANGE, "rgbmp_gat_backward: needs H == 1 or C in {4,8,16,32,64,128}");
    LPH = C / 4;
  }
  if (!al16(Xp, ldx) || !al16(dout, ldd) || !al16(dXp, lddx))
    return fail(RGBMP_EALIGN, "rgbmp_gat_backward: Xp/dout/dXp need 16-byte aligned rows");
  if (ws_bytes < rgbmp_gat_backward_workspace_bytes(gT, n_dst, H, C))
    return fail(RGBMP_EWORKSPACE, "rgbmp_gat_backward: workspace %zu < %zu", ws_bytes,
                rgbmp_gat_backward_workspace_bytes(gT, n_dst, H, C));
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RGBMP_EINVAL, "rgbmp_gat_backward: bad device");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = gT->n_rows;
  if (n == 0) return 0;
  GatParams p = {};
  fill_graph(p, gT);
  p.Xp = Xp; p.ldx = ldx; p.a_src = a_src; p.a_dst = a_dst; p.H = H; p.C = C; p.LPH = LPH; p.slope = slope;
  p.drop = drop; p.tpos = tpos; p.dout = dout; p.ldd = ldd; p.dXp = dXp; p.lddx = lddx; p.da_src = da_src; p.da_dst = da_dst;
  float4* stats = reinterpret_cast<float4*>(ws);
  const size_t stats_bytes = align_up((size_t)n_dst * H * sizeof(float4), 256);
  if (!carve_split(p, (char*)ws + stats_bytes, ws_bytes - stats_bytes, HC))
    return fail(RGBMP_EWORKSPACE, "rgbmp_gat_backward: workspace carve");
  p.stats = stats;
  if (n_dst > 0) {
    gat_bwd_prep_kernel<<<(unsigned)ceil_div(n_dst * H, 256), 256, 0, st>>>(dout, ldd, out, ldo, a_dst, rowmax, rowsum, n_dst,
                                                                           H, C, stats);
    RGBMP_LAUNCH_CHECK("gat_bwd_prep_kernel");
  }
  GAT_LAUNCH_G(gat_bwd_rows_kernel, ROWS_GRID, p)
  RGBMP_LAUNCH_CHECK("gat_bwd_rows_kernel");
  if (p.n_items > 0) {
    GAT_LAUNCH_G(gat_bwd_long_kernel, ITEM_GRID, p)
    RGBMP_LAUNCH_CHECK("gat_bwd_long_kernel");
    gat_bwd_combine_kernel<<<(unsigned)ceil_div(p.n_long * HC, 256), 256, 0, st>>>(p);
    RGBMP_LAUNCH_CHECK("gat_bwd_combine_kernel");
  }
  return 0;
}

}  // extern "C"
