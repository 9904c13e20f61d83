// Fused GATConv edge-softmax + aggregate, forward and backward (sm_100a), v3.
//
// Gather / segment-reduce over CSR: L2/HBM-bound by bytes, but the ncu profile of v2 showed the
// Reddit-shaped layer ISSUE-bound (86 % issue slots busy, 1.3 ms of 4.7 ms in a separate row-max
// pass).  v3 therefore
//   * runs ONE pass with an online softmax: a running maximum per (row, head) rescales the running
//     sum and accumulator when it grows (once per batch of U edges, a rarely taken branch), so the
//     lane-dense max pre-pass and its second read of the column ids are gone; the result equals
//     torch_geometric.utils.softmax (max, exp(e - max), /(sum + 1e-16)) up to fp32
//     rounding of the rescales;
//   * gives a lane VPL = 2 consecutive float4 (a whole 8-channel head) whenever the shape allows:
//     the logit -> exp work is shared by 8 channels instead of 4 and, for C = 8, the per-head dot
//     product of the backward needs no shuffle;
//   * works in the log2 domain: a_src, a_dst are pre-multiplied by log2(e) (leaky_relu commutes with
//     a positive scale), so exp(x) is a bare ex2.approx;
//   * keeps the SpMM machinery: degree-sorted row schedule, one coalesced load of G column ids per
//     batch + shuffle broadcast, software-pipelined full batches, rows longer than `chunk` split
//     into CTA work items whose (max, sum, acc) partials are merged in item order -- deterministic,
//     no atomics;
//   * backward on the transpose CSR with the per-target statistics (a_dst*log2e, max*log2e,
//     1/(sum+1e-16), S = <dout_i, out_i>) packed into one float4 per (node, head).
// The only atomics are the da_dst accumulation of the backward (one float per edge and head).
#include "common.cuh"

namespace rgbmp {

constexpr unsigned FULLMASK = 0xffffffffu;
constexpr int GAT_THREADS = 256;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float leaky_relu(float x, float slope) { return x > 0.f ? x : x * slope; }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void fma4(float4& a, float w, const float4& x) {
  a.x = fmaf(w, x.x, a.x);
  a.y = fmaf(w, x.y, a.y);
  a.z = fmaf(w, x.z, a.z);
  a.w = fmaf(w, x.w, a.w);
}
__device__ __forceinline__ void scale4(float4& a, float w) {
  a.x *= w; a.y *= w; a.z *= w; a.w *= w;
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

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
  const float* a_src;   // forward: a_src * log2(e) (scaled copy); backward: the caller's a_src
  const float* a_dst;   // forward: a_dst * log2(e)
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
  const float4* stats;   // [n_dst, H]: (a_dst*log2e, rowmax*log2e, 1/(rowsum+1e-16), S)
  const float* dout;
  int64_t ldd;
  float* dXp;
  int64_t lddx;
  float* da_src;
  float* da_dst;
  // long-row scratch
  float* part_m;     // [n_items, H]  (forward: running max of the item, log2 domain)
  float* part_acc;   // [n_items, ldp]
  float* part_s;     // [n_items, H]
  int64_t ldp;
};

// lane -> feature offsets of its VPL float4 vectors, head, number of vectors that exist.  Vectors
// beyond H*C shadow the lane's first vector (valid memory, never stored); a lane without any
// vector shadows vector 0 of the row.
template <int VPL>
__device__ __forceinline__ void lane_slot(const GatParams& p, int gl, int (&fo)[VPL], int& h, int& nact) {
  const int HC = p.H * p.C;
  int f = gl * 4 * VPL;
  nact = 0;
#pragma unroll
  for (int v = 0; v < VPL; ++v) nact += (f + 4 * v < HC) ? 1 : 0;
  if (nact == 0) f = 0;
#pragma unroll
  for (int v = 0; v < VPL; ++v) fo[v] = (v < nact) ? f + 4 * v : f;
  h = (p.H == 1) ? 0 : f / p.C;
}
__device__ __forceinline__ bool head_lead(const GatParams& p, int gl, int f0, int nact) {
  return nact > 0 && (p.H == 1 ? gl == 0 : (f0 % p.C) == 0);
}

// scaled copies of the attention terms: out[i] = in[i] * log2(e)
__global__ void __launch_bounds__(256) gat_scale_kernel(const float* __restrict__ in, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] * LOG2E;
}

// ------------------------------------------------------------------------------------------
// forward: online softmax over edges [k0,k1) of one target row.  State per lane (= per head):
// running max m (log2 domain), running sum s, accumulator acc.  Warp-uniform trip counts.
// ------------------------------------------------------------------------------------------
template <int VPL, int U>
__device__ __forceinline__ void gat_consume(const GatParams& p, const float4 (&x)[U][VPL], const float (&as)[U], int nvalid,
                                            int64_t kbase, int h, float ad, float& m, float& s, float4 (&acc)[VPL]) {
  float e[U];
  float mb = -INFINITY;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    e[u] = leaky_relu(as[u] + ad, p.slope);
    if (u < nvalid) mb = fmaxf(mb, e[u]);
  }
  if (mb > m) {                      // rare after the first batches of a row
    const float sc = ex2(m - mb);    // m = -inf on the first batch: ex2(-inf) = 0
    s *= sc;
#pragma unroll
    for (int v = 0; v < VPL; ++v) scale4(acc[v], sc);
    m = mb;
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (u < nvalid) {
      const float pe = ex2(e[u] - m);
      s += pe;
      const float w = p.drop ? pe * __ldg(p.drop + (kbase + u) * p.H + h) : pe;
#pragma unroll
      for (int v = 0; v < VPL; ++v) fma4(acc[v], w, x[u][v]);
    }
  }
}

template <int G, int VPL, int U>
__device__ __forceinline__ void gat_fwd_range(const GatParams& p, int64_t k0, int64_t k1, int gl, const int (&fo)[VPL], int h,
                                              float ad, float& m, float& s, float4 (&acc)[VPL]) {
  const int len = (k1 > k0) ? (int)(k1 - k0) : 0;
  const int maxlen = __reduce_max_sync(FULLMASK, len);
  if (maxlen == 0) return;
  const int32_t* __restrict__ col = p.col + k0;
  const char* xb[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) xb[v] = reinterpret_cast<const char*>(p.Xp + fo[v]);
  const uint32_t row_bytes = (uint32_t)(p.ldx * 4);
  const char* asb = reinterpret_cast<const char*>(p.a_src + h);
  const uint32_t as_bytes = (uint32_t)(p.H * 4);
  auto gather = [&](uint32_t c, float4 (&xx)[VPL], float& aa) {
#pragma unroll
    for (int v = 0; v < VPL; ++v) xx[v] = __ldg(reinterpret_cast<const float4*>(xb[v] + (size_t)c * row_bytes));
    aa = __ldg(reinterpret_cast<const float*>(asb + (size_t)c * as_bytes));
  };
  int32_t cl = (gl < len) ? __ldcs(col + gl) : 0;
  for (int off = 0; off < maxlen; off += G) {
    int nb = len - off;
    nb = nb < 0 ? 0 : (nb > G ? G : nb);
    int32_t cn = 0;
    if (off + G + gl < len) cn = __ldcs(col + off + G + gl);
    if (G >= 2 * U && __all_sync(FULLMASK, nb == G)) {
      // full batch in every group: software-pipelined -- the gathers of step j+U are in flight
      // while step j is consumed
      float4 x[U][VPL];
      float as[U];
#pragma unroll
      for (int u = 0; u < U; ++u) gather((uint32_t)__shfl_sync(FULLMASK, cl, u, G), x[u], as[u]);
#pragma unroll 1
      for (int j = 0; j < G; j += U) {
        float4 xn[U][VPL];
        float asn[U];
        const int jn = (j + U < G) ? j + U : j;      // last step re-requests itself (L1 hit, unused)
#pragma unroll
        for (int u = 0; u < U; ++u) gather((uint32_t)__shfl_sync(FULLMASK, cl, jn + u, G), xn[u], asn[u]);
        gat_consume<VPL, U>(p, x, as, U, k0 + off + j, h, ad, m, s, acc);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          as[u] = asn[u];
#pragma unroll
          for (int v = 0; v < VPL; ++v) x[u][v] = xn[u][v];
        }
      }
    } else {
      const int nbmax = __reduce_max_sync(FULLMASK, nb);
#pragma unroll 1
      for (int j = 0; j < nbmax; j += U) {
        float4 x[U][VPL];
        float as[U];
#pragma unroll
        for (int u = 0; u < U; ++u)   // lanes past nb hold id 0: a valid row
          gather((uint32_t)__shfl_sync(FULLMASK, cl, (j + u) & (G - 1), G), x[u], as[u]);
        int nv = nb - j;
        nv = nv < 0 ? 0 : (nv > U ? U : nv);
        gat_consume<VPL, U>(p, x, as, nv, k0 + off + j, h, ad, m, s, acc);
      }
    }
    cl = cn;
  }
}

template <int G, int VPL, int U>
__global__ void __launch_bounds__(GAT_THREADS, 4) gat_fwd_rows_kernel(const GatParams p) {
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
  int fo[VPL], h, nact;
  lane_slot<VPL>(p, gl, fo, h, nact);
  const float ad = (row >= 0) ? __ldg(p.a_dst + row * p.H + h) : 0.f;
  float m = -INFINITY, s = 0.f;
  float4 acc[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  gat_fwd_range<G, VPL, U>(p, k0, k1, gl, fo, h, ad, m, s, acc);
  if (row < 0 || nact == 0) return;
  const float inv = (k1 > k0) ? 1.0f / (s + 1e-16f) : 0.f;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    if (v < nact) {
      scale4(acc[v], inv);
      __stcs(reinterpret_cast<float4*>(p.out + row * p.ldo + fo[v]), acc[v]);
    }
  }
  if (head_lead(p, gl, fo[0], nact)) {
    p.rowsum[row * p.H + h] = s;
    p.rowmax[row * p.H + h] = (k1 > k0) ? m * (1.0f / LOG2E) : 0.f;
  }
}

// long rows: one CTA per work item; the CTA's Q groups take contiguous sub-ranges, their
// (max, sum, acc) are merged through shared memory in group order
template <int G, int VPL, int U>
__global__ void __launch_bounds__(GAT_THREADS, 4) gat_fwd_long_kernel(const GatParams p) {
  constexpr int Q = GAT_THREADS / G;
  constexpr int W = G * 4 * VPL;
  __shared__ float sm_acc[Q * W];
  __shared__ float sm_m[Q][G];      // heads per row <= lanes per row
  __shared__ float sm_s[Q][G];
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
  int fo[VPL], h, nact;
  lane_slot<VPL>(p, gl, fo, h, nact);
  const float ad = __ldg(p.a_dst + row * p.H + h);
  float m = -INFINITY, s = 0.f;
  float4 acc[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  gat_fwd_range<G, VPL, U>(p, k0, k1, gl, fo, h, ad, m, s, acc);
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const bool act = v < nact;
    float* d = sm_acc + q * W + (gl * VPL + v) * 4;
    d[0] = act ? acc[v].x : 0.f;
    d[1] = act ? acc[v].y : 0.f;
    d[2] = act ? acc[v].z : 0.f;
    d[3] = act ? acc[v].w : 0.f;
  }
  if (head_lead(p, gl, fo[0], nact)) {
    sm_m[q][h] = m;
    sm_s[q][h] = s;
  }
  __syncthreads();
  const int HC = p.H * p.C;
  for (int t = threadIdx.x; t < W; t += GAT_THREADS) {
    if (t >= HC) continue;
    const int ht = (p.H == 1) ? 0 : t / p.C;
    float M = -INFINITY;
    for (int qq = 0; qq < Q; ++qq) M = fmaxf(M, sm_m[qq][ht]);
    float v = 0.f;
    for (int qq = 0; qq < Q; ++qq) {
      const float mq = sm_m[qq][ht];
      if (mq > -INFINITY) v = fmaf(sm_acc[qq * W + t], ex2(mq - M), v);
    }
    p.part_acc[item * p.ldp + t] = v;
  }
  if (threadIdx.x < p.H) {
    const int ht = threadIdx.x;
    float M = -INFINITY;
    for (int qq = 0; qq < Q; ++qq) M = fmaxf(M, sm_m[qq][ht]);
    float v = 0.f;
    for (int qq = 0; qq < Q; ++qq) {
      const float mq = sm_m[qq][ht];
      if (mq > -INFINITY) v = fmaf(sm_s[qq][ht], ex2(mq - M), v);
    }
    p.part_m[item * p.H + ht] = M;
    p.part_s[item * p.H + ht] = v;
  }
}

// one thread per (long row, feature): merge the items in order, normalise, record the statistics
__global__ void __launch_bounds__(256) gat_fwd_combine_kernel(const GatParams p) {
  const int HC = p.H * p.C;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t slot = t / HC;
  const int f = (int)(t - slot * HC);
  if (slot >= p.n_long) return;
  const int64_t row = p.long_rows[slot];
  const int h = (p.H == 1) ? 0 : f / p.C;
  const int32_t i0 = p.long_item_ptr[slot], i1 = p.long_item_ptr[slot + 1];
  float M = -INFINITY;
  for (int32_t it = i0; it < i1; ++it) M = fmaxf(M, p.part_m[(int64_t)it * p.H + h]);
  float s = 0.f, acc = 0.f;
  for (int32_t it = i0; it < i1; ++it) {
    const float mq = p.part_m[(int64_t)it * p.H + h];
    if (mq > -INFINITY) {
      const float sc = ex2(mq - M);
      s = fmaf(p.part_s[(int64_t)it * p.H + h], sc, s);
      acc = fmaf(p.part_acc[(int64_t)it * p.ldp + f], sc, acc);
    }
  }
  p.out[row * p.ldo + f] = acc * (1.0f / (s + 1e-16f));
  if (p.H == 1 ? (f == 0) : (f % p.C == 0)) {
    p.rowsum[row * p.H + h] = s;
    p.rowmax[row * p.H + h] = M * (1.0f / LOG2E);
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
  stats[t] = make_float4(a_dst[t] * LOG2E, rowmax[t] * LOG2E, 1.0f / (rowsum[t] + 1e-16f), s);
}

template <int G, int VPL, int U>
__device__ __forceinline__ void gat_bwd_range(const GatParams& p, int64_t k0, int64_t k1, int gl, const int (&fo)[VPL], int h,
                                              bool lead, float as2, const float4 (&xj)[VPL], float4 (&acc)[VPL], float& das) {
  const int len = (k1 > k0) ? (int)(k1 - k0) : 0;
  const int maxlen = __reduce_max_sync(FULLMASK, len);
  if (maxlen == 0) return;
  const int32_t* __restrict__ col = p.col + k0;
  const char* db[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) db[v] = reinterpret_cast<const char*>(p.dout + fo[v]);
  const uint32_t row_bytes = (uint32_t)(p.ldd * 4);
  const char* stb = reinterpret_cast<const char*>(p.stats + h);
  const uint32_t st_bytes = (uint32_t)(p.H * 16);
  int32_t cl = (gl < len) ? __ldcs(col + gl) : 0;
  for (int off = 0; off < maxlen; off += G) {
    int nb = len - off;
    nb = nb < 0 ? 0 : (nb > G ? G : nb);
    int32_t cn = 0;
    if (off + G + gl < len) cn = __ldcs(col + off + G + gl);
    const int nbmax = __reduce_max_sync(FULLMASK, nb);
#pragma unroll 1
    for (int j = 0; j < nbmax; j += U) {
      float4 d[U][VPL], st[U];
      uint32_t ci[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        ci[u] = (uint32_t)__shfl_sync(FULLMASK, cl, (j + u) & (G - 1), G);
#pragma unroll
        for (int v = 0; v < VPL; ++v) d[u][v] = __ldg(reinterpret_cast<const float4*>(db[v] + (size_t)ci[u] * row_bytes));
        st[u] = __ldg(reinterpret_cast<const float4*>(stb + (size_t)ci[u] * st_bytes));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        // per-head dot product: lane-local, then reduced over the LPH lanes of the head (whole warp executes)
        float dot = 0.f;
#pragma unroll
        for (int v = 0; v < VPL; ++v) dot += dot4(d[u][v], xj[v]);
        for (int o = p.LPH >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(FULLMASK, dot, o);
        if (j + u < nb) {
          const float raw = as2 + st[u].x;
          const float alpha = ex2(leaky_relu(raw, p.slope) - st[u].y) * st[u].z;
          const float dr = p.drop ? __ldg(p.drop + (int64_t)__ldg(p.tpos + k0 + off + j + u) * p.H + h) : 1.0f;
          const float w = alpha * dr;
#pragma unroll
          for (int v = 0; v < VPL; ++v) fma4(acc[v], w, d[u][v]);
          const float dl = alpha * (dr * dot - st[u].w) * (raw > 0.f ? 1.0f : p.slope);
          if (lead) {
            das += dl;
            atomicAdd(p.da_dst + (size_t)ci[u] * p.H + h, dl);
          }
        }
      }
    }
    cl = cn;
  }
}

template <int G, int VPL, int U>
__global__ void __launch_bounds__(GAT_THREADS, 4) gat_bwd_rows_kernel(const GatParams p) {
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
  int fo[VPL], h, nact;
  lane_slot<VPL>(p, gl, fo, h, nact);
  const bool lead = head_lead(p, gl, fo[0], nact);
  float as2 = 0.f;
  float4 xj[VPL], acc[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) xj[v] = acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row >= 0) {
    as2 = __ldg(p.a_src + row * p.H + h) * LOG2E;
#pragma unroll
    for (int v = 0; v < VPL; ++v)
      if (v < nact) xj[v] = __ldg(reinterpret_cast<const float4*>(p.Xp + row * p.ldx + fo[v]));
  }
  float das = 0.f;
  gat_bwd_range<G, VPL, U>(p, k0, k1, gl, fo, h, lead, as2, xj, acc, das);
  if (row < 0 || nact == 0) return;
#pragma unroll
  for (int v = 0; v < VPL; ++v)
    if (v < nact) __stcs(reinterpret_cast<float4*>(p.dXp + row * p.lddx + fo[v]), acc[v]);
  if (lead) p.da_src[row * p.H + h] = das;
}

template <int G, int VPL, int U>
__global__ void __launch_bounds__(GAT_THREADS, 4) gat_bwd_long_kernel(const GatParams p) {
  constexpr int Q = GAT_THREADS / G;
  constexpr int W = G * 4 * VPL;
  __shared__ float sm_acc[Q * W];
  __shared__ float sm_s[Q][G];
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
  int fo[VPL], h, nact;
  lane_slot<VPL>(p, gl, fo, h, nact);
  const bool lead = head_lead(p, gl, fo[0], nact);
  const float as2 = __ldg(p.a_src + row * p.H + h) * LOG2E;
  float4 xj[VPL], acc[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    xj[v] = (v < nact) ? __ldg(reinterpret_cast<const float4*>(p.Xp + row * p.ldx + fo[v])) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float das = 0.f;
  gat_bwd_range<G, VPL, U>(p, k0, k1, gl, fo, h, lead, as2, xj, acc, das);
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const bool act = v < nact;
    float* d = sm_acc + q * W + (gl * VPL + v) * 4;
    d[0] = act ? acc[v].x : 0.f;
    d[1] = act ? acc[v].y : 0.f;
    d[2] = act ? acc[v].z : 0.f;
    d[3] = act ? acc[v].w : 0.f;
  }
  if (lead) sm_s[q][h] = das;
  __syncthreads();
  const int HC = p.H * p.C;
  for (int t = threadIdx.x; t < W; t += GAT_THREADS) {
    float v = 0.f;
    for (int qq = 0; qq < Q; ++qq) v += sm_acc[qq * W + t];
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
// launch shape: VPL = 2 float4 per lane when a lane's 8 floats stay inside one head, else 1;
// G = pow2 >= lanes needed per row
static void gat_shape(int H, int C, int* G, int* VPL) {
  const int HC = H * C;
  const int vpl = (H == 1 || (C % 8) == 0) ? 2 : 1;
  const int lanes = (int)ceil_div(ceil_div(HC, 4), vpl);
  int g = 1;
  while (g < lanes) g <<= 1;
  *G = g;
  *VPL = vpl;
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

// scratch of the long-row split: part_m [n_items,H] | part_s [n_items,H] | part_acc [n_items, ldp]
static size_t split_bytes(int64_t n_items, int H, int HC) {
  if (n_items <= 0) return 0;
  return 3 * 256 + align_up((size_t)n_items * H * 4, 256) * 2 + (size_t)n_items * align_up((size_t)HC, 4) * 4;
}

static bool carve_split(GatParams& p, Carver& cv, int HC) {
  p.part_m = p.part_s = p.part_acc = nullptr;
  p.ldp = (int64_t)align_up((size_t)HC, 4);
  if (p.n_items <= 0) return true;
  p.part_m = cv.take<float>((size_t)p.n_items * p.H);
  p.part_s = cv.take<float>((size_t)p.n_items * p.H);
  p.part_acc = cv.take<float>((size_t)p.n_items * p.ldp);
  return cv.ok();
}

#define GAT_CASE(KERNEL, G_, V_, U_, GRID, ...) \
  KERNEL<G_, V_, U_><<<(unsigned)(GRID(G_)), GAT_THREADS, 0, st>>>(__VA_ARGS__); break;
// VPL = 2 runs U = 2 edges per step (32 gather registers with the pipeline), VPL = 1 runs U = 4
#define GAT_LAUNCH(KERNEL, GRID, ...)                                   \
  if (VPL == 2) {                                                       \
    switch (G) {                                                        \
      case 1: GAT_CASE(KERNEL, 1, 2, 2, GRID, __VA_ARGS__)              \
      case 2: GAT_CASE(KERNEL, 2, 2, 2, GRID, __VA_ARGS__)              \
      case 4: GAT_CASE(KERNEL, 4, 2, 2, GRID, __VA_ARGS__)              \
      case 8: GAT_CASE(KERNEL, 8, 2, 2, GRID, __VA_ARGS__)              \
      default: GAT_CASE(KERNEL, 16, 2, 2, GRID, __VA_ARGS__)            \
    }                                                                   \
  } else {                                                              \
    switch (G) {                                                        \
      case 1: GAT_CASE(KERNEL, 1, 1, 4, GRID, __VA_ARGS__)              \
      case 2: GAT_CASE(KERNEL, 2, 1, 4, GRID, __VA_ARGS__)              \
      case 4: GAT_CASE(KERNEL, 4, 1, 4, GRID, __VA_ARGS__)              \
      case 8: GAT_CASE(KERNEL, 8, 1, 4, GRID, __VA_ARGS__)              \
      case 16: GAT_CASE(KERNEL, 16, 1, 4, GRID, __VA_ARGS__)            \
      default: GAT_CASE(KERNEL, 32, 1, 4, GRID, __VA_ARGS__)            \
    }                                                                   \
  }

}  // namespace rgbmp

using namespace rgbmp;

extern "C" {

size_t rgbmp_gat_workspace_bytes(const rgbmp_graph_t* g, int H, int C) {
  if (!g || H <= 0 || C <= 0) return 256;
  return 1024 + align_up((size_t)g->n_cols * H * 4, 256) + align_up((size_t)g->n_rows * H * 4, 256) +
         split_bytes(g->n_items, H, H * C);
}

size_t rgbmp_gat_backward_workspace_bytes(const rgbmp_graph_t* gT, int64_t n_dst, int H, int C) {
  if (!gT || H <= 0 || C <= 0 || n_dst < 0) return 256;
  return 1024 + align_up((size_t)n_dst * H * sizeof(float4), 256) + split_bytes(gT->n_items, H, H * C);
}

int rgbmp_gat_forward(const rgbmp_graph_t* g, const float* Xp, int64_t ldx, const float* a_src, const float* a_dst,
                      int H, int C, float slope, const float* drop, float* out, int64_t ldo, float* rowmax,
                      float* rowsum, void* ws, size_t ws_bytes, int device, void* stream) {
  if (!g || !g->rowptr || g->n_rows < 0 || g->nnz < 0 || (g->nnz > 0 && !g->col))
    return fail(RGBMP_EINVAL, "rgbmp_gat_forward: bad graph descriptor");
  if (!Xp || !a_src || !a_dst || !out || !rowmax || !rowsum || H <= 0 || C <= 0 || !ws)
    return fail(RGBMP_EINVAL, "rgbmp_gat_forward: null pointer / bad H,C");
  const int HC = H * C;
  if (HC > 128 || H > 32) return fail(RGBMP_ERANGE, "rgbmp_gat_forward: H*C = %d > 128 (use the unfused kernels)", HC);
  if (H != 1 && (C % 4) != 0) return fail(RGBMP_ERANGE, "rgbmp_gat_forward: needs H == 1 or C %% 4 == 0");
  if (!al16(Xp, ldx) || !al16(out, ldo) || ldx < (int64_t)align_up(HC, 4) || ldo < (int64_t)align_up(HC, 4))
    return fail(RGBMP_EALIGN, "rgbmp_gat_forward: Xp/out need 16-byte aligned rows with ld >= roundup(H*C,4)");
  if (ws_bytes < rgbmp_gat_workspace_bytes(g, H, C))
    return fail(RGBMP_EWORKSPACE, "rgbmp_gat_forward: workspace %zu < %zu", ws_bytes, rgbmp_gat_workspace_bytes(g, H, C));
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RGBMP_EINVAL, "rgbmp_gat_forward: bad device");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = g->n_rows;
  if (n == 0) return 0;
  GatParams p = {};
  fill_graph(p, g);
  p.H = H;                                   // carve_split sizes the partials with it
  p.C = C;
  Carver cv(ws, ws_bytes);
  float* as2 = cv.take<float>((size_t)g->n_cols * H);
  float* ad2 = cv.take<float>((size_t)n * H);
  if (!carve_split(p, cv, HC) || !cv.ok()) return fail(RGBMP_EWORKSPACE, "rgbmp_gat_forward: workspace carve");
  p.Xp = Xp; p.ldx = ldx; p.a_src = as2; p.a_dst = ad2; p.H = H; p.C = C; p.slope = slope; p.drop = drop;
  p.out = out; p.ldo = ldo; p.rowmax = rowmax; p.rowsum = rowsum;
  gat_scale_kernel<<<(unsigned)ceil_div(g->n_cols * H, 256), 256, 0, st>>>(a_src, g->n_cols * H, as2);
  gat_scale_kernel<<<(unsigned)ceil_div(n * H, 256), 256, 0, st>>>(a_dst, n * H, ad2);
  RGBMP_LAUNCH_CHECK("gat_scale_kernel");
  int G, VPL;
  gat_shape(H, C, &G, &VPL);
#define ROWS_GRID(g_) ceil_div(n, GAT_THREADS / (g_))
#define ITEM_GRID(g_) p.n_items
  GAT_LAUNCH(gat_fwd_rows_kernel, ROWS_GRID, p)
  RGBMP_LAUNCH_CHECK("gat_fwd_rows_kernel");
  if (p.n_items > 0) {
    GAT_LAUNCH(gat_fwd_long_kernel, ITEM_GRID, p)
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
  int G, VPL;
  gat_shape(H, C, &G, &VPL);
  int LPH = G;                               // H == 1: every lane of the group belongs to the one head
  if (H != 1) {
    if ((C % 4) != 0 || ((C / 4) & (C / 4 - 1)) != 0)
      return fail(RGBMP_ERANGE, "rgbmp_gat_backward: needs H == 1 or C in {4,8,16,32,64,128}");
    LPH = C / (4 * VPL);                     // lanes per head (1 when a lane holds the whole head)
    if (LPH < 1) LPH = 1;
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
  Carver cv(ws, ws_bytes);
  float4* stats = cv.take<float4>((size_t)(n_dst > 0 ? n_dst : 1) * H);
  if (!carve_split(p, cv, HC) || !cv.ok()) return fail(RGBMP_EWORKSPACE, "rgbmp_gat_backward: workspace carve");
  p.stats = stats;
  if (n_dst > 0) {
    gat_bwd_prep_kernel<<<(unsigned)ceil_div(n_dst * H, 256), 256, 0, st>>>(dout, ldd, out, ldo, a_dst, rowmax, rowsum, n_dst,
                                                                           H, C, stats);
    RGBMP_LAUNCH_CHECK("gat_bwd_prep_kernel");
  }
  GAT_LAUNCH(gat_bwd_rows_kernel, ROWS_GRID, p)
  RGBMP_LAUNCH_CHECK("gat_bwd_rows_kernel");
  if (p.n_items > 0) {
    GAT_LAUNCH(gat_bwd_long_kernel, ITEM_GRID, p)
    RGBMP_LAUNCH_CHECK("gat_bwd_long_kernel");
    gat_bwd_combine_kernel<<<(unsigned)ceil_div(p.n_long * HC, 256), 256, 0, st>>>(p);
    RGBMP_LAUNCH_CHECK("gat_bwd_combine_kernel");
  }
  return 0;
}

}  // extern "C"
