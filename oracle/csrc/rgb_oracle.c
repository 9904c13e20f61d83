/* Plain-C restatement of the hot path -- TEST INFRASTRUCTURE ONLY (third statement of the oracle,
 * beside oracle/pyg_restated.py and the dense paper forms of tests/test_oracle_dense_forms.py).
 *
 * Scalar loops in EDGE ORDER: a CPU `scatter_add_` over dim 0 adds the rows of the message matrix
 * to their targets one edge after the other, so every output element is accumulated in edge order --
 * exactly what the loops below do; compiled with -ffp-contract=off (no FMA contraction: torch
 * rounds the product w*x before it adds) the float results are bit-identical to the torch oracle
 * and to the reference's own golden vectors (tests/test_oracle_c.py).
 *
 * Conventions: row = edge_index[0] = SOURCE j, col = edge_index[1] = TARGET i.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_LOOP_NONE 0
#define ORC_LOOP_ADD 1            /* add_self_loops                       graphsage.py:56       */
#define ORC_LOOP_ADD_REMAINING 2  /* add_remaining_self_loops             dagnn.py:22-23        */
#define ORC_LOOP_REMOVE_THEN_ADD 3 /* remove_self_loops + add_self_loops  graphsage.py:55-56    */

/* A1-A3: kept edges in original order, then the loops 0..N-1.  out_* hold E + N entries.
 * Returns nnz, or -1 for an id outside [0, N). */
int64_t orc_edit_loops(const int64_t* src, const int64_t* dst, int64_t E, int64_t N, int mode,
                       int64_t* out_src, int64_t* out_dst) {
  const int filter = (mode == ORC_LOOP_ADD_REMAINING || mode == ORC_LOOP_REMOVE_THEN_ADD);
  int64_t n = 0;
  for (int64_t e = 0; e < E; ++e) {
    if (src[e] < 0 || src[e] >= N || dst[e] < 0 || dst[e] >= N) return -1;
    if (filter && src[e] == dst[e]) continue;
    out_src[n] = src[e];
    out_dst[n] = dst[e];
    ++n;
  }
  if (mode != ORC_LOOP_NONE)
    for (int64_t i = 0; i < N; ++i) {
      out_src[n] = i;
      out_dst[n] = i;
      ++n;
    }
  return n;
}

/* SURVEY 8c bit-exact definition: perm = argsort(key, stable); rowptr = [0, cumsum(bincount(key))];
 * col = other[perm]; eid = perm.  A counting sort IS the stable argsort for integer keys. */
void orc_csr_build(const int64_t* key, const int64_t* other, int64_t nnz, int64_t N, int64_t* rowptr,
                   int64_t* col, int64_t* eid) {
  memset(rowptr, 0, (size_t)(N + 1) * sizeof(int64_t));
  for (int64_t e = 0; e < nnz; ++e) rowptr[key[e] + 1] += 1;
  for (int64_t i = 0; i < N; ++i) rowptr[i + 1] += rowptr[i];
  int64_t* next = (int64_t*)malloc((size_t)(N > 0 ? N : 1) * sizeof(int64_t));
  memcpy(next, rowptr, (size_t)N * sizeof(int64_t));
  for (int64_t e = 0; e < nnz; ++e) {
    const int64_t p = next[key[e]]++;
    col[p] = other[e];
    eid[p] = e;
  }
  free(next);
}

/* gcn_norm with unit input weights (models/dagnn.py:27-31): deg = scatter_add(1, col); dinv = deg^-1/2,
 * inf -> 0; w_e = dinv[row] * 1 * dinv[col] in that multiplication order.  dinv: float [N] out. */
void orc_gcn_norm_weights(const int64_t* row, const int64_t* col, int64_t nnz, int64_t N, float* dinv, float* w) {
  for (int64_t i = 0; i < N; ++i) dinv[i] = 0.0f;
  for (int64_t e = 0; e < nnz; ++e) dinv[col[e]] += 1.0f;               /* exact: integer-valued floats */
  for (int64_t i = 0; i < N; ++i) {
    const float d = 1.0f / sqrtf(dinv[i]);                              /* IEEE, = torch CPU pow(-0.5) */
    dinv[i] = isinf(d) ? 0.0f : d;
  }
  for (int64_t e = 0; e < nnz; ++e) w[e] = (dinv[row[e]] * 1.0f) * dinv[col[e]];
}

/* MessagePassing.propagate, aggr='add' (graphsage.py:58 / dagnn.py:46,57-59 / GCNConv): out[col[e]] += w[e] * x[row[e]]
 * for e = 0..nnz-1 in order; w may be NULL (message = x_j).  out: [N, F], zeroed here. */
void orc_propagate_add(const int64_t* row, const int64_t* col, const float* w, int64_t nnz, const float* x,
                       int64_t N, int F, float* out) {
  memset(out, 0, (size_t)N * F * sizeof(float));
  for (int64_t e = 0; e < nnz; ++e) {
    const float* xs = x + row[e] * F;
    float* o = out + col[e] * F;
    if (w) {
      const float we = w[e];
      for (int f = 0; f < F; ++f) {
        const float m = we * xs[f];                                     /* the message is rounded first */
        o[f] = o[f] + m;
      }
    } else {
      for (int f = 0; f < F; ++f) o[f] = o[f] + xs[f];
    }
  }
}

/* aggr='mean' (graphsage.py:39): sum, then divide by max(count, 1). */
void orc_propagate_mean(const int64_t* row, const int64_t* col, int64_t nnz, const float* x, int64_t N, int F,
                        float* out) {
  orc_propagate_add(row, col, NULL, nnz, x, N, F, out);
  float* cnt = (float*)calloc((size_t)(N > 0 ? N : 1), sizeof(float));
  for (int64_t e = 0; e < nnz; ++e) cnt[col[e]] += 1.0f;
  for (int64_t i = 0; i < N; ++i) {
    const float c = cnt[i] < 1.0f ? 1.0f : cnt[i];
    for (int f = 0; f < F; ++f) out[i * F + f] = out[i * F + f] / c;
  }
  free(cnt);
}

/* APPNP (appnp_stack.py:22): x = h; K x { x = propagate(x); x = x * (1 - alpha); x = x + alpha * h }.
 * tmp: [N, F] scratch; z: [N, F] result. */
void orc_appnp(const int64_t* row, const int64_t* col, const float* w, int64_t nnz, const float* h, int64_t N, int F,
               int K, double alpha_d, float* tmp, float* z) {
  /* torch multiplies a float tensor by the Python scalars (1 - alpha) and alpha after casting them to float */
  const float a = (float)(1.0 - alpha_d), alpha = (float)alpha_d;
  memcpy(z, h, (size_t)N * F * sizeof(float));
  for (int k = 0; k < K; ++k) {
    orc_propagate_add(row, col, w, nnz, z, N, F, tmp);
    for (int64_t t = 0; t < N * F; ++t) {
      const float s = tmp[t] * a;
      const float r = alpha * h[t];
      z[t] = s + r;
    }
  }
}

/* torch_geometric.utils.softmax over the in-edges of each target + GATConv aggregate (gat.py:18-21; SURVEY
 * A10/A11): e = leaky_relu(a_src[row] + a_dst[col]); alpha = exp(e - max_i) / (sum_i + 1e-16);
 * out[col] += alpha * xp[row].  xp: [N, H, C]; a_src, a_dst: [N, H]; alpha_out: [nnz, H]; out: [N, H, C]. */
void orc_gat_aggregate(const int64_t* row, const int64_t* col, int64_t nnz, const float* xp, const float* a_src,
                       const float* a_dst, int64_t N, int H, int C, double slope_d, float* alpha_out, float* out) {
  const float slope = (float)slope_d;
  float* mx = (float*)malloc((size_t)(N > 0 ? N : 1) * H * sizeof(float));
  float* sm = (float*)calloc((size_t)(N > 0 ? N : 1) * H, sizeof(float));
  for (int64_t t = 0; t < N * H; ++t) mx[t] = -INFINITY;
  for (int64_t e = 0; e < nnz; ++e)
    for (int h = 0; h < H; ++h) {
      float v = a_src[row[e] * H + h] + a_dst[col[e] * H + h];
      v = v > 0.0f ? v : v * slope;
      alpha_out[e * H + h] = v;
      if (v > mx[col[e] * H + h]) mx[col[e] * H + h] = v;
    }
  for (int64_t e = 0; e < nnz; ++e)
    for (int h = 0; h < H; ++h) {
      const float v = expf(alpha_out[e * H + h] - mx[col[e] * H + h]);
      alpha_out[e * H + h] = v;
      sm[col[e] * H + h] += v;
    }
  memset(out, 0, (size_t)N * H * C * sizeof(float));
  for (int64_t e = 0; e < nnz; ++e)
    for (int h = 0; h < H; ++h) {
      const float al = alpha_out[e * H + h] / (sm[col[e] * H + h] + 1e-16f);
      alpha_out[e * H + h] = al;
      const float* xs = xp + (row[e] * H + h) * C;
      float* o = out + (col[e] * H + h) * C;
      for (int c = 0; c < C; ++c) {
        const float m = xs[c] * al;
        o[c] = o[c] + m;
      }
    }
  free(mx);
  free(sm);
}

int orc_version(void) { return 1; }
