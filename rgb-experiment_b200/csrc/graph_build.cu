// Integer graph-build kernels (sm_100a): self-loop edit, stable radix sort -> CSR / transpose CSR,
// degree normalisation, long-row work lists.  Bit-exact against oracle.pyg_restated
// {edit_loops, csr_build, degree}.  All HBM-bound integer work: coalesced streaming reads,
// shared-memory digit counters, grids sized from the data (one tile per CTA).
#include <stdlib.h>
#include "common.cuh"

namespace rgbmp {

// ---------------------------------------------------------------------------------------------
// generic exclusive scan (3-phase, in place)
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename T>
__device__ __forceinline__ T warp_inclusive_scan(T v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, total via smem
template <typename T, int THREADS>
__device__ __forceinline__ T block_exclusive_scan(T v, T* total, T* smem /*THREADS/32+1*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T inc = warp_inclusive_scan(v, lane);
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    T w = (lane < THREADS / 32) ? smem[lane] : T(0);
    T winc = warp_inclusive_scan(w, lane);
    if (lane < THREADS / 32) smem[lane] = winc - w;
    if (lane == THREADS / 32 - 1) smem[THREADS / 32] = winc;
  }
  __syncthreads();
  T res = inc - v + smem[warp];
  *total = smem[THREADS / 32];
  __syncthreads();
  return res;
}

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_kernel(T* data, int64_t n, T* sums) {
  __shared__ T sm[SCAN_THREADS / 32 + 1];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  T v[SCAN_ITEMS];
  T local = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = (base + i < n) ? data[base + i] : T(0);
    local += v[i];
  }
  T total;
  T pre = block_exclusive_scan<T, SCAN_THREADS>(local, &total, sm);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) data[base + i] = pre;
    pre += v[i];
  }
  if (threadIdx.x == 0 && sums) sums[blockIdx.x] = total;
}

template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_add_kernel(T* data, int64_t n, const T* sums) {
  const T add = sums[blockIdx.x];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < n) data[base + i] += add;
}

inline size_t scan_ws_elems(int64_t n) {
  size_t tot = 0;
  while (n > 1) {
    n = ceil_div(n, SCAN_TILE);
    tot += align_up((size_t)n, 64);
    if (n == 1) break;
  }
  return tot + 64;
}

// in-place exclusive scan of data[0..n); `ws` holds scan_ws_elems(n) elements of T
template <typename T>
cudaError_t exclusive_scan(T* data, int64_t n, T* ws, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t nb = ceil_div(n, SCAN_TILE);
  if (nb == 1) {
    scan_tile_kernel<T><<<1, SCAN_THREADS, 0, st>>>(data, n, nullptr);
    return cudaGetLastError();
  }
  scan_tile_kernel<T><<<(unsigned)nb, SCAN_THREADS, 0, st>>>(data, n, ws);
  cudaError_t e = exclusive_scan<T>(ws, nb, ws + align_up((size_t)nb, 64), st);
  if (e != cudaSuccess) return e;
  scan_add_kernel<T><<<(unsigned)nb, SCAN_THREADS, 0, st>>>(data, n, ws);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// edge edit: stable compaction of non-loop edges + loop append   
// ---------------------------------------------------------------------------------------------
constexpr int EE_THREADS = 256;
constexpr int EE_ITEMS = 8;
constexpr int EE_TILE = EE_THREADS * EE_ITEMS;

__global__ void __launch_bounds__(EE_THREADS)
ee_count_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E, int64_t N,
                int filter, int32_t* __restrict__ blockcnt, int32_t* __restrict__ errflag) {
  __shared__ int32_t sm[EE_THREADS / 32 + 1];
  const int64_t base = (int64_t)blockIdx.x * EE_TILE;
  int32_t c = 0;
  bool bad = false;
#pragma unroll
  for (int i = 0; i < EE_ITEMS; ++i) {
    const int64_t e = base + (int64_t)i * EE_THREADS + threadIdx.x;
    if (e < E) {
      const int64_t s = src[e], d = dst[e];
      bad |= ((uint64_t)s >= (uint64_t)N) | ((uint64_t)d >= (uint64_t)N);
      c += (!filter || s != d) ? 1 : 0;
    }
  }
  int32_t total;
  block_exclusive_scan<int32_t, EE_THREADS>(c, &total, sm);
  if (threadIdx.x == 0) blockcnt[blockIdx.x] = total;
  if (bad) atomicOr(errflag, 1);
}

// thread t owns EE_ITEMS CONSECUTIVE edges so that the compaction is stable with one scan
__global__ void __launch_bounds__(EE_THREADS)
ee_write_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E, int filter,
                const int32_t* __restrict__ blockoff, int32_t* __restrict__ e_src, int32_t* __restrict__ e_dst) {
  __shared__ int32_t sm[EE_THREADS / 32 + 1];
  const int64_t base = (int64_t)blockIdx.x * EE_TILE + (int64_t)threadIdx.x * EE_ITEMS;
  int32_t s[EE_ITEMS], d[EE_ITEMS];
  bool keep[EE_ITEMS];
  int32_t c = 0;
#pragma unroll
  for (int i = 0; i < EE_ITEMS; ++i) {
    const int64_t e = base + i;
    keep[i] = false;
    if (e < E) {
      s[i] = (int32_t)src[e];
      d[i] = (int32_t)dst[e];
      keep[i] = (!filter || s[i] != d[i]);
    }
    c += keep[i] ? 1 : 0;
  }
  int32_t total;
  int32_t pos = blockoff[blockIdx.x] + block_exclusive_scan<int32_t, EE_THREADS>(c, &total, sm);
#pragma unroll
  for (int i = 0; i < EE_ITEMS; ++i) {
    if (keep[i]) {
      e_src[pos] = s[i];
      e_dst[pos] = d[i];
      ++pos;
    }
  }
}

// v2 of the write: coalesced loads (lane-consecutive edges, EE_ITEMS rounds), ballot ranks give every kept edge its
// stable slot inside the tile, the tile is compacted in shared memory and leaves as one contiguous burst.
__global__ void __launch_bounds__(EE_THREADS)
ee_write2_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E, int filter,
                 const int32_t* __restrict__ blockoff, int32_t* __restrict__ e_src, int32_t* __restrict__ e_dst) {
  constexpr int W = EE_THREADS / 32;
  __shared__ int32_t woff[EE_ITEMS * W];          // kept edges per (round, warp) -> exclusive offsets
  __shared__ int32_t ss[EE_TILE], sd[EE_TILE];
  __shared__ int32_t total_sm;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * EE_TILE;
  int32_t s[EE_ITEMS], d[EE_ITEMS];
  uint32_t bal[EE_ITEMS];
#pragma unroll
  for (int i = 0; i < EE_ITEMS; ++i) {
    const int64_t e = base + (int64_t)i * EE_THREADS + threadIdx.x;
    bool keep = false;
    s[i] = d[i] = 0;
    if (e < E) {
      s[i] = (int32_t)src[e];
      d[i] = (int32_t)dst[e];
      keep = (!filter || s[i] != d[i]);
    }
    bal[i] = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) woff[i * W + warp] = __popc(bal[i]);
  }
  __syncthreads();
  if (warp == 0) {                                 // exclusive scan of the EE_ITEMS * W (= 64) counters, 2 per lane
    static_assert(EE_ITEMS * W == 64, "scan below assumes 64 counters");
    const int32_t a = woff[2 * lane], b = woff[2 * lane + 1];
    const int32_t inc = warp_inclusive_scan<int32_t>(a + b, lane);
    woff[2 * lane] = inc - a - b;
    woff[2 * lane + 1] = inc - b;
    if (lane == 31) total_sm = inc;
  }
  __syncthreads();
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < EE_ITEMS; ++i) {
    if ((bal[i] >> lane) & 1u) {
      const int32_t p = woff[i * W + warp] + __popc(bal[i] & lt);
      ss[p] = s[i];
      sd[p] = d[i];
    }
  }
  __syncthreads();
  const int32_t total = total_sm;
  const int64_t out0 = blockoff[blockIdx.x];
  for (int j = threadIdx.x; j < total; j += EE_THREADS) {
    e_src[out0 + j] = ss[j];
    e_dst[out0 + j] = sd[j];
  }
}

// the count kernel must see the same thread->edge mapping only in aggregate (per block), so the
// strided mapping there is fine; both kernels cover edges [block*TILE, (block+1)*TILE).

__global__ void ee_loops_kernel(int64_t N, int add_loops, const int32_t* __restrict__ blockoff, int64_t nblocks,
                                const int32_t* __restrict__ lastcnt_src, int64_t E_fixed,
                                int32_t* __restrict__ e_src, int32_t* __restrict__ e_dst,
                                int64_t* __restrict__ nnz_dev, const int32_t* __restrict__ errflag) {
  // kept = (filter path) blockoff[nblocks-1] + lastcnt  |  (no filter) E_fixed
  int64_t kept = E_fixed;
  if (blockoff) kept = (nblocks > 0) ? (int64_t)blockoff[nblocks - 1] + (int64_t)lastcnt_src[0] : 0;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (add_loops && i < N) {
    e_src[kept + i] = (int32_t)i;
    e_dst[kept + i] = (int32_t)i;
  }
  if (i == 0) nnz_dev[0] = (*errflag) ? -1 : kept + (add_loops ? N : 0);
}

__global__ void ee_save_last_kernel(const int32_t* blockcnt, int64_t nblocks, int32_t* last) {
  if (threadIdx.x == 0 && blockIdx.x == 0) last[0] = nblocks > 0 ? blockcnt[nblocks - 1] : 0;
}

// ---------------------------------------------------------------------------------------------
// stable LSD radix sort (8-bit digits) of (key, value=edge id)
// ---------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096 keys per CTA
constexpr int RS_BINS = 256;

__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const int32_t* __restrict__ keys, int64_t n, int shift, int64_t nblocks,
               int32_t* __restrict__ blockhist /*[256][nblocks]*/) {
  __shared__ int32_t h[RS_BINS];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
  int32_t kk[RS_ITEMS];
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {                 // all loads in flight before the first shared-memory atomic
    const int64_t k = base + (int64_t)i * RS_THREADS + threadIdx.x;
    kk[i] = (k < n) ? __ldcs(keys + k) : 0;
  }
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    const int64_t k = base + (int64_t)i * RS_THREADS + threadIdx.x;
    if (k < n) atomicAdd(&h[((uint32_t)kk[i] >> shift) & 0xFF], 1);
  }
  __syncthreads();
  blockhist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// Warp w ranks the contiguous slice [w*512, (w+1)*512) of the tile in 16 rounds of 32 keys, so the
// (round, lane) order is the input order: ranks are stable.
template <bool FIRST>
__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const int32_t* __restrict__ keys_in, const int32_t* __restrict__ vals_in, int64_t n, int shift,
                  int64_t nblocks, const int32_t* __restrict__ blockoff /*scanned [256][nblocks]*/,
                  int32_t* __restrict__ keys_out, int32_t* __restrict__ vals_out) {
  __shared__ int32_t cnt[RS_WARPS][RS_BINS];
  __shared__ int32_t gbase[RS_BINS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) cnt[w][threadIdx.x] = 0;
  __syncthreads();

  const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * (32 * RS_ITEMS);
  int32_t key[RS_ITEMS], rank[RS_ITEMS];
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int64_t k = wbase + r * 32 + lane;
    const bool valid = k < n;
    key[r] = valid ? keys_in[k] : 0;
    const uint32_t d = valid ? (((uint32_t)key[r] >> shift) & 0xFF) : 0xFFFFFFFFu;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(peers) - 1;
    int32_t old = 0;
    if (valid && lane == leader) {
      old = cnt[warp][d];
      cnt[warp][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[r] = old + __popc(peers & lt);
    __syncwarp();
  }
  __syncthreads();
  {  // thread d: exclusive scan over warps for digit d; fetch the tile's global base
    const int d = threadIdx.x;
    int32_t off = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      const int32_t c = cnt[w][d];
      cnt[w][d] = off;
      off += c;
    }
    gbase[d] = blockoff[(int64_t)d * nblocks + blockIdx.x];
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int64_t k = wbase + r * 32 + lane;
    if (k < n) {
      const uint32_t d = ((uint32_t)key[r] >> shift) & 0xFF;
      const int32_t pos = gbase[d] + cnt[warp][d] + rank[r];
      keys_out[pos] = key[r];
      vals_out[pos] = FIRST ? (int32_t)k : vals_in[k];
    }
  }
}

// Lanes of the warp that hold the same 8-bit digit (and the same validity) as the caller.  Eight ballots + LOP3s on
// the ALU pipe instead of one match.any: the ncu profile of the scatter (profiles/r01_build_scatter3_ncu_full.txt)
// showed the ADU pipe, where MATCH executes, 57-60 % busy at 37 % occupancy -- the kernel's limiter.
__device__ __forceinline__ uint32_t digit_peers(uint32_t d, bool valid) {
  const uint32_t vb = __ballot_sync(0xffffffffu, valid);
  uint32_t peers = valid ? vb : ~vb;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const bool bit = (d >> b) & 1u;
    const uint32_t bal = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? bal : ~bal;
  }
  return peers;
}

// v2 of the scatter: the tile is first reordered by digit in shared memory (same stable ranks), then written out
// by consecutive threads -- every (tile, digit) run becomes one contiguous burst instead of 4-byte stores spread over
// up to 32 bins per warp instruction.  Output positions are identical to v1 (bit-exact, stable).
template <bool FIRST, bool PAY>
__global__ void __launch_bounds__(RS_THREADS, PAY ? 3 : 4)
rs_scatter2_kernel(const int32_t* __restrict__ keys_in, const int32_t* __restrict__ vals_in, int64_t n, int shift,
                   int64_t nblocks, const int32_t* __restrict__ blockoff /*scanned [256][nblocks]*/,
                   int32_t* __restrict__ keys_out, int32_t* __restrict__ vals_out,
                   const int32_t* __restrict__ pay_in, int32_t* __restrict__ pay_out) {
  __shared__ int32_t cnt[RS_WARPS][RS_BINS];
  __shared__ int32_t dstart[RS_BINS];   // tile-local start of digit d
  __shared__ int32_t gdelta[RS_BINS];   // global base of digit d for this tile - dstart[d]
  __shared__ int32_t skey[RS_TILE];
  __shared__ int32_t sval[RS_TILE];
  __shared__ uint8_t sdig[PAY ? RS_TILE : 4];   // digit of the element at sorted slot j (payload pass only)
  __shared__ int32_t scan_sm[RS_THREADS / 32 + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) cnt[w][threadIdx.x] = 0;
  __syncthreads();

  const int64_t tbase = (int64_t)blockIdx.x * RS_TILE;
  const int64_t wbase = tbase + (int64_t)warp * (32 * RS_ITEMS);
  int32_t key[RS_ITEMS], rank[RS_ITEMS];
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int64_t k = wbase + r * 32 + lane;
    const bool valid = k < n;
    key[r] = valid ? keys_in[k] : 0;
    const uint32_t d = ((uint32_t)key[r] >> shift) & 0xFF;
    const uint32_t peers = digit_peers(d, valid);
    const int leader = __ffs(peers) - 1;
    int32_t old = 0;
    if (valid && lane == leader) {
      old = cnt[warp][d];
      cnt[warp][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[r] = old + __popc(peers & lt);
    __syncwarp();
  }
  __syncthreads();
  int32_t tot = 0;
  {  // thread d: exclusive scan over warps for digit d
    const int d = threadIdx.x;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      const int32_t c = cnt[w][d];
      cnt[w][d] = tot;
      tot += c;
    }
  }
  int32_t tile_total;
  const int32_t ds = block_exclusive_scan<int32_t, RS_THREADS>(tot, &tile_total, scan_sm);
  dstart[threadIdx.x] = ds;
  gdelta[threadIdx.x] = blockoff[(int64_t)threadIdx.x * nblocks + blockIdx.x] - ds;
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int64_t k = wbase + r * 32 + lane;
    if (k < n) {
      const uint32_t d = ((uint32_t)key[r] >> shift) & 0xFF;
      const int32_t lp = dstart[d] + cnt[warp][d] + rank[r];
      skey[lp] = key[r];
      sval[lp] = FIRST ? (int32_t)k : vals_in[k];
      if (PAY) sdig[lp] = (uint8_t)d;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    const int j = i * RS_THREADS + threadIdx.x;
    if (j < tile_total) {
      const int32_t kk = skey[j];
      const int32_t pos = gdelta[((uint32_t)kk >> shift) & 0xFF] + j;
      keys_out[pos] = kk;
      vals_out[pos] = sval[j];
    }
  }
  if (!PAY) return;
  // optional second payload (the CSR's column ids) rides along through the same slots: it is read at the element's
  // CURRENT position -- in the first pass that is the original edge order, a coalesced stream -- so that no pass
  // ever gathers it through the permutation (the old col = other[eid] gather moved 113 DRAM bytes per 4-byte id)
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int64_t k = wbase + r * 32 + lane;
    if (k < n) {
      const uint32_t d = ((uint32_t)key[r] >> shift) & 0xFF;
      skey[dstart[d] + cnt[warp][d] + rank[r]] = pay_in[k];
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    const int j = i * RS_THREADS + threadIdx.x;
    if (j < tile_total) pay_out[gdelta[sdig[j]] + j] = skey[j];
  }
}

// v3 of the scatter, CSR build only: the two values that follow a key -- its edge id and the column id -- travel as
// ONE 8-byte slot (a single shared-memory exchange, 8-byte stores, no second round); the first pass creates the
// slot from the element's position and a coalesced read of `other`, the last pass splits it into eid[] and col[].
// 58 KB of dynamic shared memory, 3 CTAs per SM.
constexpr size_t RS3_SMEM = (size_t)RS_WARPS * RS_BINS * 4 + 2 * RS_BINS * 4 + (size_t)RS_TILE * 4 + (size_t)RS_TILE * 8 + 64;

template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(RS_THREADS, 3)
rs_scatter3_kernel(const int32_t* __restrict__ keys_in, const int2* __restrict__ pay_in,
                   const int32_t* __restrict__ other_first, int64_t n, int shift, int64_t nblocks,
                   const int32_t* __restrict__ blockoff /*scanned [256][nblocks]*/, int32_t* __restrict__ keys_out,
                   int2* __restrict__ pay_out, int32_t* __restrict__ eid_out, int32_t* __restrict__ col_out) {
  extern __shared__ __align__(16) unsigned char rs3_raw[];
  int2* spay = reinterpret_cast<int2*>(rs3_raw);                                  // [RS_TILE]
  int32_t* skey = reinterpret_cast<int32_t*>(rs3_raw + (size_t)RS_TILE * 8);      // [RS_TILE]
  int32_t(*cnt)[RS_BINS] = reinterpret_cast<int32_t(*)[RS_BINS]>(skey + RS_TILE); // [RS_WARPS][RS_BINS]
  int32_t* dstart = &cnt[0][0] + RS_WARPS * RS_BINS;                              // [RS_BINS]
  int32_t* gdelta = dstart + RS_BINS;                                             // [RS_BINS]
  int32_t* scan_sm = gdelta + RS_BINS;                                            // [RS_THREADS/32 + 1]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int w = 0; w < RS_WARPS; ++w) cnt[w][threadIdx.x] = 0;
  __syncthreads();

  const int64_t tbase = (int64_t)blockIdx.x * RS_TILE;
  const int64_t wbase = tbase + (int64_t)warp * (32 * RS_ITEMS);
  int32_t key[RS_ITEMS], rank[RS_ITEMS];
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int64_t k = wbase + r * 32 + lane;
    const bool valid = k < n;
    key[r] = valid ? keys_in[k] : 0;
    const uint32_t d = ((uint32_t)key[r] >> shift) & 0xFF;
    const uint32_t peers = digit_peers(d, valid);
    const int leader = __ffs(peers) - 1;
    int32_t old = 0;
    if (valid && lane == leader) {
      old = cnt[warp][d];
      cnt[warp][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[r] = old + __popc(peers & lt);
    __syncwarp();
  }
  __syncthreads();
  int32_t tot = 0;
  {
    const int d = threadIdx.x;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      const int32_t c = cnt[w][d];
      cnt[w][d] = tot;
      tot += c;
    }
  }
  int32_t tile_total;
  const int32_t ds = block_exclusive_scan<int32_t, RS_THREADS>(tot, &tile_total, scan_sm);
  dstart[threadIdx.x] = ds;
  gdelta[threadIdx.x] = blockoff[(int64_t)threadIdx.x * nblocks + blockIdx.x] - ds;
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int64_t k = wbase + r * 32 + lane;
    if (k < n) {
      const uint32_t d = ((uint32_t)key[r] >> shift) & 0xFF;
      const int32_t lp = dstart[d] + cnt[warp][d] + rank[r];
      skey[lp] = key[r];
      spay[lp] = FIRST ? make_int2((int32_t)k, other_first[k]) : pay_in[k];
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    const int j = i * RS_THREADS + threadIdx.x;
    if (j < tile_total) {
      const int32_t kk = skey[j];
      const int32_t pos = gdelta[((uint32_t)kk >> shift) & 0xFF] + j;
      const int2 pv = spay[j];
      keys_out[pos] = kk;
      if (LAST) {
        eid_out[pos] = pv.x;
        col_out[pos] = pv.y;
      } else {
        pay_out[pos] = pv;
      }
    }
  }
}

// rowptr[r] = number of sorted keys < r = lower_bound(sorted, r), r in [0, N]: no atomics, no scan; neighbouring
// rows walk nearly the same search path, so the probes are cache hits.
__global__ void __launch_bounds__(256)
rowptr_lower_bound_kernel(const int32_t* __restrict__ sorted, int64_t nnz, int64_t N, int64_t* __restrict__ rowptr) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > N) return;
  int64_t lo = 0, hi = nnz;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)__ldg(sorted + mid) < r) lo = mid + 1;
    else hi = mid;
  }
  rowptr[r] = lo;
}

// RGBMP_BUILD_VARIANT (read once): 1 = first-round kernels (direct scatter, atomic degree histogram + scan,
// col = other[eid] gather), 2 = shared-memory reorder with the column ids as a separate payload round,
// default 3 = packed (eid, col) slots in the CSR sort.  All three are bit-identical; 1 and 2 stay for A/B timing.
static int build_variant() {
  static const int v = [] {
    const char* e = getenv("RGBMP_BUILD_VARIANT");
    if (e && e[1] == 0 && (e[0] == '1' || e[0] == '2')) return e[0] - '0';
    return 3;
  }();
  return v;
}

__global__ void deg_hist_kernel(const int32_t* __restrict__ key, int64_t nnz, unsigned long long* __restrict__ deg) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride)
    atomicAdd(&deg[key[k]], 1ull);
}

__global__ void gather_i32_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ idx, int64_t n,
                                  int32_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) out[k] = src[idx[k]];
}

__global__ void iota_i32_kernel(int32_t* out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) out[k] = (int32_t)k;
}

// ---------------------------------------------------------------------------------------------
// node / edge normalisation
// ---------------------------------------------------------------------------------------------
__global__ void degree_norm_kernel(const int64_t* __restrict__ rowptr, int64_t N, int mode, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float deg = (float)(rowptr[i + 1] - rowptr[i]);
  float v;
  if (mode == RGBMP_NORM_INV_SQRT) {
    v = (deg > 0.f) ? __fdiv_rn(1.0f, __fsqrt_rn(deg)) : 0.f;  // IEEE 1/sqrt == torch CPU pow(-0.5); inf -> 0
  } else if (mode == RGBMP_NORM_INV_MEAN) {
    v = __fdiv_rn(1.0f, fmaxf(deg, 1.0f));
  } else {
    v = fmaxf(deg, 1.0f);
  }
  out[i] = v;
}

__global__ void gcn_edge_weight_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                       int64_t n_rows, const float* __restrict__ dinv_row,
                                       const float* __restrict__ dinv_col, float* __restrict__ val) {
  // one warp per row
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const float di = dinv_row[row];
  const int64_t s = rowptr[row], e = rowptr[row + 1];
  for (int64_t k = s + lane; k < e; k += 32)
    val[k] = __fmul_rn(__fmul_rn(dinv_col[col[k]], 1.0f), di);  // dinv[src]*w*dinv[dst], w = 1
}

__global__ void edge_permute_kernel(const float* __restrict__ in, const int32_t* __restrict__ eid, int64_t nnz, int H,
                                    int scatter, float* __restrict__ out) {
  const int64_t total = nnz * H;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t k = t / H;
    const int h = (int)(t - k * H);
    const int64_t e = eid[k];
    if (scatter) out[e * H + h] = in[t];
    else out[t] = in[e * H + h];
  }
}

// ---------------------------------------------------------------------------------------------
// long-row work lists
// ---------------------------------------------------------------------------------------------
__global__ void longrow_count_kernel(const int64_t* __restrict__ rowptr, int64_t n_rows, int32_t chunk,
                                     int32_t long_chunk, unsigned long long* __restrict__ counts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long nl = 0, ni = 0;
  if (i < n_rows) {
    const int64_t deg = rowptr[i + 1] - rowptr[i];
    if (deg > chunk) {
      nl = 1;
      ni = (unsigned long long)((deg + long_chunk - 1) / long_chunk);
    }
  }
  // warp-aggregate before the global atomics
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nl += __shfl_down_sync(0xffffffffu, nl, o);
    ni += __shfl_down_sync(0xffffffffu, ni, o);
  }
  if ((threadIdx.x & 31) == 0 && nl) {
    atomicAdd(&counts[0], nl);
    atomicAdd(&counts[1], ni);
  }
}

// position p of the list order <-> row order[p] (order == NULL: natural order).  Listing the long rows in the
// order of the row schedule keeps the work items of one locality group adjacent in the long-row launch.
__global__ void longrow_flags_kernel(const int64_t* __restrict__ rowptr, int64_t n_rows, int32_t chunk,
                                     int32_t long_chunk, const int32_t* __restrict__ order, int32_t* __restrict__ slot,
                                     int32_t* __restrict__ items) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_rows) return;
  const int64_t i = order ? order[p] : p;
  const int64_t deg = rowptr[i + 1] - rowptr[i];
  const bool lg = deg > chunk;
  slot[p] = lg ? 1 : 0;
  items[p] = lg ? (int32_t)((deg + long_chunk - 1) / long_chunk) : 0;
}

__global__ void longrow_fill_kernel(const int64_t* __restrict__ rowptr, int64_t n_rows, int32_t chunk,
                                    int32_t long_chunk, const int32_t* __restrict__ order, const int32_t* __restrict__ slot,
                                    const int32_t* __restrict__ items, int64_t n_long, int64_t n_items,
                                    int32_t* __restrict__ long_rows, int32_t* __restrict__ long_item_ptr,
                                    int32_t* __restrict__ item_long, int64_t* __restrict__ item_start) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p == 0) long_item_ptr[n_long] = (int32_t)n_items;
  if (p >= n_rows) return;
  const int64_t i = order ? order[p] : p;
  const int64_t s = rowptr[i], deg = rowptr[i + 1] - s;
  if (deg <= chunk) return;
  const int32_t sl = slot[p], it0 = items[p];
  long_rows[sl] = (int32_t)i;
  long_item_ptr[sl] = it0;
  const int32_t cnt = (int32_t)((deg + long_chunk - 1) / long_chunk);
  for (int32_t j = 0; j < cnt; ++j) {
    item_long[it0 + j] = sl;
    item_start[it0 + j] = s + (int64_t)j * long_chunk;
  }
}

// stable LSD radix sort of (key, index) pairs on the low `bits` bits of the key; the permutation
// (original index of every sorted element) lands in vfinal.  kA/kB/vA: scratch of n int32 each.
static int sort_pairs_i32(const int32_t* key, int64_t n, int bits, int32_t* kA, int32_t* kB, int32_t* vA,
                          int32_t* vfinal, int32_t* bh, int32_t* sc32, int64_t nb, cudaStream_t st,
                          const int32_t** sorted_keys = nullptr, const int32_t* pay = nullptr, int32_t* pA = nullptr,
                          int32_t* pB = nullptr, int32_t* pay_final = nullptr) {
  // pay (optional, v2 only): a second int32 per element that travels with the pair; the last pass writes pay_final
  const int passes = (bits + 7) / 8;
  const bool v2 = build_variant() != 1;
  const int32_t* kin = key;
  const int32_t* vin = nullptr;
  const int32_t* pin = pay;
  for (int p = 0; p < passes; ++p) {
    const int shift = 8 * p;
    int32_t* kout = (p & 1) ? kB : kA;
    int32_t* vout = (((passes - 1 - p) & 1) == 0) ? vfinal : vA;   // the LAST pass writes vfinal
    int32_t* pout = !pay ? nullptr : (p == passes - 1) ? pay_final : ((p & 1) ? pB : pA);
    rs_hist_kernel<<<(unsigned)nb, RS_THREADS, 0, st>>>(kin, n, shift, nb, bh);
    RGBMP_LAUNCH_CHECK("rs_hist_kernel");
    RGBMP_CUDA(exclusive_scan<int32_t>(bh, (int64_t)RS_BINS * nb, sc32, st));
    if (p == 0) {
      if (v2 && pay) rs_scatter2_kernel<true, true><<<(unsigned)nb, RS_THREADS, 0, st>>>(kin, vin, n, shift, nb, bh, kout, vout, pin, pout);
      else if (v2) rs_scatter2_kernel<true, false><<<(unsigned)nb, RS_THREADS, 0, st>>>(kin, vin, n, shift, nb, bh, kout, vout, nullptr, nullptr);
      else rs_scatter_kernel<true><<<(unsigned)nb, RS_THREADS, 0, st>>>(kin, vin, n, shift, nb, bh, kout, vout);
    } else {
      if (v2 && pay) rs_scatter2_kernel<false, true><<<(unsigned)nb, RS_THREADS, 0, st>>>(kin, vin, n, shift, nb, bh, kout, vout, pin, pout);
      else if (v2) rs_scatter2_kernel<false, false><<<(unsigned)nb, RS_THREADS, 0, st>>>(kin, vin, n, shift, nb, bh, kout, vout, nullptr, nullptr);
      else rs_scatter_kernel<false><<<(unsigned)nb, RS_THREADS, 0, st>>>(kin, vin, n, shift, nb, bh, kout, vout);
    }
    RGBMP_LAUNCH_CHECK("rs_scatter_kernel");
    kin = kout;
    vin = vout;
    pin = pout;
  }
  if (sorted_keys) *sorted_keys = kin;
  return 0;
}

template <bool FIRST, bool LAST>
static cudaError_t launch_scatter3(const int32_t* kin, const int2* pin, const int32_t* other, int64_t n, int shift, int64_t nb,
                                   const int32_t* bh, int32_t* kout, int2* pout, int32_t* eid, int32_t* col, cudaStream_t st) {
  auto fn = rs_scatter3_kernel<FIRST, LAST>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS3_SMEM);
  if (e != cudaSuccess) return e;
  fn<<<(unsigned)nb, RS_THREADS, RS3_SMEM, st>>>(kin, pin, other, n, shift, nb, bh, kout, pout, eid, col);
  return cudaGetLastError();
}

// CSR sort with packed (eid, col) slots: sorts `key` (stable, low `bits` bits), writes eid[] and col[] = other[eid[]];
// *sorted_keys = the sorted keys (in kA or kB).  pA/pB: n int2 each.
static int sort_csr_packed(const int32_t* key, const int32_t* other, int64_t n, int bits, int32_t* kA, int32_t* kB,
                           int2* pA, int2* pB, int32_t* eid, int32_t* col, int32_t* bh, int32_t* sc32, int64_t nb,
                           cudaStream_t st, const int32_t** sorted_keys) {
  const int passes = (bits + 7) / 8;
  const int32_t* kin = key;
  const int2* pin = nullptr;
  for (int p = 0; p < passes; ++p) {
    const int shift = 8 * p;
    const bool first = p == 0, last = p == passes - 1;
    int32_t* kout = (p & 1) ? kB : kA;
    int2* pout = (p & 1) ? pB : pA;
    rs_hist_kernel<<<(unsigned)nb, RS_THREADS, 0, st>>>(kin, n, shift, nb, bh);
    RGBMP_LAUNCH_CHECK("rs_hist_kernel");
    RGBMP_CUDA(exclusive_scan<int32_t>(bh, (int64_t)RS_BINS * nb, sc32, st));
    cudaError_t e;
    if (first && last) e = launch_scatter3<true, true>(kin, pin, other, n, shift, nb, bh, kout, pout, eid, col, st);
    else if (first) e = launch_scatter3<true, false>(kin, pin, other, n, shift, nb, bh, kout, pout, eid, col, st);
    else if (last) e = launch_scatter3<false, true>(kin, pin, other, n, shift, nb, bh, kout, pout, eid, col, st);
    else e = launch_scatter3<false, false>(kin, pin, other, n, shift, nb, bh, kout, pout, eid, col, st);
    if (e != cudaSuccess) return cuda_fail(e, "rs_scatter3_kernel");
    kin = kout;
    pin = pout;
  }
  *sorted_keys = kin;
  return 0;
}

__global__ void row_key_kernel(const int64_t* __restrict__ rowptr, int64_t n_rows, int64_t window,
                               const int32_t* __restrict__ group, int32_t* __restrict__ keys) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  int64_t deg = rowptr[i + 1] - rowptr[i];
  if (deg > 0xFFFF) deg = 0xFFFF;
  const int64_t major = group ? (int64_t)group[i] : i / window;   // locality group, or window of consecutive ids
  keys[i] = (int32_t)((major << 16) | (0xFFFF - deg));            // major first, longest rows first inside
}

__global__ void __launch_bounds__(256)
col_freq_kernel(const int32_t* __restrict__ col, int64_t nnz, int32_t* __restrict__ freq) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride)
    atomicAdd(freq + (col[k] & 0x7fffffff), 1);
}

__global__ void __launch_bounds__(256)
col_tag_kernel(const int32_t* __restrict__ col, int64_t nnz, const int32_t* __restrict__ freq, int32_t thresh,
               int32_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) {
    const int32_t c = col[k] & 0x7fffffff;
    out[k] = (__ldg(freq + c) >= thresh) ? (int32_t)((uint32_t)c | 0x80000000u) : c;
  }
}

// ---------------------------------------------------------------------------------------------
// coalesce / to_undirected: sort by (row, col), drop duplicates   
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
co_expand_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t E, int64_t N, int symmetrize,
                 int32_t* __restrict__ r, int32_t* __restrict__ c, int32_t* __restrict__ errflag) {
  const int64_t n = symmetrize ? 2 * E : E;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const bool rev = i >= E;
    const int64_t e = rev ? i - E : i;
    const int64_t a = src[e], b = dst[e];
    bad |= ((uint64_t)a >= (uint64_t)N) | ((uint64_t)b >= (uint64_t)N);
    r[i] = (int32_t)(rev ? b : a);
    c[i] = (int32_t)(rev ? a : b);
  }
  if (bad) atomicOr(errflag, 1);
}

__global__ void __launch_bounds__(256)
co_flag_kernel(const int32_t* __restrict__ rs, const int32_t* __restrict__ cs, int64_t n, int32_t* __restrict__ flag) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    flag[i] = (i == 0 || rs[i] != rs[i - 1] || cs[i] != cs[i - 1]) ? 1 : 0;
}

// pos = exclusive scan of the first-of-run flags; element i is a run head iff pos[i+1] != pos[i]
__global__ void __launch_bounds__(256)
co_write_kernel(const int32_t* __restrict__ rs, const int32_t* __restrict__ cs, int64_t n, const int32_t* __restrict__ pos,
                const int32_t* __restrict__ errflag, int64_t* __restrict__ out_src, int64_t* __restrict__ out_dst,
                int64_t* __restrict__ count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t last = n - 1;
  const bool last_head = (n == 1) || rs[last] != rs[last - 1] || cs[last] != cs[last - 1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const bool head = (i == 0) || rs[i] != rs[i - 1] || cs[i] != cs[i - 1];
    if (head) {
      out_src[pos[i]] = rs[i];
      out_dst[pos[i]] = cs[i];
    }
    if (i == last) *count = (*errflag) ? -1 : (int64_t)pos[last] + (last_head ? 1 : 0);
  }
}

}  // namespace rgbmp

using namespace rgbmp;

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int rgbmp_version(void) { return RGBMP_VERSION; }
const char* rgbmp_last_error(void) { return err_buf(); }

size_t rgbmp_edge_edit_workspace_bytes(int64_t E, int64_t N) {
  (void)N;
  const size_t nb = (size_t)ceil_div(E > 0 ? E : 1, EE_TILE);
  return (align_up(nb, 64) + scan_ws_elems((int64_t)nb) + 128) * sizeof(int32_t) + 1024;
}

int rgbmp_edge_edit(const int64_t* src, const int64_t* dst, int64_t E, int64_t N, int loop_mode, int32_t* e_src,
                    int32_t* e_dst, int64_t* nnz_dev, void* ws, size_t ws_bytes, int device, void* stream) {
  if (E < 0 || N < 0 || (E > 0 && (!src || !dst)) || !e_src || !e_dst || !nnz_dev || !ws)
    return fail(RGBMP_EINVAL, "rgbmp_edge_edit: null pointer or negative size");
  if (loop_mode < 0 || loop_mode > 3) return fail(RGBMP_EINVAL, "rgbmp_edge_edit: bad loop_mode %d", loop_mode);
  if (N >= (1ll << 31) - 1 || E + N >= (1ll << 31) - 1)
    return fail(RGBMP_ERANGE, "rgbmp_edge_edit: N=%lld, E+N=%lld exceed the int32 id range", (long long)N,
                (long long)(E + N));
  if (ws_bytes < rgbmp_edge_edit_workspace_bytes(E, N)) return fail(RGBMP_EWORKSPACE, "rgbmp_edge_edit: workspace");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RGBMP_EINVAL, "rgbmp_edge_edit: bad device %d", device);
  cudaStream_t st = (cudaStream_t)stream;
  const int filter = (loop_mode == RGBMP_LOOP_ADD_REMAINING || loop_mode == RGBMP_LOOP_REMOVE_THEN_ADD) ? 1 : 0;
  const int add_loops = loop_mode != RGBMP_LOOP_NONE;
  const int64_t nb = ceil_div(E, EE_TILE);
  Carver cv(ws, ws_bytes);
  int32_t* errflag = cv.take<int32_t>(1);
  int32_t* last = cv.take<int32_t>(1);
  int32_t* blockcnt = cv.take<int32_t>(align_up((size_t)(nb > 0 ? nb : 1), 64));
  int32_t* scanws = cv.take<int32_t>(scan_ws_elems(nb > 0 ? nb : 1));
  RGBMP_CUDA(cudaMemsetAsync(errflag, 0, sizeof(int32_t), st));
  if (nb > 0) {
    ee_count_kernel<<<(unsigned)nb, EE_THREADS, 0, st>>>(src, dst, E, N, filter, blockcnt, errflag);
    RGBMP_LAUNCH_CHECK("ee_count_kernel");
    ee_save_last_kernel<<<1, 32, 0, st>>>(blockcnt, nb, last);
    RGBMP_CUDA(exclusive_scan<int32_t>(blockcnt, nb, scanws, st));
    if (build_variant() != 1) ee_write2_kernel<<<(unsigned)nb, EE_THREADS, 0, st>>>(src, dst, E, filter, blockcnt, e_src, e_dst);
    else ee_write_kernel<<<(unsigned)nb, EE_THREADS, 0, st>>>(src, dst, E, filter, blockcnt, e_src, e_dst);
    RGBMP_LAUNCH_CHECK("ee_write_kernel");
  } else {
    RGBMP_CUDA(cudaMemsetAsync(last, 0, sizeof(int32_t), st));
  }
  const int64_t lb = ceil_div(N > 0 ? N : 1, 256);
  ee_loops_kernel<<<(unsigned)lb, 256, 0, st>>>(N, add_loops, nb > 0 ? blockcnt : nullptr, nb, last, 0, e_src, e_dst,
                                                nnz_dev, errflag);
  RGBMP_LAUNCH_CHECK("ee_loops_kernel");
  return 0;
}

size_t rgbmp_csr_build_workspace_bytes(int64_t nnz, int64_t N) {
  (void)N;
  const size_t n = (size_t)(nnz > 0 ? nnz : 1);
  const size_t nb = (size_t)ceil_div((int64_t)n, RS_TILE);
  size_t b = 0;
  b += 2 * align_up(n * sizeof(int32_t), 256) + 2 * align_up(2 * n * sizeof(int32_t), 256);   // keys A/B, (eid, col) slots A/B
  b += align_up(RS_BINS * nb * sizeof(int32_t), 256);                              // block histograms
  b += align_up(scan_ws_elems((int64_t)(RS_BINS * nb)) * sizeof(int32_t), 256);    // scan scratch (i32)
  b += align_up(scan_ws_elems(N + 1) * sizeof(int64_t), 256);                      // scan scratch (i64)
  return b + 4096;
}

int rgbmp_csr_build(const int32_t* key, const int32_t* other, int64_t nnz, int64_t N, int64_t* rowptr, int32_t* col,
                    int32_t* eid, void* ws, size_t ws_bytes, int device, void* stream) {
  if (nnz < 0 || N < 0 || !rowptr || (nnz > 0 && (!key || !other || !col || !eid)) || !ws)
    return fail(RGBMP_EINVAL, "rgbmp_csr_build: null pointer or negative size");
  if (nnz >= (1ll << 31) - 1 || N >= (1ll << 31) - 1)
    return fail(RGBMP_ERANGE, "rgbmp_csr_build: nnz=%lld or N=%lld exceeds int32", (long long)nnz, (long long)N);
  if (ws_bytes < rgbmp_csr_build_workspace_bytes(nnz, N)) return fail(RGBMP_EWORKSPACE, "rgbmp_csr_build: workspace");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RGBMP_EINVAL, "rgbmp_csr_build: bad device %d", device);
  cudaStream_t st = (cudaStream_t)stream;

  const size_t n = (size_t)(nnz > 0 ? nnz : 1);
  const int64_t nb = ceil_div((int64_t)n, RS_TILE);
  Carver cv(ws, ws_bytes);
  int32_t* kA = cv.take<int32_t>(n);
  int32_t* kB = cv.take<int32_t>(n);
  int32_t* vA = cv.take<int32_t>(2 * n);   // variants 1/2: vals A in the first half, payload A in the second; variant 3: packed slots A
  int32_t* pA = vA + n;
  int32_t* pB = cv.take<int32_t>(2 * n);   // variant 2: payload B; variant 3: packed slots B
  int32_t* bh = cv.take<int32_t>((size_t)RS_BINS * nb);
  int32_t* sc32 = cv.take<int32_t>(scan_ws_elems(RS_BINS * nb));
  int64_t* sc64 = cv.take<int64_t>(scan_ws_elems(N + 1));
  if (!cv.ok()) return fail(RGBMP_EWORKSPACE, "rgbmp_csr_build: workspace carve");
  const bool v2 = build_variant() != 1;
  if (!v2 || nnz == 0) {
    // rowptr = exclusive scan of the key histogram (int64)
    RGBMP_CUDA(cudaMemsetAsync(rowptr, 0, (size_t)(N + 1) * sizeof(int64_t), st));
    if (nnz > 0) {
      deg_hist_kernel<<<kSMs * 8, 256, 0, st>>>(key, nnz, (unsigned long long*)rowptr);
      RGBMP_LAUNCH_CHECK("deg_hist_kernel");
    }
    RGBMP_CUDA(exclusive_scan<int64_t>(rowptr, N + 1, sc64, st));
  }
  if (nnz == 0) return 0;

  int bits = 1;
  while (bits < 31 && (1ll << bits) < N) ++bits;
  const int32_t* sorted = nullptr;
  if (v2) {   // the column ids ride through the sort; rowptr straight from the sorted keys
    const int rc_sort = build_variant() == 2
        ? sort_pairs_i32(key, nnz, bits, kA, kB, vA, eid, bh, sc32, nb, st, &sorted, other, pA, pB, col)
        : sort_csr_packed(key, other, nnz, bits, kA, kB, (int2*)vA, (int2*)pB, eid, col, bh, sc32, nb, st, &sorted);
    if (rc_sort) return rc_sort;
    rowptr_lower_bound_kernel<<<(unsigned)ceil_div(N + 1, 256), 256, 0, st>>>(sorted, nnz, N, rowptr);
    RGBMP_LAUNCH_CHECK("rowptr_lower_bound_kernel");
    return 0;
  }
  const int rc_sort = sort_pairs_i32(key, nnz, bits, kA, kB, vA, eid, bh, sc32, nb, st, &sorted);
  if (rc_sort) return rc_sort;
  gather_i32_kernel<<<kSMs * 8, 256, 0, st>>>(other, eid, nnz, col);
  RGBMP_LAUNCH_CHECK("gather_i32_kernel");
  return 0;
}

int rgbmp_degree_norm(const int64_t* rowptr, int64_t N, int mode, float* out, int device, void* stream) {
  if (!rowptr || !out || N < 0 || mode < 0 || mode > 2) return fail(RGBMP_EINVAL, "rgbmp_degree_norm: bad argument");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RGBMP_EINVAL, "rgbmp_degree_norm: bad device");
  if (N == 0) return 0;
  degree_norm_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, (cudaStream_t)stream>>>(rowptr, N, mode, out);
  RGBMP_LAUNCH_CHECK("degree_norm_kernel");
  return 0;
}

int rgbmp_gcn_edge_weight(const int64_t* rowptr, const int32_t* col, int64_t n_rows, const float* dinv_row,
                          const float* dinv_col, float* val, int device, void* stream) {
  if (!rowptr || !dinv_row || !dinv_col || n_rows < 0) return fail(RGBMP_EINVAL, "rgbmp_gcn_edge_weight: bad argument");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RGBMP_EINVAL, "rgbmp_gcn_edge_weight: bad device");
  if (n_rows == 0) return 0;
  gcn_edge_weight_kernel<<<(unsigned)ceil_div(n_rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      rowptr, col, n_rows, dinv_row, dinv_col, val);
  RGBMP_LAUNCH_CHECK("gcn_edge_weight_kernel");
  return 0;
}

int rgbmp_edge_permute(const float* in, const int32_t* eid, int64_t nnz, int H, int scatter, float* out, int device,
                       void* stream) {
  if (nnz < 0 || H <= 0 || (nnz > 0 && (!in || !eid || !out))) return fail(RGBMP_EINVAL, "rgbmp_edge_permute: bad argument");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RGBMP_EINVAL, "rgbmp_edge_permute: bad device");
  if (nnz == 0) return 0;
  edge_permute_kernel<<<kSMs * 8, 256, 0, (cudaStream_t)stream>>>(in, eid, nnz, H, scatter, out);
  RGBMP_LAUNCH_CHECK("edge_permute_kernel");
  return 0;
}

int rgbmp_longrow_count(const int64_t* rowptr, int64_t n_rows, int32_t chunk, int32_t long_chunk, int64_t* counts_dev,
                        int device, void* stream) {
  if (!rowptr || !counts_dev || n_rows < 0 || chunk <= 0 || long_chunk <= 0)
    return fail(RGBMP_EINVAL, "rgbmp_longrow_count: bad argument");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RGBMP_EINVAL, "rgbmp_longrow_count: bad device");
  cudaStream_t st = (cudaStream_t)stream;
  RGBMP_CUDA(cudaMemsetAsync(counts_dev, 0, 2 * sizeof(int64_t), st));
  if (n_rows == 0) return 0;
  longrow_count_kernel<<<(unsigned)ceil_div(n_rows, 256), 256, 0, st>>>(rowptr, n_rows, chunk, long_chunk,
                                                                      (unsigned long long*)counts_dev);
  RGBMP_LAUNCH_CHECK("longrow_count_kernel");
  return 0;
}

size_t rgbmp_longrow_fill_workspace_bytes(int64_t n_rows) {
  const size_t n = (size_t)(n_rows > 0 ? n_rows : 1);
  return 2 * align_up(n * sizeof(int32_t), 256) + align_up(scan_ws_elems((int64_t)n) * sizeof(int32_t), 256) + 4096;
}

int rgbmp_longrow_fill(const int64_t* rowptr, int64_t n_rows, int32_t chunk, int32_t long_chunk, int64_t n_long,
                       int64_t n_items, int32_t* long_rows, int32_t* long_item_ptr, int32_t* item_long,
                       int64_t* item_start, void* ws, size_t ws_bytes, int device, void* stream) {
  return rgbmp_longrow_fill_ordered(rowptr, n_rows, chunk, long_chunk, nullptr, n_long, n_items, long_rows, long_item_ptr,
                                    item_long, item_start, ws, ws_bytes, device, stream);
}

int rgbmp_longrow_fill_ordered(const int64_t* rowptr, int64_t n_rows, int32_t chunk, int32_t long_chunk,
                               const int32_t* order, int64_t n_long, int64_t n_items, int32_t* long_rows,
                               int32_t* long_item_ptr, int32_t* item_long, int64_t* item_start, void* ws, size_t ws_bytes,
                               int device, void* stream) {
  if (!rowptr || n_rows <= 0 || n_long <= 0 || n_items <= 0 || !long_rows || !long_item_ptr || !item_long ||
      !item_start || !ws)
    return fail(RGBMP_EINVAL, "rgbmp_longrow_fill: bad argument");
  if (ws_bytes < rgbmp_longrow_fill_workspace_bytes(n_rows)) return fail(RGBMP_EWORKSPACE, "rgbmp_longrow_fill: workspace");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RGBMP_EINVAL, "rgbmp_longrow_fill: bad device");
  cudaStream_t st = (cudaStream_t)stream;
  Carver cv(ws, ws_bytes);
  int32_t* slot = cv.take<int32_t>((size_t)n_rows);
  int32_t* items = cv.take<int32_t>((size_t)n_rows);
  int32_t* sc = cv.take<int32_t>(scan_ws_elems(n_rows));
  const unsigned nb = (unsigned)ceil_div(n_rows, 256);
  longrow_flags_kernel<<<nb, 256, 0, st>>>(rowptr, n_rows, chunk, long_chunk, order, slot, items);
  RGBMP_LAUNCH_CHECK("longrow_flags_kernel");
  RGBMP_CUDA(exclusive_scan<int32_t>(slot, n_rows, sc, st));
  RGBMP_CUDA(exclusive_scan<int32_t>(items, n_rows, sc, st));
  longrow_fill_kernel<<<nb, 256, 0, st>>>(rowptr, n_rows, chunk, long_chunk, order, slot, items, n_long, n_items, long_rows,
                                          long_item_ptr, item_long, item_start);
  RGBMP_LAUNCH_CHECK("longrow_fill_kernel");
  return 0;
}

size_t rgbmp_row_order_workspace_bytes(int64_t n_rows) {
  const size_t n = (size_t)(n_rows > 0 ? n_rows : 1);
  const size_t nb = (size_t)ceil_div((int64_t)n, RS_TILE);
  return 4 * align_up(n * sizeof(int32_t), 256) + align_up(RS_BINS * nb * sizeof(int32_t), 256) +
         align_up(scan_ws_elems((int64_t)(RS_BINS * nb)) * sizeof(int32_t), 256) + 4096;
}

int rgbmp_row_order(const int64_t* rowptr, int64_t n_rows, int64_t window, int32_t* order, void* ws, size_t ws_bytes,
                    int device, void* stream) {
  return rgbmp_row_order_grouped(rowptr, n_rows, window, nullptr, 0, order, ws, ws_bytes, device, stream);
}

int rgbmp_row_order_grouped(const int64_t* rowptr, int64_t n_rows, int64_t window, const int32_t* group, int32_t n_groups,
                            int32_t* order, void* ws, size_t ws_bytes, int device, void* stream) {
  if (!rowptr || !order || n_rows <= 0 || (!group && window <= 0) || !ws) return fail(RGBMP_EINVAL, "rgbmp_row_order: bad argument");
  if (group && (n_groups <= 0 || n_groups > 32768)) return fail(RGBMP_ERANGE, "rgbmp_row_order: n_groups must be in 1..32768");
  if (n_rows >= (1ll << 31) - 1) return fail(RGBMP_ERANGE, "rgbmp_row_order: n_rows exceeds int32");
  if (ws_bytes < rgbmp_row_order_workspace_bytes(n_rows)) return fail(RGBMP_EWORKSPACE, "rgbmp_row_order: workspace");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RGBMP_EINVAL, "rgbmp_row_order: bad device");
  cudaStream_t st = (cudaStream_t)stream;
  if (!group && window < ceil_div(n_rows, 32768)) window = ceil_div(n_rows, 32768);   // window id must fit 15 bits
  const int64_t nwin = group ? (int64_t)n_groups : ceil_div(n_rows, window);
  int wbits = 0;
  while ((1ll << wbits) < nwin) ++wbits;
  const int64_t nb = ceil_div(n_rows, RS_TILE);
  Carver cv(ws, ws_bytes);
  int32_t* keys = cv.take<int32_t>((size_t)n_rows);
  int32_t* kA = cv.take<int32_t>((size_t)n_rows);
  int32_t* kB = cv.take<int32_t>((size_t)n_rows);
  int32_t* vA = cv.take<int32_t>((size_t)n_rows);
  int32_t* bh = cv.take<int32_t>((size_t)RS_BINS * nb);
  int32_t* sc32 = cv.take<int32_t>(scan_ws_elems(RS_BINS * nb));
  if (!cv.ok()) return fail(RGBMP_EWORKSPACE, "rgbmp_row_order: workspace carve");
  row_key_kernel<<<(unsigned)ceil_div(n_rows, 256), 256, 0, st>>>(rowptr, n_rows, group ? 1 : window, group, keys);
  RGBMP_LAUNCH_CHECK("row_key_kernel");
  return sort_pairs_i32(keys, n_rows, 16 + wbits, kA, kB, vA, order, bh, sc32, nb, st);
}

static size_t coalesce_ws(int64_t n, int64_t N) {
  (void)N;
  const size_t nn = (size_t)(n > 0 ? n : 1);
  const size_t nb = (size_t)ceil_div((int64_t)nn, RS_TILE);
  size_t b = 10 * align_up(nn * sizeof(int32_t), 256);                             // r, c, r1, c1, kA, kB, vA, p1, p2, flag
  b += align_up(RS_BINS * nb * sizeof(int32_t), 256);
  b += align_up(scan_ws_elems((int64_t)(RS_BINS * nb)) * sizeof(int32_t), 256);
  b += align_up(scan_ws_elems((int64_t)nn) * sizeof(int32_t), 256);
  return b + 4096;
}

size_t rgbmp_coalesce_workspace_bytes(int64_t E, int64_t N, int symmetrize) {
  return coalesce_ws(symmetrize ? 2 * E : E, N);
}

int rgbmp_coalesce(const int64_t* src, const int64_t* dst, int64_t E, int64_t N, int symmetrize, int64_t* out_src,
                   int64_t* out_dst, int64_t* count_dev, void* ws, size_t ws_bytes, int device, void* stream) {
  if (E < 0 || N < 0 || !count_dev || !ws || (E > 0 && (!src || !dst || !out_src || !out_dst)))
    return fail(RGBMP_EINVAL, "rgbmp_coalesce: null pointer or negative size");
  const int64_t n = symmetrize ? 2 * E : E;
  if (n >= (1ll << 31) - 1 || N >= (1ll << 31) - 1)
    return fail(RGBMP_ERANGE, "rgbmp_coalesce: %lld entries or N=%lld exceeds int32", (long long)n, (long long)N);
  if (ws_bytes < coalesce_ws(n, N)) return fail(RGBMP_EWORKSPACE, "rgbmp_coalesce: workspace");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RGBMP_EINVAL, "rgbmp_coalesce: bad device %d", device);
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    RGBMP_CUDA(cudaMemsetAsync(count_dev, 0, sizeof(int64_t), st));
    return 0;
  }
  const size_t nn = (size_t)n;
  const int64_t nb = ceil_div(n, RS_TILE);
  Carver cv(ws, ws_bytes);
  int32_t* r = cv.take<int32_t>(nn);
  int32_t* c = cv.take<int32_t>(nn);
  int32_t* r1 = cv.take<int32_t>(nn);
  int32_t* c1 = cv.take<int32_t>(nn);
  int32_t* kA = cv.take<int32_t>(nn);
  int32_t* kB = cv.take<int32_t>(nn);
  int32_t* vA = cv.take<int32_t>(nn);
  int32_t* p1 = cv.take<int32_t>(nn);
  int32_t* p2 = cv.take<int32_t>(nn);
  int32_t* flag = cv.take<int32_t>(nn);
  int32_t* bh = cv.take<int32_t>((size_t)RS_BINS * nb);
  int32_t* sc32 = cv.take<int32_t>(scan_ws_elems(RS_BINS * nb));
  int32_t* scf = cv.take<int32_t>(scan_ws_elems(n));
  if (!cv.ok()) return fail(RGBMP_EWORKSPACE, "rgbmp_coalesce: workspace carve");
  int32_t* err = bh;                       // one word, consumed before the histograms are written
  RGBMP_CUDA(cudaMemsetAsync(err, 0, sizeof(int32_t), st));
  int32_t* errkeep = scf + scan_ws_elems(n) - 1;   // last scratch word: survives until co_write
  co_expand_kernel<<<kSMs * 8, 256, 0, st>>>(src, dst, E, N, symmetrize, r, c, err);
  RGBMP_LAUNCH_CHECK("co_expand_kernel");
  RGBMP_CUDA(cudaMemcpyAsync(errkeep, err, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  int bits = 1;
  while (bits < 31 && (1ll << bits) < N) ++bits;
  // LSD over the pair: stable sort by col, then stable sort by row
  int rc = sort_pairs_i32(c, n, bits, kA, kB, vA, p1, bh, sc32, nb, st);
  if (rc) return rc;
  gather_i32_kernel<<<kSMs * 8, 256, 0, st>>>(r, p1, n, r1);
  gather_i32_kernel<<<kSMs * 8, 256, 0, st>>>(c, p1, n, c1);
  RGBMP_LAUNCH_CHECK("gather_i32_kernel");
  rc = sort_pairs_i32(r1, n, bits, kA, kB, vA, p2, bh, sc32, nb, st);
  if (rc) return rc;
  gather_i32_kernel<<<kSMs * 8, 256, 0, st>>>(r1, p2, n, r);      // r, c now hold the sorted pairs
  gather_i32_kernel<<<kSMs * 8, 256, 0, st>>>(c1, p2, n, c);
  RGBMP_LAUNCH_CHECK("gather_i32_kernel");
  co_flag_kernel<<<kSMs * 8, 256, 0, st>>>(r, c, n, flag);
  RGBMP_LAUNCH_CHECK("co_flag_kernel");
  RGBMP_CUDA(exclusive_scan<int32_t>(flag, n, scf, st));
  co_write_kernel<<<kSMs * 8, 256, 0, st>>>(r, c, n, flag, errkeep, out_src, out_dst, count_dev);
  RGBMP_LAUNCH_CHECK("co_write_kernel");
  return 0;
}

int rgbmp_col_freq(const int32_t* col, int64_t nnz, int64_t n_cols, int32_t* freq, int device, void* stream) {
  if (nnz < 0 || n_cols < 0 || (nnz > 0 && !col) || (n_cols > 0 && !freq)) return fail(RGBMP_EINVAL, "rgbmp_col_freq: bad argument");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RGBMP_EINVAL, "rgbmp_col_freq: bad device");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_cols > 0) RGBMP_CUDA(cudaMemsetAsync(freq, 0, (size_t)n_cols * sizeof(int32_t), st));
  if (nnz == 0) return 0;
  col_freq_kernel<<<kSMs * 16, 256, 0, st>>>(col, nnz, freq);
  RGBMP_LAUNCH_CHECK("col_freq_kernel");
  return 0;
}

int rgbmp_col_tag(const int32_t* col, int64_t nnz, const int32_t* freq, int32_t thresh, int32_t* out, int device,
                  void* stream) {
  if (nnz < 0 || (nnz > 0 && (!col || !freq || !out))) return fail(RGBMP_EINVAL, "rgbmp_col_tag: bad argument");
  DeviceGuard dg(device);
  if (!dg.ok) return fail(RGBMP_EINVAL, "rgbmp_col_tag: bad device");
  if (nnz == 0) return 0;
  col_tag_kernel<<<kSMs * 16, 256, 0, (cudaStream_t)stream>>>(col, nnz, freq, thresh, out);
  RGBMP_LAUNCH_CHECK("col_tag_kernel");
  return 0;
}

}  // extern "C"
