"""Deterministic synthetic graphs of the BASELINE.json shapes.

Everything is derived from counter-based integer hashes evaluated with torch int64 ops, so the
same (shape, seed) gives the SAME graph on CPU and on any GPU -- the oracle (CPU) and the CUDA
path are compared on identical inputs, and partitions of a multi-GPU run agree on the graph.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

MASK63 = (1 << 63) - 1

SHAPES = {
    # name: (nodes, directed edges, features, classes)
    "cora": (2_708, 10_556, 1_433, 7),
    "arxiv": (169_343, 2_315_598, 128, 40),
    "reddit": (232_965, 114_615_892, 602, 41),
    "products": (2_449_029, 123_718_280, 100, 47),
}


def _mix(x: torch.Tensor) -> torch.Tensor:
    """63-bit integer hash (xorshift-multiply rounds; int64 multiply wraps, the mask keeps it >= 0)."""
    x = x & MASK63
    x = ((x ^ (x >> 30)) * 0x3F58476D1CE4E5B9) & MASK63
    x = ((x ^ (x >> 27)) * 0x14D049BB133111EB) & MASK63
    return x ^ (x >> 31)


def _uniform(idx: torch.Tensor, seed: int, stream: int) -> torch.Tensor:
    """U[0,1) float64 per index."""
    h = _mix(idx * 0x2545F4914F6CDD1D + (seed * 0x9E3779B97F4A7C15 + stream * 0x632BE59BD9B4E019 & MASK63))
    return (h >> 10).to(torch.float64) * (1.0 / (1 << 53))


def rmat_pairs(n_nodes: int, n_pairs: int, seed: int, device="cpu", a=0.57, b=0.19, c=0.19,
               uniform: bool = False, chunk: int = 1 << 24):
    """n_pairs (u, v) node pairs, u != v.  R-MAT (a,b,c,d) quadrant recursion on ceil(log2 N) bit
    levels, scaled to [0,N); or uniform (Erdos-Renyi-like) when uniform=True."""
    k = max(1, math.ceil(math.log2(max(n_nodes, 2))))
    us, vs = [], []
    for s in range(0, n_pairs, chunk):
        idx = torch.arange(s, min(n_pairs, s + chunk), dtype=torch.int64, device=device)
        if uniform:
            u = (_uniform(idx, seed, 1) * n_nodes).to(torch.int64)
            v = (_uniform(idx, seed, 2) * n_nodes).to(torch.int64)
        else:
            u = torch.zeros_like(idx)
            v = torch.zeros_like(idx)
            for lvl in range(k):
                r = _uniform(idx, seed, 10 + lvl)
                ub = (r >= a + b).to(torch.int64)                        # quadrants c,d -> row bit 1
                vb = (((r >= a) & (r < a + b)) | (r >= a + b + c)).to(torch.int64)   # b,d -> col bit 1
                u = (u << 1) | ub
                v = (v << 1) | vb
            # scramble the ids so hubs are not the low ids, then scale 2^k -> N
            u = (_mix(u + 0x51ED27) % (1 << k)) * n_nodes >> k
            v = (_mix(v + 0x51ED27) % (1 << k)) * n_nodes >> k
        v = torch.where(u == v, (v + 1) % n_nodes, v)
        us.append(u)
        vs.append(v)
    return torch.cat(us), torch.cat(vs)


@dataclass
class SynthGraph:
    edge_index: torch.Tensor     # int64 [2, E], symmetric, no self loops
    x: torch.Tensor              # float32 [N, F]
    y: torch.Tensor              # int64 [N]
    num_nodes: int
    num_classes: int


def make_graph(n_nodes: int, n_edges: int, n_feat: int, n_classes: int, *, seed: int = 20261018,
               feat_seed: int = 1, device="cpu", power_law: bool = True, homophily: float = 0.8,
               features: bool = True, feat_dtype=torch.float32) -> SynthGraph:
    """Symmetric graph with n_edges directed edges (n_edges/2 sampled pairs, both directions),
    planted classes y_i = hash(i) % C, homophily: with prob. `homophily` the second endpoint is moved
    to the nearest id carrying the first endpoint's class; x = onehot(y).M + N(0,1)."""
    pairs = n_edges // 2
    u, v = rmat_pairs(n_nodes, pairs, seed, device, uniform=not power_law)
    C = n_classes
    idx = torch.arange(pairs, dtype=torch.int64, device=device)
    same = _uniform(idx, seed, 99) < homophily
    v2 = v - (v % C) + (u % C)                       # class of node i is i % C
    v2 = torch.where(v2 >= n_nodes, v2 - C, v2)
    v2 = torch.where(v2 < 0, v, v2)
    v = torch.where(same & (v2 != u), v2, v)
    ei = torch.stack([torch.cat([u, v]), torch.cat([v, u])])
    y = torch.arange(n_nodes, dtype=torch.int64, device=device) % C
    x = None
    if features:
        g = torch.Generator(device="cpu").manual_seed(feat_seed)
        M = torch.randn(C, n_feat, generator=g).to(device)
        gd = torch.Generator(device=device).manual_seed(feat_seed)
        x = torch.randn(n_nodes, n_feat, generator=gd, device=device)
        x += M[y]
        x = x.to(feat_dtype)
    return SynthGraph(ei, x, y, n_nodes, C)


def make_named(name: str, **kw) -> SynthGraph:
    n, e, f, c = SHAPES[name]
    return make_graph(n, e, f, c, **kw)


# --------------------------------------------------------------------------------------------
# Row-generated directed graphs (papers100M-shaped, 111 M nodes / 3.2 B
# edges): the in-neighbours of target row i are a pure function of (seed, i, k), so any rank can
# generate exactly its own row block, as a CSR by construction (no edge list, no sort), and the
# graph is identical for every partition count.
# --------------------------------------------------------------------------------------------
PAPERS100M = (111_059_956, 3_231_371_744, 128)     # nodes, directed edges, features


def rowgen_degrees(n_nodes: int, n_edges: int, lo: int, hi: int, seed: int, device="cpu") -> torch.Tensor:
    """In-degree (without the self loop) of rows [lo, hi): a Pareto-like law d = mean/2 * u^-1/2
    clipped to [0, 64*mean], whose expectation is ~ the mean degree n_edges / n_nodes."""
    mean = n_edges / n_nodes
    idx = torch.arange(lo, hi, dtype=torch.int64, device=device)
    u = _uniform(idx, seed, 201).clamp_min(1e-12)
    d = (0.5 * mean) * u.pow(-0.5)
    return d.clamp_max(64.0 * mean).to(torch.int64)


def rowgen_block(n_nodes: int, n_edges: int, lo: int, hi: int, *, seed: int = 20261018, device="cpu",
                 skew: float = 2.0, locality: float = 0.0, self_loops: bool = True, rows_per_chunk: int = 1 << 21):
    """CSR of target rows [lo, hi): rowptr int64 [hi-lo+1], col int32 [nnz] (global source ids).
    Row i lists i itself first (the self loop GCN adds) and then deg_i sources: with probability
    `locality` a node within +-N/64 of i, otherwise scramble(floor(N * u^skew)) -- popular sources
    (a power law over columns) spread over the whole id range."""
    deg = rowgen_degrees(n_nodes, n_edges, lo, hi, seed, device)
    extra = 1 if self_loops else 0
    rowptr = torch.zeros(hi - lo + 1, dtype=torch.int64, device=device)
    torch.cumsum(deg + extra, 0, out=rowptr[1:])
    nnz = int(rowptr[-1].item())
    col = torch.empty(nnz, dtype=torch.int32, device=device)
    k_bits = max(1, math.ceil(math.log2(max(n_nodes, 2))))
    for r0 in range(lo, hi, rows_per_chunk):
        r1 = min(hi, r0 + rows_per_chunk)
        d = deg[r0 - lo:r1 - lo] + extra
        e0, e1 = int(rowptr[r0 - lo].item()), int(rowptr[r1 - lo].item())
        rows = torch.repeat_interleave(torch.arange(r0, r1, dtype=torch.int64, device=device), d)
        k = torch.arange(e0, e1, dtype=torch.int64, device=device) - rowptr[rows - lo]     # position inside the row
        key = rows * 0x100000001B3 + k
        u = _uniform(key, seed, 202)
        c = (u.pow(skew) * n_nodes).to(torch.int64).clamp_max(n_nodes - 1)
        c = (_mix(c + 0x51ED27) % (1 << k_bits)) * n_nodes >> k_bits            # scramble: hubs are not the low ids
        if locality > 0:
            near = _uniform(key, seed, 203) < locality
            span = max(1, n_nodes // 64)
            off = (_uniform(key, seed, 204) * (2 * span + 1)).to(torch.int64) - span
            c = torch.where(near, (rows + off) % n_nodes, c)
        if self_loops:
            c = torch.where(k == 0, rows, c)
        col[e0:e1] = c.to(torch.int32)
        del rows, k, key, u, c
    return rowptr, col
