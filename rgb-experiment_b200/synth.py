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
