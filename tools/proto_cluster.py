#!/usr/bin/env python
"""Prototype (torch ops, GPU) of the LOCALITY-AWARE ROW SCHEDULE: does processing target rows community by
community raise the L2 hit rate of the gathered feature rows enough to pay?  

The SpMM gathers 192-byte rows of a 470 MB matrix; L2 (126 MB) hit rate 42 %, real DRAM traffic 15.6 GB per hop
against 1.95 GB compulsory.  Rows are scheduled by degree only, so the rows in flight at any time are unrelated
and the cold (low-degree) sources -- 16 gathers each per hop -- miss every time.  If the rows in flight belong to
one community, most of their sources do too and the community's slice of the matrix stays in L2.

Schedules compared on the products-shaped graph (APPNP hop, folded, F=47):
  base     global degree sort (round 1)
  lpa      seeded label propagation from the S highest-degree nodes (edges only, no generator knowledge),
           clusters chained greedily by normalised connectivity, rows sorted by (cluster rank, -degree)
  oracle   the generator's planted class (id % 47): the upper bound a perfect clustering would reach

    python tools/proto_cluster.py [--seeds 1024] [--iters 4] [--variants base,lpa,oracle]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def time_ms(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def lpa_labels(rowptr, col, n, seeds, iters, taus=(0.3, 0.15, 0.05, 0.0, 0.0, 0.0)):
    """Seeded label propagation, leaves first: the S highest-degree nodes keep their own label; in round t an
    UNLABELLED node takes the most frequent label among its labelled in-neighbours (ties: smallest label) once at
    least taus[t] of its neighbours carry one, and keeps it.  The threshold makes hubs wait for their own leaves
    instead of copying a bigger hub across one of the (few, but heavy) hub-hub edges."""
    dev = rowptr.device
    deg = rowptr[1:] - rowptr[:-1]
    order = torch.argsort(deg, descending=True, stable=True)
    S = seeds
    label = torch.full((n,), -1, dtype=torch.int64, device=dev)
    label[order[:S]] = torch.arange(S, device=dev)
    row = torch.repeat_interleave(torch.arange(n, device=dev), deg)
    colL = col.long()
    for it in range(iters):
        tau = taus[min(it, len(taus) - 1)]
        lj = label[colL]
        m = lj >= 0
        nlab = torch.zeros(n, dtype=torch.int64, device=dev).index_add_(0, row[m], torch.ones_like(row[m]))
        key = row[m] * S + lj[m]
        uk, cnt = torch.unique(key, return_counts=True)
        r, l = uk // S, uk % S
        score = cnt * S + (S - 1 - l)
        best = torch.full((n,), -1, dtype=torch.int64, device=dev)
        best.scatter_reduce_(0, r, score, reduce="amax", include_self=True)
        ok = (label < 0) & (best >= 0) & (nlab.double() >= tau * deg.double())
        new = torch.where(ok, S - 1 - (best % S), label)
        changed = int((new != label).sum())
        label = new
        print(json.dumps({"lpa_iter": it, "tau": tau, "changed": changed, "unlabelled": int((label < 0).sum())}), flush=True)
        del lj, m, key, uk, cnt, r, l, score, best, nlab
    label = torch.where(label < 0, torch.zeros_like(label), label)
    return label, row


def chain_clusters(label, row, colL, S):
    """Order the clusters so that strongly connected ones are adjacent: greedy chain on the normalised
    connectivity W[a,b] / (vol_a * vol_b) with an exponentially decayed affinity to the recently placed."""
    key = label[row] * S + label[colL]
    uk, cnt = torch.unique(key, return_counts=True)
    W = torch.zeros(S * S, dtype=torch.float64, device=label.device)
    W[uk] = cnt.double()
    W = W.view(S, S).cpu().numpy()
    intra = float(np.trace(W) / W.sum())
    vol = W.sum(1) + 1e-9
    Wn = W / vol[:, None] / vol[None, :]
    np.fill_diagonal(Wn, 0.0)
    placed = np.zeros(S, dtype=bool)
    cur = int(np.argmax(vol))
    rank = np.zeros(S, dtype=np.int64)
    aff = np.zeros(S)
    for p in range(S):
        rank[cur] = p
        placed[cur] = True
        aff = 0.5 * aff + Wn[cur]
        a = np.where(placed, -1.0, aff)
        cur = int(np.argmax(a))
    return torch.from_numpy(rank).to(label.device), intra


def apply_schedule(csr, group_rank_of_row):
    """Overwrite the row schedule and the long-row lists of a built CSR (prototype: torch ops)."""
    from rgb_experiment_b200._lib import ptr
    dev = csr.device
    n = csr.n_rows
    deg = (csr.rowptr[1:] - csr.rowptr[:-1])
    key = group_rank_of_row * 65536 + (65535 - deg.clamp(max=65535))
    csr.row_order = torch.argsort(key, stable=True).to(torch.int32)
    if csr.n_long > 0:
        lr = csr.long_rows.long()
        o = torch.argsort(group_rank_of_row[lr], stable=True)
        lr = lr[o]
        cnt = (deg[lr] + csr.long_chunk - 1) // csr.long_chunk
        ptr_ = torch.zeros(csr.n_long + 1, dtype=torch.int64, device=dev)
        ptr_[1:] = torch.cumsum(cnt, 0)
        item_long = torch.repeat_interleave(torch.arange(csr.n_long, device=dev), cnt)
        j = torch.arange(int(ptr_[-1]), device=dev) - ptr_[item_long]
        csr.long_rows = lr.to(torch.int32)
        csr.long_item_ptr = ptr_.to(torch.int32)
        csr.item_long = item_long.to(torch.int32)
        csr.item_start = (csr.rowptr[lr][item_long] + j * csr.long_chunk).contiguous()
    csr._tagged = {}
    csr.struct = csr._make_struct(csr.col, 0)
    csr.ref = C.byref(csr.struct)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="products")
    ap.add_argument("--F", type=int, default=47)
    ap.add_argument("--seeds", default="1024")
    ap.add_argument("--iters", type=int, default=6)
    ap.add_argument("--variants", default="base,lpa,oracle")
    ap.add_argument("--policies", default="default,off")
    ap.add_argument("--hops", type=int, default=10)
    ap.add_argument("--tunes", default="0", help="comma list of G:V:U launch shapes (0 = the library default)")
    args = ap.parse_args()
    import rgb_experiment_b200 as P
    import rgb_experiment_b200.synth as S
    from rgb_experiment_b200 import graph as G_
    dev = torch.device("cuda:0")
    sg = S.make_named(args.workload, device=dev, features=False)
    N = sg.num_nodes
    z0 = torch.randn(N, args.F, device=dev)
    z0p, _ = P.ops.as_rows(z0)

    def bench(g, tag, extra):
        for tn in args.tunes.split(","):
            P.ops.TUNE_OVERRIDE = 0
            if tn != "0":
                G__, V__, U__ = (int(v) for v in tn.split(":"))
                P.ops.TUNE_OVERRIDE = G__ | (V__ << 8) | (U__ << 16)
            for pol in args.policies.split(","):
                hot_saved = G_.HOT_L2_BYTES
                if pol == "off":
                    G_.HOT_L2_BYTES = 0
                g.fwd._tagged = {}
                ms = time_ms(lambda: P.ops._appnp_khop(g.fwd, g, z0p, args.hops, 0.1, False, True)) / args.hops
                G_.HOT_L2_BYTES = hot_saved
                print(json.dumps({"variant": tag, "policy": pol, "tune": tn, "ms_per_hop": round(ms, 4),
                                  "gteps": round(g.nnz / ms / 1e6, 2), **extra}), flush=True)
        P.ops.TUNE_OVERRIDE = 0

    variants = args.variants.split(",")
    g = P.Graph(sg.edge_index, N, P.LOOP_ADD_REMAINING)
    ref = P.ops._appnp_khop(g.fwd, g, z0p, 2, 0.1, False, True).clone()
    if "base" in variants:
        bench(g, "base", {})
    csr = g.fwd
    if "oracle" in variants:
        cls = torch.arange(N, device=dev) % 47
        apply_schedule(csr, cls)
        out = P.ops._appnp_khop(g.fwd, g, z0p, 2, 0.1, False, True)
        bench(g, "oracle", {"maxdiff_vs_base": float((out - ref).abs().max())})
    if "lpa" in variants:
        for S_ in [int(s) for s in args.seeds.split(",")]:
            label, row = lpa_labels(csr.rowptr, csr.col, N, S_, args.iters)
            colL = csr.col.long()
            rank, intra = chain_clusters(label, row, colL, S_)
            sizes = torch.bincount(label, minlength=S_)
            # how pure are the clusters w.r.t. the planted classes (diagnostic only, never used by the schedule)
            cls = torch.arange(N, device=dev) % 47
            pur = torch.zeros(S_ * 47, device=dev).index_add_(0, label * 47 + cls, torch.ones(N, device=dev)).view(S_, 47)
            purity = float(pur.max(1).values.sum() / N)
            del row, colL
            apply_schedule(csr, rank[label])
            out = P.ops._appnp_khop(g.fwd, g, z0p, 2, 0.1, False, True)
            bench(g, "lpa", {"seeds": S_, "iters": args.iters, "intra_cluster_edge_frac": round(intra, 4),
                             "largest_cluster": int(sizes.max()), "empty_clusters": int((sizes == 0).sum()),
                             "purity_vs_planted": round(purity, 4), "maxdiff_vs_base": float((out - ref).abs().max())})


if __name__ == "__main__":
    main()
