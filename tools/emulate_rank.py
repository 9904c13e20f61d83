#!/usr/bin/env python
"""One rank's hop of the partitioned products-shaped APPNP, alone on one GPU (no NVLink, no NCCL): how long does the
row block's SpMM itself take?  Separates the kernel's efficiency on (rows/Pr) x (F/Pf) blocks from exchange effects.
    python tools/emulate_rank.py [--Pr 4] [--Pf 2]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--Pr", default="1,2,4,8")
    ap.add_argument("--Pf", default="1,2")
    ap.add_argument("--F", type=int, default=47)
    ap.add_argument("--chunks", default="1024:4096", help="comma list of chunk:long_chunk")
    ap.add_argument("--by-community", action="store_true",
                    help="rank 0 owns the first N/Pr nodes in (locality group, id) order instead of the ids [0, N/Pr)")
    args = ap.parse_args()
    import rgb_experiment_b200 as P
    import rgb_experiment_b200.partition as PT
    import rgb_experiment_b200.synth as S
    from rgb_experiment_b200.graph import CSR
    dev = torch.device("cuda:0")
    sg = S.make_named("products", device=dev, features=False)
    N = sg.num_nodes
    g = P.Graph(sg.edge_index, N, P.LOOP_ADD_REMAINING)
    dinv = g.dinv()
    for Pr in [int(v) for v in args.Pr.split(",")]:
        R = PT.rows_per_rank(N, Pr)
        for Pf in [int(v) for v in args.Pf.split(",")]:
            Fl = (args.F + Pf - 1) // Pf
            ld = P.ops.padded_width(Fl)
            res = []
            for rp in range(1):
                lo, hi = PT.row_range(N, rp, Pr)
                if args.by_community and g.groups is not None:
                    # position of every node in (group, id) order; the block owns positions [lo, hi); columns keep their ids
                    order = torch.argsort(g.groups[0].long() * N + torch.arange(N, device=dev), stable=True)
                    pos = torch.empty(N, dtype=torch.int64, device=dev)
                    pos[order] = torch.arange(N, device=dev)
                    pd = pos[g.e_dst.long()]
                    m = (pd >= lo) & (pd < hi)
                    key, other = (pd[m] - lo).to(torch.int32), g.e_src[m].to(torch.int32)
                    grp_local = g.groups[0][order[lo:hi]]
                else:
                    key, other = PT.local_edges(g.e_src, g.e_dst, lo, hi)
                    grp_local = g.groups[0][lo:hi] if g.groups is not None else None
                for use_groups, chunk, lchunk in [(ug, int(c.split(":")[0]), int(c.split(":")[1])) for c in args.chunks.split(",")
                                                  for ug in (True, False)]:
                    groups = None
                    if use_groups and g.groups is not None:
                        mine = torch.zeros(R, dtype=torch.int32, device=dev)
                        mine[: hi - lo] = grp_local
                        groups = (mine, g.groups[1])
                    csr = CSR(key, other, R, R * Pr, chunk=chunk, long_chunk=lchunk, groups=groups)
                    x = torch.randn(R * Pr, ld, device=dev)
                    out = torch.empty(R, ld, device=dev)
                    d = torch.zeros(R, device=dev)
                    d[: hi - lo] = dinv[lo:hi]
                    z0 = torch.randn(R, ld, device=dev)
                    ep = P.ops.make_epilogue(row_scale=d, a=0.9, b=0.1, T=z0, ldt=ld, out2_scale=d)
                    fn = lambda: P.ops.spmm_raw(csr, x[:, :Fl], None, ep=ep, out=out[:, :Fl])
                    for _ in range(3):
                        fn()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(10):
                        fn()
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / 10
                    res.append({"rp": rp, "groups": use_groups, "chunk": f"{chunk}:{lchunk}", "ms": round(ms, 4), "n_items": csr.n_items})
                    del csr, x, out
            print(json.dumps({"by_community": bool(args.by_community), "Pr": Pr, "Pf": Pf, "F_local": Fl, "rows": R, "runs": res,
                              "whole_job_gteps_if_all_ranks_like_this": round(g.nnz / max(r["ms"] for r in res if r["groups"]) / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    main()
