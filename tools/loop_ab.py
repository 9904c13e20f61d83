#!/usr/bin/env python
"""HISTORICAL: the tool that produced profiles/r02_loop_kernel_ab.txt.  The looping kernel it times was measured slower and
is not in the tree (kernel side kept as profiles/r02_loop_kernel.patch; the `row_sched` descriptor field, its builder call in
graph.py and `rgbmp_set_rows_per_group` in the header / _lib.py went with it), so this script does not run as is.
A/B of the looping short-row SpMM kernel (spmm_rows_loop_kernel) against the one-row-per-group kernel:
ms per hop of the K-hop families and per launch of the plain aggregations, for rows-per-group 0 (old kernel), 1, 2, ...
    python tools/loop_ab.py [--workloads products,arxiv,reddit] [--rpg 0,1,4,8,16,32]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def time_ms(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="products,arxiv,reddit")
    ap.add_argument("--rpg", default="0,1,2,4,8,16,32")
    ap.add_argument("--shapes", default="", help="extra G:V:U shapes to force, comma list (default: the heuristic only)")
    args = ap.parse_args()
    import rgb_experiment_b200 as P
    import rgb_experiment_b200.memo as memo
    import rgb_experiment_b200.synth as S
    memo.MIN_WORK = float("inf")
    L = P._lib.lib()
    dev = torch.device("cuda:0")
    plans = {"products": [("appnp", 47), ("sgc", 100), ("gcn_w", 47), ("sum", 64)],
             "arxiv": [("mean", 256), ("gcn_w", 40), ("appnp", 40)],
             "reddit": [("sum", 64), ("gcn_w", 41), ("mean", 602)]}
    shapes = [0] + [int(g) | (int(v) << 8) | (int(u) << 16) for g, v, u in
                    (s.split(":") for s in args.shapes.split(",") if s)]
    for wl in args.workloads.split(","):
        sg = S.make_named(wl, device=dev, features=False)
        N = sg.num_nodes
        g = P.Graph(sg.edge_index, N, P.LOOP_ADD_REMAINING)
        for kind, F in plans[wl]:
            x = torch.randn(N, F, device=dev)
            K = 10 if kind == "appnp" else (2 if kind == "sgc" else 1)
            for tune in shapes:
                P.ops.TUNE_OVERRIDE = tune
                if kind == "appnp":
                    fn = lambda: P.ops.appnp(x, g, K, 0.1)
                elif kind == "sgc":
                    fn = lambda: P.ops.gcn_power(x, g, K)
                elif kind == "gcn_w":
                    fn = lambda: P.ops.propagate(x, g, "gcn")
                else:
                    fn = lambda: P.ops.propagate(x, g, kind)
                res = {}
                for rpg in [int(r) for r in args.rpg.split(",")]:
                    L.rgbmp_set_rows_per_group(rpg)
                    res[rpg] = round(time_ms(fn) / K, 4)
                print(json.dumps({"workload": wl, "kind": kind, "F": F, "tune": hex(tune), "K": K, "ms_per_hop_by_rows_per_group": res}),
                      flush=True)
            del x
        del g, sg
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
