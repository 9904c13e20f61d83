#!/usr/bin/env python
"""Small-graph K-hop latency on the Cora-shaped graph: the library default (long rows split at 32 edges) against
larger long-row thresholds and a second launch shape.  GPU time per call (CUDA events, calls queued back to back)
and wall time per synchronised call.  (A cooperative all-hops-in-one-launch kernel was tried and removed: the hop
latency is the longest row's chain of dependent gathers, not the launches -- profiles/r02_small_graph_latency.txt.)
    python tools/khop_latency.py"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import rgb_experiment_b200 as P
    import rgb_experiment_b200.synth as S
    dev = torch.device("cuda:0")
    sg = S.make_named("cora", device=dev, features=False)
    N = sg.num_nodes
    g = P.Graph(sg.edge_index, N, P.LOOP_ADD_REMAINING)
    g0 = P.Graph(sg.edge_index, N, P.LOOP_NONE)
    deg = g.fwd.degree()
    print(json.dumps({"N": N, "nnz": g.nnz, "max_degree": int(deg.max()), "rows_over_64": int((deg > 64).sum())}), flush=True)
    gs = P.Graph(sg.edge_index, N, P.LOOP_ADD_REMAINING, chunk=64, long_chunk=512)       # long rows split at 64 edges
    gs2 = P.Graph(sg.edge_index, N, P.LOOP_ADD_REMAINING, chunk=1024, long_chunk=4096)   # the large-graph thresholds (round 1)
    other = 8 | (2 << 8) | (4 << 16)
    for name, gr, F, K in (("appnp K=10 F=7", g, 7, 10), ("appnp K=10 F=7 chunk 64/512", gs, 7, 10), ("appnp K=10 F=7 chunk 1024/4096", gs2, 7, 10),
                           ("C&S LP 50 hops F=7", g0, 7, 50), ("sgc K=2 F=1433", g, 1433, 2),
                           ("appnp K=10 F=64", g, 64, 10), ("appnp K=10 F=64 chunk 64/512", gs, 64, 10)):
        x = torch.randn(N, F, device=dev)
        xb, ldx = P.ops.as_rows(x)
        val = gr.gcn_val(False)
        ep = P.ops.make_epilogue(a=0.9, b=0.1, T=xb, ldt=ldx)
        for label, tune in (("default launch shape", 0), ("shape G8 V2 U4", other)):
            fn = lambda: P.ops.khop_raw(gr.fwd, xb, K, val=val, ep=ep, tune=tune)
            for _ in range(20):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(200):
                fn()
            e1.record()
            torch.cuda.synchronize()
            queued = e0.elapsed_time(e1) / 200 * 1e3
            t0 = time.perf_counter()
            for _ in range(200):
                fn()
                torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) / 200 * 1e6
            print(json.dumps({"case": name, "path": label, "us_per_call_queued": round(queued, 1), "us_per_call_synchronised": round(wall, 1),
                              "us_per_hop_queued": round(queued / K, 2)}), flush=True)


if __name__ == "__main__":
    main()
