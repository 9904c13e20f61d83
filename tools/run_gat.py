#!/usr/bin/env python
"""One GATConv attention layer (H=8, C=8) forward + backward on the Reddit-shaped graph, a few
times -- the command profiled by ncu for the GAT kernels (profiles/r01_gat_v2_ncu.txt).
    python tools/run_gat.py [--iters 3] [--heads 8] [--channels 8]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--heads", type=int, default=8)
    ap.add_argument("--channels", type=int, default=8)
    ap.add_argument("--workload", default="reddit")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--long-chunk", type=int, default=0)
    args = ap.parse_args()
    import rgb_experiment_b200 as P
    import rgb_experiment_b200.synth as S
    dev = torch.device("cuda:0")
    sg = S.make_named(args.workload, device=dev, features=False)
    N, H, C = sg.num_nodes, args.heads, args.channels
    g = P.Graph(sg.edge_index, N, P.LOOP_REMOVE_THEN_ADD, chunk=args.chunk or None, long_chunk=args.long_chunk or None)
    _ = g.bwd
    gen = torch.Generator(device=dev).manual_seed(0)
    xp = torch.randn(N, H * C, device=dev, generator=gen, requires_grad=True)
    a_s = torch.randn(N, H, device=dev, generator=gen, requires_grad=True)
    a_d = torch.randn(N, H, device=dev, generator=gen, requires_grad=True)
    dout = torch.randn(N, H * C, device=dev, generator=gen)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for it in range(args.iters + 1):
        ev[0].record()
        out = P.ops.gat(xp, a_s, a_d, g, H, C, 0.2)
        ev[1].record()
        out.backward(dout)
        ev[2].record()
        torch.cuda.synchronize()
        if it > 0:
            tf += ev[0].elapsed_time(ev[1])
            tb += ev[1].elapsed_time(ev[2])
        xp.grad = a_s.grad = a_d.grad = None
    nnz = g.nnz
    fwd_bytes = nnz * (H * C * 4 + H * 4 + 4) + N * H * C * 4 + 2 * N * H * 4
    print(json.dumps({"workload": args.workload, "nnz": nnz, "H": H, "C": C, "fwd_ms": round(tf / args.iters, 3),
                      "bwd_ms": round(tb / args.iters, 3), "fwd_gteps": round(nnz / (tf / args.iters) / 1e6, 2),
                      "fwd_algorithmic_GBps": round(fwd_bytes / (tf / args.iters) / 1e6, 1),
                      "n_long_fwd": g.fwd.n_long, "n_items_fwd": g.fwd.n_items, "chunk": g.fwd.chunk, "long_chunk": g.fwd.long_chunk}))


if __name__ == "__main__":
    main()
