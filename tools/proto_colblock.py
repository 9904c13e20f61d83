#!/usr/bin/env python
"""Prototype measurement: column-blocked SpMM (one launch per block of source nodes whose feature
slice fits L2, partial sums accumulated through the epilogue's acc_in) vs the single-pass kernel.
    python tools/proto_colblock.py [--workload products] [--F 47,100] [--blocks 1,3,4,5,6,8,10]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def time_ms(fn, iters=5):
    for _ in range(2):
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
    ap.add_argument("--workload", default="products")
    ap.add_argument("--F", default="47,100")
    ap.add_argument("--blocks", default="3,4,5,6,8,10")
    args = ap.parse_args()
    import rgb_experiment_b200 as P
    import rgb_experiment_b200.synth as S
    from rgb_experiment_b200 import graph as G_
    from rgb_experiment_b200._lib import check, lib, ptr, stream_of
    dev = torch.device("cuda:0")
    sg = S.make_named(args.workload, device=dev, features=False)
    N = sg.num_nodes
    g = P.Graph(sg.edge_index, N, P.LOOP_ADD_REMAINING)
    del sg
    dinv = g.dinv()
    val = g.gcn_val(False)
    for F in [int(f) for f in args.F.split(",")]:
        x = torch.randn(N, F, device=dev)
        xb, ld = P.ops.as_rows(x)
        z0, _ = P.ops.as_rows(torch.randn(N, F, device=dev))
        ref, _ = P.ops.alloc_rows(N, F, torch.float32, dev)
        ep_full = lambda **kw: P.ops.make_epilogue(a=0.9, b=0.1, T=z0, ldt=ld, **kw)
        base = lambda: P.ops.spmm_raw(g.fwd, xb, val, ep=ep_full(), out=ref)
        ms0 = time_ms(base)
        print(json.dumps({"F": F, "blocks": 1, "ms": round(ms0, 4), "gteps": round(g.nnz / ms0 / 1e6, 2)}), flush=True)
        for B in [int(b) for b in args.blocks.split(",")]:
            Wc = (N + B - 1) // B
            blocks = []
            for b in range(B):
                m = (g.e_src >= b * Wc) & (g.e_src < (b + 1) * Wc)
                csr = G_.CSR(g.e_dst[m].contiguous(), g.e_src[m].contiguous(), N, N)
                v = torch.empty(max(csr.nnz, 1), dtype=torch.float32, device=dev)
                check(lib().rgbmp_gcn_edge_weight(ptr(csr.rowptr), ptr(csr.col), N, ptr(dinv), ptr(dinv), ptr(v),
                                                  dev.index, stream_of(dev)), "w")
                blocks.append((csr, v[:csr.nnz]))
            y, _ = P.ops.alloc_rows(N, F, torch.float32, dev)
            eps = []
            for b in range(B):
                kw = {}
                if b > 0:
                    kw.update(acc_in=y, ld_acc=y.stride(0))
                if b == B - 1:
                    eps.append(ep_full(**kw))
                else:
                    eps.append(P.ops.make_epilogue(skip_empty=(b > 0), **kw))

            def run():
                for b, (csr, v) in enumerate(blocks):
                    P.ops.spmm_raw(csr, xb, v, ep=eps[b], out=y, hot=False)

            ms = time_ms(run)
            err = float((y[:, :F] - ref[:, :F]).abs().max() / ref[:, :F].abs().max())
            per = []
            for b, (csr, v) in enumerate(blocks):
                per.append(round(time_ms(lambda: P.ops.spmm_raw(csr, xb, v, ep=eps[b], out=y, hot=False), 3), 3))
            print(json.dumps({"F": F, "blocks": B, "ms": round(ms, 4), "gteps": round(g.nnz / ms / 1e6, 2),
                              "speedup": round(ms0 / ms, 3), "relerr_vs_single_pass": err, "per_block_ms": per,
                              "n_items": [c.n_items for c, _ in blocks]}), flush=True)
            del blocks, eps


if __name__ == "__main__":
    main()
