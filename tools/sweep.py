#!/usr/bin/env python
"""On-GPU sweep of the SpMM launch shapes (G lanes per row, V vectors per lane, U edges in flight)
for the BASELINE shapes.  Prints one JSON line per (workload, F, dtype, G, V, U) with the CUDA-event
time per launch and the algorithmic GB/s, so that choose_shape() in csrc/spmm.cu can be set from
measurements rather than guesses.   python tools/sweep.py [--workloads products,arxiv] [--quick]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def time_ms(fn, iters):
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
    ap.add_argument("--workloads", default="products,arxiv")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--chunks", default="")
    ap.add_argument("--windows", default="16384", help="comma list of row-schedule windows (0 = natural order, -1 = global)")
    ap.add_argument("--us", default="2,4,8,18,20")
    ap.add_argument("--minimal", action="store_true", help="only the shapes with the fewest idle lanes")
    ap.add_argument("--nostream", action="store_true", help="turn the evict-first policy of the streamed arrays off")
    ap.add_argument("--policies", default="default", help="comma list: default | off (untagged) | hHcC with H,C in 0..2 "
                    "(L2 priority of hot / cold gathered rows: 0 normal, 1 evict-first, 2 evict-last)")
    ap.add_argument("--hot-mb", default="64", help="comma list of L2 budgets (MB) for the hot rows")
    ap.add_argument("--persist-mb", type=int, default=-1, help="set the persisting-L2 set-aside (MB) first")
    ap.add_argument("--F", default="", help="override the feature widths of the workload plan, e.g. 8,16,24,32 (fp32)")
    ap.add_argument("--shapes", default="", help="only these G:V:U shapes, e.g. 8:2:18,16:1:20")
    args = ap.parse_args()
    import rgb_experiment_b200 as P
    import rgb_experiment_b200.synth as S
    from rgb_experiment_b200 import graph as G_
    dev = torch.device("cuda:0")
    if args.persist_mb >= 0:
        import ctypes
        from rgb_experiment_b200._lib import lib, check
        got = ctypes.c_size_t(0)
        torch.zeros(1, device=dev)
        check(lib().rgbmp_l2_persist(0, args.persist_mb << 20, ctypes.addressof(got)), "l2_persist")
        print(json.dumps({"persisting_l2_bytes": got.value}), flush=True)
    plans = {"products": [(47, torch.float32), (100, torch.float32)],
             "arxiv": [(256, torch.float32), (40, torch.float32), (128, torch.bfloat16)],
             "reddit": [(64, torch.float32), (41, torch.float32)]}
    for wl in args.workloads.split(","):
        sg = S.make_named(wl, device=dev, features=False)
        N = sg.num_nodes
        chunk_opts = [(1024, 4096)]
        if args.chunks:
            chunk_opts = [tuple(int(v) for v in c.split(":")) for c in args.chunks.split(",")]
        combos = [(c, l, int(w), pol, int(mb)) for (c, l) in chunk_opts for w in args.windows.split(",")
                  for pol in args.policies.split(",") for mb in args.hot_mb.split(",")]
        g = None
        for chunk, lchunk, window, pol, hot_mb in combos:
            G_.HOT_L2_BYTES = hot_mb << 20
            hot = pol != "off"
            pbits = 0
            if pol not in ("default", "off"):
                pbits = (1 << 29) | (int(pol[3]) << 25) | (int(pol[1]) << 27)
            del g
            g = P.Graph(sg.edge_index, N, P.LOOP_ADD_REMAINING, chunk=chunk, long_chunk=lchunk,
                        window=(N if window < 0 else window))
            val = g.gcn_val(False)
            deg = g.fwd.degree()
            print(json.dumps({"workload": wl, "N": N, "nnz": g.nnz, "max_deg": int(deg.max()), "n_long": g.fwd.n_long,
                              "n_items": g.fwd.n_items, "chunk": chunk, "long_chunk": lchunk, "window": window,
                              "policy": pol, "hot_mb": hot_mb}), flush=True)
            plan = plans[wl] if not args.F else [(int(f), torch.float32) for f in args.F.split(",")]
            for F, dt in plan:
                x = torch.randn(N, F, device=dev).to(dt)
                xb, ld = P.ops.as_rows(x)
                esz = 2 if dt == torch.bfloat16 else 4
                nvec = (F * esz + 15) // 16
                out, _ = P.ops.alloc_rows(N, F, dt, dev)
                iters = 3 if args.quick else 5
                for weighted in (True, False):
                    res = []
                    for G in (1, 2, 4, 8, 16, 32):
                        for V in (1, 2):
                            if G * V < nvec and nvec <= 128:
                                continue
                            if G * V >= 2 * nvec + 4:
                                continue
                            if args.minimal and (V > 2 or G * V >= 2 * nvec):
                                continue
                            for U in [int(u) for u in args.us.split(",")]:
                                if args.shapes and f"{G}:{V}:{U}" not in args.shapes.split(","):
                                    continue
                                tune = G | (V << 8) | (U << 16) | ((1 << 24) if args.nostream else 0) | pbits
                                fn = lambda: P.ops.spmm_raw(g.fwd, xb, val if weighted else None, tune=tune, out=out, hot=hot)
                                ms = time_ms(fn, iters)
                                B = g.nnz * (F * esz + 4 + (4 if weighted else 0)) + N * F * esz + (N + 1) * 8
                                r = {"workload": wl, "F": F, "dtype": str(dt).split(".")[-1], "weighted": weighted,
                                     "G": G, "V": V, "U": U, "ms": round(ms, 4), "GBps": round(B / ms / 1e6, 1),
                                     "gteps": round(g.nnz / ms / 1e6, 3), "chunk": chunk, "window": window, "policy": pol, "hot_mb": hot_mb}
                                res.append(r)
                                print(json.dumps(r), flush=True)
                    if not res:
                        continue
                    best = min(res, key=lambda r: r["ms"])
                    default_ms = time_ms(lambda: P.ops.spmm_raw(g.fwd, xb, val if weighted else None, out=out, hot=hot,
                                                                tune=pbits), iters)
                    print(json.dumps({"BEST": best, "default_ms": round(default_ms, 4)}), flush=True)
        del g
        del sg
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
