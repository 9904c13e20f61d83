#!/usr/bin/env python
"""Graph build only (edge edit + forward CSR + transpose CSR + long-row lists + row order) on one
named workload -- the target of the per-kernel ncu launch list of the integer kernels and of the
RGBMP_BUILD_VARIANT=1|2 A/B.   python tools/build_only.py [--workload products] [--reps 3]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="products")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import rgb_experiment_b200 as P
    import rgb_experiment_b200.synth as S
    dev = torch.device("cuda:0")
    sg = S.make_named(args.workload, device=dev, features=False)
    N, E = sg.num_nodes, sg.edge_index.size(1)
    torch.cuda.synchronize()
    ts, parts = [], None
    for _ in range(args.reps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        t0 = time.perf_counter()
        ev[0].record()
        g = P.Graph(sg.edge_index, N, P.LOOP_ADD_REMAINING)
        ev[1].record()
        _ = g.bwd
        ev[2].record()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
        parts = (ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]))
        nnz = g.nnz
        del g
    # algorithmic bytes of the build: edit 2 passes over int64 src,dst + int32 pair out;
    # per CSR: 3-4 radix passes of (hist 4 + read 8 + write 8) B, col gather 12 B, rowptr
    passes = (max(N - 1, 1).bit_length() + 7) // 8
    algo = E * 16 * 2 + nnz * 8 + 2 * (nnz * (passes * 20 + 12) + (N + 1) * 8)
    best = min(ts)
    print(json.dumps({"config": "graph_build(edit+CSR+transpose CSR)", "workload": args.workload, "N": N, "E": E, "nnz": nnz,
                      "variant": os.environ.get("RGBMP_BUILD_VARIANT", "3"), "ms": round(best, 3),
                      "edit_plus_fwd_ms": round(parts[0], 3), "transpose_ms": round(parts[1], 3),
                      "Medges_per_s": round(E / best / 1e3, 1), "algorithmic_GB": round(algo / 1e9, 2),
                      "algorithmic_GBps": round(algo / best / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    main()
