#!/usr/bin/env python
"""Small-graph K-hop latency on the Cora-shaped graph: the one-cluster shared-memory path (csrc/khop_cta.cu, the default for
graphs whose iterate fits a 16-CTA cluster) against the K-launch path (long rows split at 32 edges; larger thresholds for
comparison).  GPU time per call (CUDA events, calls queued back to back) and wall time per synchronised call, for the raw
C-ABI call and for the public op (ops.appnp / label_propagation: padding, pre-scale and autograd glue included).
    python tools/khop_latency.py"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


_big = None


def heat(ms=150):
    """Busy the whole GPU for ~ms so that the timed calls start at boost clocks."""
    global _big
    if _big is None:
        _big = torch.randn(8192, 8192, device="cuda:0", dtype=torch.bfloat16)
    t0 = time.perf_counter()
    while (time.perf_counter() - t0) * 1e3 < ms:
        _big @ _big
        torch.cuda.synchronize()


def sm_clock():
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        return pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
    except Exception:
        return None


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
    L = P._lib.lib()
    import rgb_experiment_b200.memo as memo
    memo.MIN_WORK = float('inf')
    quick = "--quick" in sys.argv
    for name, gr, F, K in (("appnp K=1 F=7", g, 7, 1), ("appnp K=10 F=7", g, 7, 10), ("appnp K=10 F=7 chunk 64/512", gs, 7, 10),
                           ("appnp K=10 F=7 chunk 1024/4096", gs2, 7, 10), ("C&S LP 50 hops F=7", g0, 7, 50), ("sgc K=2 F=1433", g, 1433, 2),
                           ("appnp K=10 F=64", g, 64, 10)):
        if quick and ("chunk" in name or "sgc" in name or "K=1 " in name):
            continue
        x = torch.randn(N, F, device=dev)
        xb, ldx = P.ops.as_rows(x)
        val = gr.gcn_val(False)
        ep = P.ops.make_epilogue(a=0.9, b=0.1, T=xb, ldt=ldx)
        calls = {"rgbmp_khop (C ABI, per-edge weights)": lambda: P.ops.khop_raw(gr.fwd, xb, K, val=val, ep=ep)}
        if name.startswith("appnp"):
            calls["ops.appnp (public op, folded norm)"] = lambda: P.ops.appnp(x, gr, K, 0.1)
        if name.startswith("C&S"):
            y0 = torch.rand(N, F, device=dev)
            calls["ops.label_propagation (public op)"] = lambda: P.ops.label_propagation(gr, y0, K, 0.8)
        if quick:
            calls = dict(list(calls.items())[:1])
        for cname, fn in calls.items():
            for label, on in (("one cluster, iterate in distributed shared memory", 1), ("K launches", 0)):
                if quick and not on:
                    continue
                L.rgbmp_set_khop_cta(on)
                before = L.rgbmp_khop_cta_calls()
                for _ in range(20):
                    fn()
                heat()                                   # a one-SM kernel alone does not pull the clocks up
                torch.cuda.synchronize()
                took = L.rgbmp_khop_cta_calls() > before
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(200):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                queued = e0.elapsed_time(e1) / 200 * 1e3
                mhz = sm_clock()
                t0 = time.perf_counter()
                for _ in range(200):
                    fn()
                    torch.cuda.synchronize()
                wall = (time.perf_counter() - t0) / 200 * 1e6
                print(json.dumps({"case": name, "call": cname, "path": label, "one_cta_path_taken": took,
                                  "us_per_call_queued": round(queued, 1), "us_per_call_synchronised": round(wall, 1),
                                  "us_per_hop_queued": round(queued / K, 2), "sm_mhz_after": mhz}), flush=True)
        L.rgbmp_set_khop_cta(1)


if __name__ == "__main__":
    main()
