#!/usr/bin/env python
"""Times the other BASELINE.json configurations on one B200 (parity-test cases of bench.py's
contract, reported in DESIGN.md / profiles): graph build, C2 GraphSAGE epoch, C3 GAT layers,
C4 SGC hops and C&S label propagation.   python tools/bench_configs.py [--only c2,c3,c4]"""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PEAK = 6536.4


def ev_ms(fn, iters=5, warm=2):
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


def out(**kw):
    print(json.dumps(kw), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="build,c1,c2,c3,att,c4")
    args = ap.parse_args()
    only = set(args.only.split(","))
    import rgb_experiment_b200 as P
    import rgb_experiment_b200.synth as S
    import importlib
    L = importlib.import_module("rgb_experiment_b200.shim.nn")
    dev = torch.device("cuda:0")

    if "build" in only:
        for wl in ("arxiv", "products", "reddit"):
            sg = S.make_named(wl, device=dev, features=False)
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                g = P.Graph(sg.edge_index, sg.num_nodes, P.LOOP_ADD_REMAINING)
                _ = g.bwd
                torch.cuda.synchronize()
                ts.append((time.perf_counter() - t0) * 1e3)
                del g
            out(config="graph_build(edit+CSR+transpose CSR)", workload=wl, E=sg.edge_index.size(1), ms=min(ts),
                Medges_per_s=sg.edge_index.size(1) / min(ts) / 1e3)
            del sg

    if "c1" in only:
        # BASELINE configs[0]: GCN 2-layer hidden 64 on the Cora-shaped graph.  4 MB per hop: latency-bound, so the
        # numbers that matter are microseconds per call and launches per epoch, not a roofline fraction.
        sg = S.make_named("cora", device=dev)
        N = sg.num_nodes

        class GCN(nn.Module):        # models/gcn.py:18-31: GCNConv -> BatchNorm -> GCNConv, no ReLU / dropout between
            def __init__(self):
                super().__init__()
                self.c1, self.c2 = L.GCNConv(sg.x.size(1), 64), L.GCNConv(64, sg.num_classes)
                self.bn = nn.BatchNorm1d(64)

            def forward(self, x, ei):
                return F.log_softmax(self.c2(self.bn(self.c1(x, ei)), ei), 1)

        m = GCN().to(dev)
        opt = torch.optim.Adam(m.parameters(), lr=0.01)
        tr = torch.arange(N, device=dev) % 10 < 6

        def epoch():
            m.train()
            opt.zero_grad()
            F.nll_loss(m(sg.x, sg.edge_index)[tr], sg.y[tr]).backward()
            opt.step()
            m.eval()
            with torch.no_grad():
                m(sg.x, sg.edge_index)
                m(sg.x, sg.edge_index)

        ms = ev_ms(epoch, 50, 5)
        g = P.get_graph(sg.edge_index, N, P.LOOP_ADD_REMAINING)
        x64 = torch.randn(N, 64, device=dev)
        hop = ev_ms(lambda: P.ops.propagate(x64, g, "gcn"), 200, 20)
        khop = ev_ms(lambda: P.ops.appnp(x64[:, :7].contiguous(), g, 10, 0.1), 100, 10)
        out(config="C1 GCN 2-layer hidden 64 Cora-shaped", nnz=g.nnz, epoch_ms=ms, hop_F64_us=hop * 1e3,
            appnp_K10_F7_us=khop * 1e3, aggregation_launches_per_epoch=2 + 2 + 2 * 2,
            note="latency-bound (4 MB per hop): 2 train-forward + 2 backward + 2x2 eval SpMM launches per epoch, graph built once")
        del sg, g, m

    if "c2" in only:
        sg = S.make_named("arxiv", device=dev)
        N = sg.num_nodes

        class SAGE(nn.Module):       # shaped like models/graphsage.py (3 layers, hidden 256, BN between)
            def __init__(self):
                super().__init__()
                dims = [128, 256, 256, 40]
                self.ll = nn.ModuleList(nn.Linear(dims[i], dims[i + 1]) for i in range(3))
                self.lr = nn.ModuleList(nn.Linear(dims[i], dims[i + 1]) for i in range(3))
                self.bn = nn.ModuleList(nn.BatchNorm1d(256) for _ in range(2))

            def forward(self, x, g):
                for i in range(3):
                    x = P.ops.propagate(self.ll[i](x), g, "mean") + self.lr[i](x)
                    if i < 2:
                        x = self.bn[i](x)
                return F.log_softmax(x, 1)

        g = P.Graph(sg.edge_index, N, P.LOOP_REMOVE_THEN_ADD)
        m = SAGE().to(dev)
        opt = torch.optim.Adam(m.parameters(), lr=0.01)
        tr = torch.arange(N, device=dev) % 10 < 6

        def epoch():
            m.train()
            opt.zero_grad()
            F.nll_loss(m(sg.x, g)[tr], sg.y[tr]).backward()
            opt.step()
            m.eval()
            with torch.no_grad():
                m(sg.x, g)
                m(sg.x, g)

        ms = ev_ms(epoch, 10, 3)
        x256 = torch.randn(N, 256, device=dev)
        hop = ev_ms(lambda: P.ops.propagate(x256, g, "mean"), 20, 3)
        B = g.nnz * (256 * 4 + 4) + N * 256 * 4 + (N + 1) * 8
        out(config="C2 GraphSAGE 3x256 arxiv-shaped", epoch_ms=ms, hop_F256_ms=hop, hop_gteps=g.nnz / hop / 1e6,
            hop_GBps=B / hop / 1e6, frac_of_measured_hbm=B / hop / 1e6 / PEAK, note="features fit L2 (173 MB vs 126 MB: partly)")
        del sg, g, m

    if "c3" in only:
        sg = S.make_named("reddit", device=dev, features=False)
        N = sg.num_nodes
        g = P.Graph(sg.edge_index, N, P.LOOP_REMOVE_THEN_ADD)
        for H, C in ((8, 8), (1, 41)):
            xp = torch.randn(N, H * C, device=dev, requires_grad=True)
            a_s = torch.randn(N, H, device=dev, requires_grad=True)
            a_d = torch.randn(N, H, device=dev, requires_grad=True)
            fwd = ev_ms(lambda: P.ops.gat(xp.detach(), a_s.detach(), a_d.detach(), g, H, C, 0.2), 5, 2)
            do = torch.randn(N, H * C, device=dev)

            def fb():
                o = P.ops.gat(xp, a_s, a_d, g, H, C, 0.2)
                o.backward(do)
                xp.grad = a_s.grad = a_d.grad = None

            _ = g.bwd
            fbm = ev_ms(fb, 3, 1)
            B = g.nnz * (H * C * 4 + H * 4 + 4) + N * H * C * 4 + 2 * N * H * 4
            out(config=f"C3 GAT layer H={H} C={C} reddit-shaped", nnz=g.nnz, fwd_ms=fwd, fwd_gteps=g.nnz / fwd / 1e6,
                fwd_GBps=B / fwd / 1e6, frac_of_measured_hbm=B / fwd / 1e6 / PEAK, fwd_bwd_ms=fbm)
        # the whole 2-layer model of models/gat.py:18-31 (GATConv 602 -> 8x8, BatchNorm, GATConv 64 -> 41 with one head) for one
        # reference epoch (1 train fwd+bwd+Adam, 2 eval fwd) -- the configuration that runs the reference out of memory
        del sg
        sgx = S.make_named("reddit", device=dev)

        class GAT(nn.Module):
            def __init__(self):
                super().__init__()
                self.c1 = L.GATConv(sgx.x.size(1), 8, 8)
                self.bn = nn.BatchNorm1d(64)
                self.c2 = L.GATConv(64, sgx.num_classes, 1, concat=False)

            def forward(self, x, ei):
                return F.log_softmax(self.c2(self.bn(self.c1(x, ei)), ei), 1)

        m = GAT().to(dev)
        opt = torch.optim.Adam(m.parameters(), lr=0.01)
        tr = torch.arange(N, device=dev) % 10 < 6

        def epoch():
            m.train()
            opt.zero_grad()
            F.nll_loss(m(sgx.x, sgx.edge_index)[tr], sgx.y[tr]).backward()
            opt.step()
            m.eval()
            with torch.no_grad():
                m(sgx.x, sgx.edge_index)
                m(sgx.x, sgx.edge_index)

        torch.cuda.reset_peak_memory_stats()
        ms = ev_ms(epoch, 5, 2)
        out(config="C3 GAT 2-layer (8x8 heads, then 1x41) reddit-shaped, full-batch epoch", epoch_ms=ms,
            peak_mem_GiB=round(torch.cuda.max_memory_allocated() / 2 ** 30, 2),
            note="1 train fwd+bwd+Adam + 2 eval fwd; PyG materialises [nnz,8,8] = 29.4 GB per layer-forward here (SURVEY 8a a4)")
        del sgx, g, m

    if "att" in only:
        # the two layers that were an unfused chain in round 1: SuperGATConv-MX 8x8 on the arxiv-shaped graph (the
        # reference's OOM cell, 最终结果.csv:69) and on the Reddit-shaped one, FAConv F=64 on arxiv / products shapes
        for wl, mode in (("arxiv", P.LOOP_REMOVE_THEN_ADD), ("reddit", P.LOOP_REMOVE_THEN_ADD)):
            sg = S.make_named(wl, device=dev, features=False)
            N = sg.num_nodes
            g = P.Graph(sg.edge_index, N, mode)
            _ = g.bwd
            H, C = 8, 8
            xp = (torch.randn(N, H * C, device=dev) * 0.3).requires_grad_(True)
            a_l = torch.randn(N, H, device=dev, requires_grad=True)
            a_r = torch.randn(N, H, device=dev, requires_grad=True)
            do = torch.randn(N, H * C, device=dev)
            P.memo.set_budget_mb(0)
            torch.cuda.reset_peak_memory_stats()
            base = torch.cuda.memory_allocated()
            with torch.no_grad():
                fwd = ev_ms(lambda: P.ops.supergat_mx(xp.detach(), a_l.detach(), a_r.detach(), g, H, C, 0.2), 5, 2)
            peak_eval = torch.cuda.max_memory_allocated() - base

            def fb():
                o = P.ops.supergat_mx(xp, a_l, a_r, g, H, C, 0.2)
                o.backward(do)
                xp.grad = a_l.grad = a_r.grad = None

            fbm = ev_ms(fb, 3, 1)
            # algorithmic bytes (gather model): forward gathers X[j] (H*C*4) + a_l[j] (H*4) + col per edge; the MX backward
            # gathers dout_i + x_i + stats per edge on the transpose and x_j + a_l on the forward CSR
            Bf = g.nnz * (H * C * 4 + H * 4 + 4) + 2 * N * H * C * 4 + 3 * N * H * 4
            out(config=f"SuperGATConv-MX H={H} C={C} fused, {wl}-shaped", nnz=g.nnz, fwd_ms=fwd, fwd_gteps=g.nnz / fwd / 1e6,
                fwd_GBps=Bf / fwd / 1e6, frac_of_measured_hbm=Bf / fwd / 1e6 / PEAK, fwd_bwd_ms=fbm,
                eval_forward_extra_bytes=peak_eval, one_nnzH_tensor_bytes=g.nnz * H * 4)
            P.memo.set_budget_mb(4096)
            del sg, g, xp, a_l, a_r, do
        for wl in ("arxiv", "products"):
            sg = S.make_named(wl, device=dev, features=False)
            N = sg.num_nodes
            g = P.Graph(sg.edge_index, N, P.LOOP_ADD_REMAINING)
            _ = g.bwd
            Fh = 64
            x = torch.randn(N, Fh, device=dev, requires_grad=True)
            a_l = torch.randn(N, 1, device=dev, requires_grad=True)
            a_r = torch.randn(N, 1, device=dev, requires_grad=True)
            do = torch.randn(N, Fh, device=dev)
            P.memo.set_budget_mb(0)
            with torch.no_grad():
                fwd = ev_ms(lambda: P.ops.faconv(x.detach(), a_l.detach(), a_r.detach(), g), 5, 2)

            def fb2():
                o = P.ops.faconv(x, a_l, a_r, g)
                o.backward(do)
                x.grad = a_l.grad = a_r.grad = None

            fbm = ev_ms(fb2, 3, 1)
            Bf = g.nnz * (Fh * 4 + 4 + 4 + 4) + 2 * N * Fh * 4 + (N + 1) * 8
            spmm = ev_ms(lambda: P.ops.propagate(x.detach(), g, "gcn"), 5, 2)
            out(config=f"FAConv F={Fh} fused, {wl}-shaped", nnz=g.nnz, fwd_ms=fwd, fwd_gteps=g.nnz / fwd / 1e6,
                fwd_GBps=Bf / fwd / 1e6, frac_of_measured_hbm=Bf / fwd / 1e6 / PEAK, fwd_bwd_ms=fbm,
                plain_gcn_spmm_same_shape_ms=spmm)
            P.memo.set_budget_mb(4096)
            del sg, g, x, a_l, a_r, do

    if "c4" in only:
        sg = S.make_named("products", device=dev, features=False)
        N = sg.num_nodes
        g = P.Graph(sg.edge_index, N, P.LOOP_ADD_REMAINING)
        x = torch.randn(N, 100, device=dev)
        ms = ev_ms(lambda: P.ops.gcn_power(x, g, 2), 5, 2)
        B = g.nnz * (100 * 4 + 8) + N * 100 * 4 + (N + 1) * 8
        out(config="C4 SGC K=2 F=100 products-shaped", ms=ms, hop_ms=ms / 2, gteps=g.nnz * 2 / ms / 1e6,
            GBps=B * 2 / ms / 1e6, frac_of_measured_hbm=B * 2 / ms / 1e6 / PEAK)
        g0 = P.Graph(sg.edge_index, N, P.LOOP_NONE)
        y = torch.softmax(torch.randn(N, 47, device=dev), -1)
        ms = ev_ms(lambda: P.ops.label_propagation(g0, y, 50, 0.8), 3, 1)
        B = g0.nnz * (47 * 4 + 8) + 2 * N * 47 * 4 + (N + 1) * 8
        out(config="C4 C&S label propagation 50 hops F=47 products-shaped", ms=ms, hop_ms=ms / 50,
            gteps=g0.nnz * 50 / ms / 1e6, GBps=B * 50 / ms / 1e6, frac_of_measured_hbm=B * 50 / ms / 1e6 / PEAK)
        z = torch.randn(N, 47, device=dev, requires_grad=True)
        _ = g.bwd

        def fb():
            o = P.ops.appnp(z, g, 10, 0.1)
            o.backward(torch.ones_like(o))
            z.grad = None

        ms = ev_ms(fb, 3, 1)
        out(config="C4 APPNP K=10 fwd+bwd F=47 products-shaped", ms=ms, gteps=g.nnz * 20 / ms / 1e6)


if __name__ == "__main__":
    main()
