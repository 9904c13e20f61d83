#!/usr/bin/env python
"""Full-batch training epoch of the reference's APPNPStack on the products-shaped synthetic graph at
1/2/4/8 GPUs -- the second half of the BASELINE metric ("full-batch epoch ms at 1/2/4/8 GPU").

The model is the reference's (rgb_experiment/models/appnp_stack.py:19-31): Linear(in, hidden) ->
BatchNorm1d -> Linear(hidden, classes) -> APPNP(K, alpha) -> log_softmax; the epoch is the reference's
(rgb_experiment/itexperiments.py:417-473): 1 train forward + NLL loss on the train mask + backward +
Adam step, then 2 eval forwards (val, test) with their losses.  Restated here because the GPU box has
neither the reference nor torch_geometric.

Multi-GPU: rows are partitioned over the grid's row blocks (the dense layers run on the local rows,
BatchNorm statistics are synchronised over a row group, weight gradients are summed over all ranks),
APPNP is partition.DistAPPNP (fused-push K-hop forward, the same on the transposed row block
backward).  With --check the logits and the weight gradients of the first step are compared with the
single-GPU model on rank 0.

    python tools/bench_epoch.py                                            # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29571 tools/bench_epoch.py [--workload products] [--epochs 5] [--check]"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as Fn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class APPNPStack(nn.Module):
    """appnp_stack.py:19-31 with the propagation injected (single-GPU shim APPNP or DistAPPNP)."""

    def __init__(self, input_dim, hidden, output_dim, prop, bn):
        super().__init__()
        self.lin1 = nn.Linear(input_dim, hidden)
        self.lin2 = nn.Linear(hidden, output_dim)
        self.bn = bn
        self.prop = prop

    def forward(self, x):
        x = self.lin2(self.bn(self.lin1(x)))
        return Fn.log_softmax(self.prop(x), dim=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="products")
    ap.add_argument("--hidden", type=int, default=64)
    ap.add_argument("--K", type=int, default=10)
    ap.add_argument("--alpha", type=float, default=0.1)
    ap.add_argument("--epochs", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--feature-groups", type=int, default=0)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--relabel", type=int, default=1, help="N > 1: row blocks = runs of whole locality groups (1) or id ranges (0)")
    ap.add_argument("--as-called", action="store_true", help="1 GPU: also time the epoch through the reference's own classes")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    lr_ = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr_)
    dev = torch.device("cuda", lr_)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    res = run(args, rank, world, dev)
    if args.as_called and world == 1:
        res["as_called"] = run_as_called(args, dev)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


def default_args(**kw):
    """The argument set of main() as a namespace (bench.py calls run() with it)."""
    d = dict(workload="products", hidden=64, K=10, alpha=0.1, epochs=5, warmup=2, feature_groups=0, check=False,
             as_called=False, relabel=1)
    d.update(kw)
    return argparse.Namespace(**d)


def reference_root():
    snap = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(snap, "rgb_experiment")):
        return snap
    return "/root/reference" if os.path.isdir("/root/reference/rgb_experiment") else None


def run_as_called(args, dev, sg=None):
    """The epoch AS THE REFERENCE CALLS IT (SURVEY.md 8d "metrics sync included"), single GPU: the reference's own
    model class (models/appnp_stack.py:19-31, imported unmodified from baseline/_ref) over the product shim's
    APPNP, and the reference's own `test()` / `compare_pred_label()` (itexperiments.py:600-664: three host syncs and
    four sklearn metric calls per evaluation) driven by the statements of its epoch loop (:417-473, restated below
    line by line because the loop lives inside the monolithic experiment()).  Wall clock per epoch, GPU drained
    at the end of every epoch by the loop's own .item() calls."""
    import copy
    root = reference_root()
    if root is None:
        return {"unavailable": "no reference snapshot (baseline/_ref) on this box"}
    import rgb_experiment_b200 as P
    import rgb_experiment_b200.synth as S
    P.install_shim()
    sys.path.insert(0, root)
    try:
        import rgb_experiment.itexperiments as it
        from rgb_experiment.models import APPNPStack as RefAPPNPStack
    finally:
        sys.path.remove(root)
    if sg is None:
        sg = S.make_named(args.workload, device=dev)
    N, Fin, C = sg.num_nodes, sg.x.size(1), sg.num_classes
    y, features, edge_index = sg.y, sg.x, sg.edge_index
    sel = (S._mix(torch.arange(N, device=dev)) % 10)
    train_mask, val_mask, test_mask = sel < 6, (sel >= 6) & (sel < 8), sel >= 8
    torch.manual_seed(14530529)
    model = RefAPPNPStack(input_dim=Fin, output_dim=C, hidden_unit=args.hidden, dropout_rate=0.5, alpha=args.alpha, K=args.K)
    model.to(dev)
    optimizer = torch.optim.Adam(model.parameters(), lr=0.01, weight_decay=0)
    criterion = nn.NLLLoss()
    model_forward_param = {"x": features, "edge_index": edge_index}
    metric_s = [0.0]
    real_cmp = it.compare_pred_label

    def timed_cmp(*a, **k):                     # same function, its host time (device sync + sklearn) accumulated
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = real_cmp(*a, **k)
        metric_s[0] += time.perf_counter() - t0
        return r

    it.compare_pred_label = timed_cmp
    state = {"best": 0.0, "best_model": None}

    def epoch(i):
        model.train()                                                               # :419
        optimizer.zero_grad()
        model_out = model(**model_forward_param)                                    # :427
        out = model_out["out"]
        loss = criterion(out[train_mask], y[train_mask])
        it.compare_pred_label(out[train_mask].max(dim=1)[1], y[train_mask], True)   # :434
        loss.item()                                                                 # :437
        loss.backward()
        optimizer.step()
        val_dict = it.test(model, model_forward_param, y, val_mask, True)           # :464
        criterion(val_dict["test_op"][val_mask], y[val_mask]).item()
        test_dict = it.test(model, model_forward_param, y, test_mask, True)         # :470
        criterion(test_dict["test_op"][test_mask], y[test_mask]).item()
        if val_dict["ACC"] >= state["best"]:                                        # :491-494
            state["best"] = val_dict["ACC"]
            state["best_model"] = copy.deepcopy(model.state_dict())
        return val_dict["ACC"], test_dict["ACC"]

    try:
        for i in range(args.warmup):
            epoch(i)
        torch.cuda.synchronize()
        metric_s[0] = 0.0
        t0 = time.perf_counter()
        for i in range(args.epochs):
            v, t = epoch(i)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / args.epochs * 1e3
    finally:
        it.compare_pred_label = real_cmp
    m_ms = metric_s[0] / args.epochs * 1e3
    return {"config": f"reference APPNPStack class + reference test()/compare_pred_label, {args.workload}-shaped, as called",
            "epoch_wall_ms": round(wall, 2), "of_which_metrics_host_ms": round(m_ms, 2),
            "epoch_wall_ms_without_metric_calls": round(wall - m_ms, 2),
            "graph_builds": P.graph.stats["builds"], "val_acc": round(v, 4), "test_acc": round(t, 4)}


def run(args, rank, world, dev):
    """One measurement; needs an initialised NCCL process group when world > 1.  Returns the result dict
    (identical on every rank up to the accuracies, which rank 0 reports)."""
    import rgb_experiment_b200 as P
    import rgb_experiment_b200.partition as PT
    import rgb_experiment_b200.synth as S

    if args.workload == "medium":
        sg = S.make_graph(300_001, 6_000_000, 64, 16, device=dev)
    else:
        sg = S.make_named(args.workload, device=dev)
    N, Fin, C = sg.num_nodes, sg.x.size(1), sg.num_classes
    y = sg.y
    sel = (S._mix(torch.arange(N, device=dev)) % 10)                 # deterministic 60/20/20 split
    masks = {"train": sel < 6, "val": (sel >= 6) & (sel < 8), "test": sel >= 8}
    torch.manual_seed(14530529)                                       # reappear_seed (itexperiments.py:57)
    ref_lin1, ref_lin2 = nn.Linear(Fin, args.hidden), nn.Linear(args.hidden, C)

    if world == 1:
        g = P.Graph(sg.edge_index, N, P.LOOP_ADD_REMAINING)
        prop = lambda h: P.ops.appnp(h, g, args.K, args.alpha)        # the shim's default form (ops.FOLD_KHOP)
        bn = nn.BatchNorm1d(args.hidden)
        lo, hi, R = 0, N, N
        grid = None
    else:
        Pf = args.feature_groups if args.feature_groups > 0 else PT.auto_feature_groups(world, C)
        grid = PT.Grid(rank, world, Pf)
        a, b = grid.feature_slice(C)
        # community row blocks: the nodes are renamed by locality group, my rows are the nodes fwd.perm[lo:hi] of the dataset's
        # numbering (features, labels and masks are read through it once, below); the transposed block takes the same naming
        fwd = PT.LocalBlock(sg.edge_index, N, P.LOOP_ADD_REMAINING, grid.rp, grid.Pr, group=grid.row_group,
                            relabel=bool(getattr(args, "relabel", 1)), row_bytes=(b - a) * 4)
        bwd = PT.LocalBlock(sg.edge_index, N, P.LOOP_ADD_REMAINING, grid.rp, grid.Pr, group=grid.row_group, transpose_of=fwd,
                            row_bytes=(b - a) * 4)
        ld = PT.DistAPPNP.slice_ld(grid, C)
        pf = PT.PartitionedAPPNP(fwd, b - a, group=grid.row_group, ld=ld)
        pb = PT.PartitionedAPPNP(bwd, b - a, group=grid.row_group, ld=ld)
        prop = PT.DistAPPNP(grid, C, args.K, args.alpha, pf.run, pb.run, col_group=grid.col_group)
        bn = nn.SyncBatchNorm(args.hidden, process_group=grid.row_group) if grid.Pr > 1 else nn.BatchNorm1d(args.hidden)
        lo, hi, R = fwd.lo, fwd.hi, fwd.R
        dist_prop = prop
        prop = lambda h: dist_prop(Fn.pad(h, (0, 0, 0, R - (hi - lo))))[: hi - lo]
    model = APPNPStack(Fin, args.hidden, C, prop, bn).to(dev)
    model.lin1.load_state_dict(ref_lin1.state_dict())
    model.lin2.load_state_dict(ref_lin2.state_dict())
    opt = torch.optim.Adam(model.parameters(), lr=0.01)

    # the dense layers (and the BatchNorm statistics) see exactly my rows; only the propagation works on
    # ceil(N/Pr)-row blocks, so its input is zero-padded and its output trimmed
    rows = slice(lo, hi)
    if world > 1 and fwd.perm is not None:
        rows = fwd.perm[lo:hi]
    x_loc, y_loc = sg.x[rows].contiguous(), y[rows].contiguous()
    m_loc = {k: m[rows].contiguous() for k, m in masks.items()}
    n_glob = {k: int(m.sum()) for k, m in masks.items()}
    params = [p for p in model.parameters()]

    def loss_of(logp, key):
        """NLLLoss(mean) over the GLOBAL mask: local sum / global count (the ranks of a row block that hold
        other feature slices see the same rows, so only the rank with fp == 0 contributes)."""
        l = Fn.nll_loss(logp[m_loc[key]], y_loc[m_loc[key]], reduction="sum") / n_glob[key]
        return l

    def train_step():
        model.train()
        opt.zero_grad(set_to_none=True)
        logp = model(x_loc)
        loss = loss_of(logp, "train")
        loss.backward()
        if world > 1:
            # every rank of a row block back-propagates only its own feature slice (partial gradients); the
            # sum over ALL ranks is the full gradient
            flat = torch.cat([p.grad.reshape(-1) for p in params])
            dist.all_reduce(flat)
            off = 0
            for p in params:
                p.grad.copy_(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
        opt.step()
        return loss, logp

    @torch.no_grad()
    def eval_step(key):
        model.eval()
        logp = model(x_loc)
        loss = loss_of(logp, key)
        pred = logp.max(dim=1)[1]
        stat = torch.stack([(pred[m_loc[key]] == y_loc[m_loc[key]]).sum().float(), loss])
        if world > 1 and grid.Pr > 1:
            dist.all_reduce(stat, group=grid.row_group)
        return stat[0].item() / n_glob[key], stat[1].item()       # .item(): the reference syncs per eval too

    phase_ev = []                                                  # CUDA events at the phase boundaries of every epoch

    def mark():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def epoch():
        ev = [mark()]
        loss, _ = train_step()
        loss.item()                                                # train_losses.append(loss.item())
        ev.append(mark())
        v = eval_step("val")
        ev.append(mark())
        t = eval_step("test")
        ev.append(mark())
        phase_ev.append(ev)
        return v, t

    check = None
    if args.check:
        # first training step: logits + weight gradients vs the single-GPU model on every rank's own GPU
        g1 = P.Graph(sg.edge_index, N, P.LOOP_ADD_REMAINING)
        m1 = APPNPStack(Fin, args.hidden, C, lambda h: P.ops.appnp(h, g1, args.K, args.alpha),
                        nn.BatchNorm1d(args.hidden)).to(dev)
        m1.lin1.load_state_dict(ref_lin1.state_dict())
        m1.lin2.load_state_dict(ref_lin2.state_dict())
        m1.train()
        lp1 = m1(sg.x)
        l1 = Fn.nll_loss(lp1[masks["train"]], y[masks["train"]])
        l1.backward()
        model.train()
        opt.zero_grad(set_to_none=True)
        lp = model(x_loc)
        l = loss_of(lp, "train")
        l.backward()
        flat = torch.cat([p.grad.reshape(-1) for p in params])
        if world > 1:
            dist.all_reduce(flat)
        flat1 = torch.cat([p.grad.reshape(-1) for p in m1.parameters()])
        e_logit = float((lp - lp1[rows]).abs().max() / lp1.abs().max())
        e_grad = float((flat - flat1).abs().max() / flat1.abs().max())
        stat = torch.tensor([e_logit, e_grad], device=dev)
        if world > 1:
            dist.all_reduce(stat, op=dist.ReduceOp.MAX)
        check = {"logits_relerr": stat[0].item(), "weight_grad_relerr": stat[1].item()}
        del m1, g1, lp1
        opt.zero_grad(set_to_none=True)

    for _ in range(args.warmup):
        epoch()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for _ in range(args.epochs):
        v, t = epoch()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - w0) / args.epochs * 1e3
    ms = torch.tensor([e0.elapsed_time(e1) / args.epochs, wall], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    import rgb_experiment_b200.memo as memo
    timed = phase_ev[-args.epochs:]
    phases = [round(sum(ev[i].elapsed_time(ev[i + 1]) for ev in timed) / len(timed), 2) for i in range(3)]
    per_epoch = [[round(ev[i].elapsed_time(ev[i + 1]), 1) for i in range(3)] for ev in phase_ev]
    res = {"config": "APPNPStack full-batch epoch (1 train fwd+bwd+Adam, 2 eval fwd), "
                     f"{args.workload}-shaped, hidden {args.hidden}, K={args.K}",
           "n_gpus": world, "grid": "1x1" if grid is None else f"{grid.Pr}x{grid.Pf}",
           "epoch_ms": round(ms[0].item(), 2), "epoch_wall_ms": round(ms[1].item(), 2),
           "phase_ms": {"train_fwd_bwd_step": phases[0], "eval_val": phases[1], "eval_test": phases[2]},
           "phase_ms_per_epoch_incl_warmup": per_epoch if os.environ.get("RGBMP_EPOCH_TRACE") else None,
           "eval_memo": memo.enabled(), "memo_stats": dict(memo.stats) if world == 1 else {"hits": dist_prop.memo_hits},
           "val_acc": round(v[0], 4), "test_acc": round(t[0], 4),
           "val_loss": round(v[1], 4), "check_vs_single_gpu": check}
    if world > 1:
        torch.cuda.synchronize()
        pf.close()
        pb.close()
    return res


if __name__ == "__main__":
    main()
