#!/usr/bin/env python
"""bench.py -- APPNP K=10 propagation on the ogbn-products-shaped synthetic graph
(BASELINE.json configs[3], the configuration the metric "GCN/APPNP propagate GTEPS & HBM GB/s vs
peak ... at 1/2/4/8 GPU" is quoted on; it fits one GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch: z = APPNP(z0) = 10 fused hops
z <- 0.9 * A_hat z + 0.1 * z0 over nnz = E + N edges, z0 [N, 47] fp32 (appnp_stack.py:22,30).
value = aggregated edges / s (GTEPS) over all ranks, inputs resident in HBM.
e2e   = the same metric through the host-buffer C-ABI call (rgbmp_appnp_host): H2D of z0 from
        pinned memory, 10 hops, D2H of z inside the timed region.
--impl reference times the reference's CPU path (the oracle's restated PyG gather -> scale ->
scatter_add_, this tier's definition) on the host cores, rank 0 only.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = "products"
K_HOPS, ALPHA, F_CLASSES = 10, 0.1, 47
GRAPH_SEED = 20261018
CPU_SAMPLE_FRACTION = 0.10          # share of the target rows used by the bounded CPU sample


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def algorithmic_bytes_per_hop(nnz: int, n: int, F: int, weighted: bool) -> int:
    """SURVEY.md 8d gather model, no cache credit: nnz*(F*4 + 4 [col] + s_w) + N*F*4 [store]
    + (N+1)*8 [rowptr] + N*F*4 [teleport read]."""
    return nnz * (F * 4 + 4 + (4 if weighted else 0)) + n * F * 4 + (n + 1) * 8 + n * F * 4


def workload_name(N: int, nnz: int, F: int) -> str:
    """config.workload, shared by both arms (the reference arm runs a bounded sample of the same workload)."""
    return (f"APPNP K={K_HOPS} alpha={ALPHA} propagate, ogbn-products-shaped R-MAT graph "
            f"(N={N}, E={nnz - N} directed + {N} self loops, F={F} fp32)")


def make_workload(device):
    import rgb_experiment_b200.synth as S
    return S.make_named(WORKLOAD, seed=GRAPH_SEED, device=device, features=False)


def cpu_sample(sg, device):
    """Bounded CPU sample of the same workload: the edges (after add_remaining_self_loops) whose
    target is one of the first 10% of the nodes, with their gcn_norm weights, and z [N, 47]."""
    N = sg.num_nodes
    ei = sg.edge_index
    loop = torch.arange(N, device=ei.device)
    src = torch.cat([ei[0], loop])
    dst = torch.cat([ei[1], loop])
    deg = torch.bincount(dst, minlength=N).float()
    dinv = deg.pow(-0.5)
    n_sub = int(N * CPU_SAMPLE_FRACTION)
    m = dst < n_sub
    s, d = src[m], dst[m]
    w = dinv[s] * dinv[d]
    z = torch.randn(N, F_CLASSES, generator=torch.Generator().manual_seed(1))
    return s.cpu(), d.cpu(), w.cpu(), z, n_sub


def cpu_hop(s, d, w, z, n_sub, alpha=ALPHA):
    """The literal PyG-on-CPU form (oracle.pyg_restated.propagate): index_select -> mul -> scatter_add_."""
    from oracle import pyg_restated as R
    out = R.scatter_add(w.view(-1, 1) * z.index_select(0, s), d, dim=0, dim_size=n_sub)
    out = out * (1 - alpha)
    return out + alpha * z[:n_sub]


def cpu_strong_baseline(s, d, w, z, n_sub, budget_s=10.0, alpha=ALPHA):
    """the same hop as one torch.sparse CSR product A @ z on the host cores -- no [nnz, F] message
    tensor, the strongest CPU form available in this image.  Reported beside the literal PyG form, never instead of it."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        perm = torch.argsort(d, stable=True)
        crow = torch.zeros(n_sub + 1, dtype=torch.int64)
        crow[1:] = torch.cumsum(torch.bincount(d, minlength=n_sub), 0)
        A = torch.sparse_csr_tensor(crow, s[perm], w[perm], size=(n_sub, z.size(0)))
        best, hops, t_all = None, 0, time.perf_counter()
        while hops < 6 and (hops < 2 or time.perf_counter() - t_all < budget_s):
            t0 = time.perf_counter()
            out = (A @ z) * (1 - alpha) + alpha * z[:n_sub]
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            hops += 1
    del out
    return {"value": s.numel() / best / 1e9, "unit": "GTEPS", "kind": "torch.sparse CSR A @ z on the same sample and weights",
            "best_of": hops}


def run_reference(args, rank):
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the other ranks have already left, so rank 0 takes every
    # host core it is allowed to run on (the same thread count a plain `python bench.py --impl reference` gets)
    try:
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        torch.set_num_threads(os.cpu_count() or 1)
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"
    sg = make_workload(dev)
    N_full, nnz_full = sg.num_nodes, sg.edge_index.size(1) + sg.num_nodes      # the generator emits no self loops
    s, d, w, z, n_sub = cpu_sample(sg, dev)
    del sg
    cores = torch.get_num_threads()
    for _ in range(args.warmup):
        cpu_hop(s, d, w, z, n_sub)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_hop(s, d, w, z, n_sub)
    dt = time.perf_counter() - t0
    val = s.numel() * args.steps / dt / 1e9
    sample = (f"1 hop per step over the {s.numel()} edges whose target is in the first "
              f"{int(CPU_SAMPLE_FRACTION * 100)}% of nodes; literal index_select->mul->scatter_add_ fp32")
    line = {"impl": "reference", "metric": "appnp_propagate_gteps", "value": val, "unit": "GTEPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(N_full, nnz_full, F_CLASSES), "hops_per_step": 1, "nnz": nnz_full,
                       "parallelism": f"{cores} host threads", "sample": sample},
            "cpu_baseline": {"value": val, "unit": "GTEPS", "cores": cores, "kind": "port", "sample": sample,
                             "strong": cpu_strong_baseline(s, d, w, z, n_sub)},
            "e2e": {"value": val, "unit": "GTEPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_epoch_guarded(rank, world, dev, line, limit_s=300):
    """The second half of the BASELINE metric ("full-batch epoch ms at 1/2/4/8 GPU"): the reference's APPNPStack
    epoch (1 train forward + backward + Adam, 2 eval forwards; itexperiments.py:417-473) on the same products-shaped
    graph, measured by tools/bench_epoch.py AFTER the timed region and reported under "epoch".  It never costs the
    headline line: an exception becomes {"error": ...}, and if nothing comes back within `limit_s` seconds (a rank
    stuck in a collective) a watchdog prints the line without it and ends the process with status 0."""
    import importlib.util
    done = threading.Event()

    def watchdog():
        if done.wait(limit_s):
            return
        if line is not None:
            line["epoch"] = {"error": f"no result within {limit_s} s"}
            print(json.dumps(line), flush=True)
        os._exit(0)

    threading.Thread(target=watchdog, daemon=True).start()
    try:
        spec = importlib.util.spec_from_file_location("_bench_epoch", os.path.join(ROOT, "tools", "bench_epoch.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        res = mod.run(mod.default_args(), rank, world, dev)
        res.pop("check_vs_single_gpu", None)
        if world == 1:
            torch.cuda.empty_cache()
            try:
                res["as_called"] = mod.run_as_called(mod.default_args(epochs=3, warmup=1), dev)
            except Exception as e:                                    # noqa: BLE001
                res["as_called"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    except Exception as e:                                        # noqa: BLE001 -- reported, never fatal for the line
        res = {"error": f"{type(e).__name__}: {e}"[:300]}
        if line is None:                                          # a peer may now be waiting for me: leave quietly
            done.set()
            os._exit(0)
    done.set()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-epoch", action="store_true",
                    help="skip the full-batch epoch measurement (the second half of the BASELINE metric) after the timed region")
    ap.add_argument("--fold", type=int, default=1,
                    help="1 (default): D^-1/2 (A+I) D^-1/2 applied as row scalings around an unweighted sum (no per-edge "
                         "weight stream); 0: per-edge gcn_norm weights exactly as PyG multiplies them")
    ap.add_argument("--feature-groups", type=int, default=0,
                    help="N>1: Pf of the Pr x Pf process grid (features split Pf ways, rows N/Pf ways); 0 = auto")
    ap.add_argument("--exchange", default="push", choices=["push", "allgather"],
                    help="N>1: fused push of finished rows into every peer over NVLink (default) or NCCL all-gather")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    import rgb_experiment_b200 as P
    import rgb_experiment_b200.partition as PT

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sg = make_workload(dev)
    N, F = sg.num_nodes, F_CLASSES
    _warm = P.Graph(torch.tensor([[0, 1, 2], [1, 2, 0]], device=dev), 3, P.LOOP_ADD_REMAINING)   # loads the build kernels
    _ = _warm.bwd
    del _warm
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if world == 1:
        g = P.get_graph(sg.edge_index, N, P.LOOP_ADD_REMAINING)      # what the layer's first forward does
        torch.cuda.synchronize()
        build_ms = (time.perf_counter() - t0) * 1e3
        nnz = g.nnz
        n_items = g.fwd.n_items
        z0 = torch.randn(N, F, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
        # the call the reference's APPNPStack makes (appnp_stack.py:22,30): the shim layer, its graph-cache lookup
        # and the autograd.Function are inside the timed region; --fold 1 is the layer's default form
        import rgb_experiment_b200.shim.nn as SN
        layer = SN.APPNP(K_HOPS, ALPHA)
        layer.fold_norm = bool(args.fold)
        ei = sg.edge_index                              # stays alive: the cached graph lives as long as this tensor

        def step():
            return layer(z0, ei)

        launches_per_step = K_HOPS * (1 + (2 if n_items > 0 else 0)) + (1 if args.fold else 0)
    else:
        Pf = args.feature_groups if args.feature_groups > 0 else PT.auto_feature_groups(world, F)
        grid = PT.Grid(rank, world, Pf)
        blk = PT.LocalBlock(sg.edge_index, N, P.LOOP_ADD_REMAINING, grid.rp, grid.Pr, group=grid.row_group)
        torch.cuda.synchronize()
        build_ms = (time.perf_counter() - t0) * 1e3
        nnz = blk.nnz_global
        flo, fhi = grid.feature_slice(F)
        F_local = fhi - flo
        prop = PT.PartitionedAPPNP(blk, F_local, group=grid.row_group, mode=args.exchange)
        z0l = torch.zeros((blk.R, prop.ld), device=dev)
        z0l[: blk.hi - blk.lo, :F_local] = torch.randn(blk.hi - blk.lo, F_local, device=dev,
                                                       generator=torch.Generator(device=dev).manual_seed(1 + rank))

        def step():
            return prop.run(z0l, K_HOPS, ALPHA)

        launches_per_step = K_HOPS * prop.launches_per_hop
    del sg

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    sampler.stop_flag = True
    sampler.join()
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    gteps = nnz * K_HOPS * args.steps / (ms * 1e-3) / 1e9

    # ---- e2e: host buffers through the C-ABI host entry point (single GPU) / host round trip (multi) ----
    e2e = None
    if world == 1:
        z0h = z0.cpu().pin_memory()
        outh = torch.empty_like(z0h).pin_memory()
        plan = P.ops.HostAppnpPlan(g, F)
        for _ in range(2):
            P.ops.appnp_host(g, z0h, outh, K_HOPS, ALPHA, plan)
        torch.cuda.synchronize()
        n_e2e = max(3, args.steps // 2)
        w0 = time.perf_counter()
        for _ in range(n_e2e):
            P.ops.appnp_host(g, z0h, outh, K_HOPS, ALPHA, plan)      # returns after the D2H completed
        wall = time.perf_counter() - w0
        e2e = {"value": nnz * K_HOPS * n_e2e / wall / 1e9, "unit": "GTEPS",
               "h2d_bytes_per_step": N * F * 4, "d2h_bytes_per_step": N * F * 4, "ms_per_step": wall / n_e2e * 1e3}
    else:
        R_, ldp = z0l.shape
        z0h = z0l.cpu().pin_memory()
        outh = torch.empty_like(z0h).pin_memory()
        n_e2e = max(3, args.steps // 2)
        barrier()
        w0 = time.perf_counter()
        for _ in range(n_e2e):
            z0l.copy_(z0h, non_blocking=True)
            outh.copy_(step(), non_blocking=True)
            torch.cuda.synchronize()
        barrier()
        wall = torch.tensor([time.perf_counter() - w0], device=dev)
        dist.all_reduce(wall, op=dist.ReduceOp.MAX)
        wall = float(wall.item())
        e2e = {"value": nnz * K_HOPS * n_e2e / wall / 1e9, "unit": "GTEPS",
               "h2d_bytes_per_step": R_ * ldp * 4 * world, "d2h_bytes_per_step": R_ * ldp * 4 * world,
               "ms_per_step": wall / n_e2e * 1e3}

    if world > 1:
        torch.cuda.synchronize()
        prop.close()
        del prop, blk
    else:
        del plan, g
    torch.cuda.empty_cache()
    if rank != 0:
        if not args.no_epoch:
            run_epoch_guarded(rank, world, dev, None)
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_kind = peaks()
    weighted = (args.exchange == "allgather") if world > 1 else not args.fold   # the push path folds D^-1/2 too
    if world > 1:                                      # per GPU: nnz/Pr edges of F/Pf-wide rows
        hop_bytes = algorithmic_bytes_per_hop(nnz // grid.Pr, N // grid.Pr, F_local, weighted=weighted)
    else:
        hop_bytes = algorithmic_bytes_per_hop(nnz, N, F, weighted=weighted)
    hop_ms = ms_per_step / K_HOPS
    achieved = hop_bytes / (hop_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_kind": peak_kind, "kernel": "spmm_rows_kernel<float,4,...> (+long/combine)",
                "bytes_per_launch": hop_bytes, "launch_ms": hop_ms}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            # one "launch" of the roofline object = one hop = rows + long-row + combine kernels: report their summed DRAM bytes
            roofline["traffic"] = json.load(open(tr)).get("hop_bytes_total")
        except Exception:
            pass

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        sg2 = make_workload(dev)
        s, d, w, z, n_sub = cpu_sample(sg2, dev)
        del sg2
        cpu_hop(s, d, w, z, n_sub)
        best = None
        tot0 = time.perf_counter()
        hops = 0
        while hops < 10 and time.perf_counter() - tot0 < 25:
            c0 = time.perf_counter()
            cpu_hop(s, d, w, z, n_sub)
            dt = time.perf_counter() - c0
            best = dt if best is None else min(best, dt)
            hops += 1
        cpu_baseline = {"value": s.numel() / best / 1e9, "unit": "GTEPS", "cores": torch.get_num_threads(),
                        "kind": "port",
                        "sample": f"best of {hops} single hops over the {s.numel()} edges whose target is in the first "
                                  f"{int(CPU_SAMPLE_FRACTION * 100)}% of nodes; literal index_select->mul->scatter_add_ fp32",
                        "strong": cpu_strong_baseline(s, d, w, z, n_sub)}

    line = {"metric": "appnp_propagate_gteps", "value": gteps, "unit": "GTEPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(N, nnz, F),
                       "hops_per_step": K_HOPS, "nnz": nnz, "parallelism": (f"{grid.Pr} row blocks x {grid.Pf} feature slices, exchange={args.exchange}"
                                       if world > 1 else "single"),
                       "l2": "inputs larger than L2 (features 470 MB, col 505 MB vs 126 MB L2)",
                       "norm": "per-edge weights" if weighted else "folded row scaling (same operator, no per-edge weight stream)", "graph_build_ms": build_ms},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "gpu_launches": launches_per_step * args.steps, "clocks": sampler.result()}
    if not args.no_epoch:
        line["epoch"] = run_epoch_guarded(rank, world, dev, line)
    print(json.dumps(line), flush=True)
    if "error" in (line.get("epoch") or {}):
        os._exit(0)                    # a failed epoch measurement may have left a collective half-entered
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
