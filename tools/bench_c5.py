#!/usr/bin/env python
"""BASELINE.json configs[4]: GCN propagation with bf16 features on the papers100M-shaped synthetic
graph (111 M nodes / 3.2 B directed edges / 128 features), row-partitioned over 2/4/8 B200s.
Each rank generates its own CSR row block on the device (synth.rowgen_block), the iterate lives in
peer-mapped buffers and every hop is ONE fused SpMM whose epilogue pushes the finished rows to the
ranks of its row group over NVLink.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29531 tools/bench_c5.py [--scale 1.0] [--hops 2] [--steps 3] [--feature-groups 1]
--scale shrinks nodes and edges together (1-GPU and smoke runs)."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--hops", type=int, default=2)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--feature-groups", type=int, default=1)
    ap.add_argument("--exchange", default="push", choices=["push", "allgather"])
    ap.add_argument("--locality", type=float, default=0.0)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    lr = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import rgb_experiment_b200.partition as PT
    import rgb_experiment_b200.synth as S
    N0, E0, F = S.PAPERS100M
    N, E = int(N0 * args.scale), int(E0 * args.scale)
    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    esz = 2 if dt == torch.bfloat16 else 4
    grid = PT.Grid(rank, world, args.feature_groups)
    t0 = time.perf_counter()
    blk = PT.LocalBlock.from_rowgen(N, E, grid.rp, grid.Pr, group=grid.row_group, device=dev, locality=args.locality)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    flo, fhi = grid.feature_slice(F, align=16 // esz)
    Fl = fhi - flo
    prop = PT.PartitionedAPPNP(blk, Fl, group=grid.row_group, mode=args.exchange, dtype=dt)
    z0l = torch.zeros((blk.R, prop.ld), dtype=dt, device=dev)
    z0l[: blk.hi - blk.lo, :Fl] = torch.randn(blk.hi - blk.lo, Fl, device=dev,
                                              generator=torch.Generator(device=dev).manual_seed(1 + rank)).to(dt)
    def timed(hops):
        for _ in range(args.warmup):
            prop.run(z0l, hops, 0.0)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            o = prop.run(z0l, hops, 0.0)
        e1.record()
        torch.cuda.synchronize()
        tt = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()) / args.steps, o

    step_ms, out = timed(args.hops)
    step_ms2, _ = timed(args.hops + 2)
    t = torch.tensor([step_ms * args.steps], device=dev)
    finite = torch.tensor([1 if bool(torch.isfinite(out.float()).all()) else 0], device=dev)
    mem = torch.tensor([torch.cuda.max_memory_allocated(dev) / 2**30], device=dev)
    if world > 1:
        dist.all_reduce(finite, op=dist.ReduceOp.MIN)
        dist.all_reduce(mem, op=dist.ReduceOp.MAX)
    marginal = (step_ms2 - step_ms) / 2.0          # steady-state hop: the initial distribution of the iterate cancels
    ms_hop = float(t.item()) / (args.steps * args.hops)
    nnz = blk.nnz_global * (1 if grid.Pf == 1 else 1)        # every feature group walks the same edges once per hop
    if rank == 0:
        per_gpu_bytes = (blk.nnz_local * (Fl * esz + 4 + 4) + blk.R * Fl * esz + (blk.R + 1) * 8)
        line = {"config": "C5 GCN propagation, papers100M-shaped row-generated graph", "scale": args.scale, "N": N,
                "nnz": blk.nnz_global, "F": F, "dtype": args.dtype, "n_gpus": world, "grid": f"{grid.Pr}x{grid.Pf}",
                "exchange": args.exchange, "locality": args.locality, "hops": args.hops, "ms_per_hop": round(ms_hop, 3),
                "gteps": round(blk.nnz_global / ms_hop / 1e6, 2),
                "steady_state_ms_per_hop": round(marginal, 3), "steady_state_gteps": round(blk.nnz_global / marginal / 1e6, 2),
                "per_gpu_algorithmic_GBps": round(per_gpu_bytes / ms_hop / 1e6, 1),
                "nvlink_rx_MB_per_hop_per_gpu": round((grid.Pr - 1) * blk.R * prop.ld * esz / 1e6, 1),
                "build_s": round(build_s, 2), "torch_peak_mem_GiB": round(float(mem.item()), 1),
                "peer_buffers_GiB": round(2 * blk.R * grid.Pr * prop.ld * esz / 2**30, 1),
                "finite": bool(finite.item())}
        print(json.dumps(line), flush=True)
    torch.cuda.synchronize()
    prop.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
