"""world_size-2 gloo test (CPU) of the multi-GPU host logic: row ranges, edge bucketing, the
per-hop all-gather exchange and the K-hop driver, with the CPU oracle standing in for the SpMM
kernel.  The partitioned result must equal the single-process oracle bit for bit."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import rgb_experiment_b200.partition as PT
        from oracle import pyg_restated as R
        torch.manual_seed(0)
        N, F, K, alpha = 101, 6, 5, 0.1                       # N not divisible by the world size
        ei = torch.randint(N, (2, 900))
        ed, w = R.gcn_norm(ei, None, N, dtype=torch.float32)
        z0 = torch.randn(N, F)
        Rr = PT.rows_per_rank(N, world)
        lo, hi = PT.row_range(N, rank, world)
        m = (ed[1] >= lo) & (ed[1] < hi)
        key, src = PT.local_edges(ed[0], ed[1], lo, hi)
        assert torch.equal(key.long() + lo, ed[1][m]) and torch.equal(src.long(), ed[0][m])
        wl = w[m]

        def spmm(x_full, z0_local, a, b):
            msg = wl.view(-1, 1) * x_full[src.long()]
            out = torch.zeros(Rr, x_full.size(1)).index_add_(0, key.long(), msg)
            out = out * a
            return out + b * z0_local

        z0_local = torch.zeros(Rr, F)
        z0_local[: hi - lo] = z0[lo:hi]
        drv = PT.PartitionedPropagator(N, rank, world, spmm)
        out_local = drv.run(z0_local, K, 1 - alpha, alpha)
        full = torch.empty(Rr * world, F)
        dist.all_gather_into_tensor(full, out_local)
        ref = R.appnp_propagate(z0, ei, K, alpha)
        ok = torch.equal(full[:N], ref)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_partitioned_khop_matches_single_process_oracle():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert ret.get(0) is True and ret.get(1) is True


def test_row_ranges_cover_all_rows():
    import rgb_experiment_b200.partition as PT
    for N in (1, 7, 100, 101, 2_449_029):
        for P in (1, 2, 4, 8):
            seen = 0
            for r in range(P):
                lo, hi = PT.row_range(N, r, P)
                assert lo <= hi and hi - lo <= PT.rows_per_rank(N, P)
                seen += hi - lo
            assert seen == N


def _grid_worker(rank, world, port, ret):
    """2 row blocks x 2 feature slices: ranks sharing a feature slice exchange rows, feature slices
    never communicate; the assembled result equals the single-process oracle bit for bit."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import rgb_experiment_b200.partition as PT
        from oracle import pyg_restated as R
        torch.manual_seed(0)
        N, F, K, alpha = 103, 10, 4, 0.1
        ei = torch.randint(N, (2, 1200))
        ed, w = R.gcn_norm(ei, None, N, dtype=torch.float32)
        z0 = torch.randn(N, F)
        grid = PT.Grid(rank, world, 2)
        assert (grid.Pr, grid.Pf, grid.rp, grid.fp) == (2, 2, rank // 2, rank % 2)
        flo, fhi = grid.feature_slice(F)
        assert (flo, fhi) == ((0, 8) if grid.fp == 0 else (8, 10))
        assert [PT.Grid.feature_slice(grid, 47, fp=f) for f in range(2)] == [(0, 24), (24, 47)]
        Rr = PT.rows_per_rank(N, grid.Pr)
        lo, hi = PT.row_range(N, grid.rp, grid.Pr)
        key, src = PT.local_edges(ed[0], ed[1], lo, hi)
        wl = w[(ed[1] >= lo) & (ed[1] < hi)]

        def spmm(x_full, z0_local, a, b):
            out = torch.zeros(Rr, x_full.size(1)).index_add_(0, key.long(), wl.view(-1, 1) * x_full[src.long()])
            return out * a + b * z0_local

        z0_local = torch.zeros(Rr, fhi - flo)
        z0_local[: hi - lo] = z0[lo:hi, flo:fhi]
        drv = PT.PartitionedPropagator(N, grid.rp, grid.Pr, spmm, group=grid.row_group)
        out_local = drv.run(z0_local, K, 1 - alpha, alpha)
        pad = torch.zeros(Rr, 8)
        pad[:, : fhi - flo] = out_local
        allb = torch.empty(world * Rr, 8)
        dist.all_gather_into_tensor(allb, pad)
        allb = allb.view(world, Rr, 8)
        full = torch.empty(Rr * grid.Pr, F)
        for r in range(world):
            a, b = (0, 8) if r % 2 == 0 else (8, 10)
            full[(r // 2) * Rr:(r // 2 + 1) * Rr, a:b] = allb[r, :, : b - a]
        ref = R.appnp_propagate(z0, ei, K, alpha)
        ret[rank] = bool(torch.equal(full[:N], ref))
    finally:
        dist.destroy_process_group()


def test_row_by_feature_grid_matches_single_process_oracle():
    world = 4
    port = 31500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_grid_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert all(ret.get(r) is True for r in range(world))


def _dist_autograd_worker(rank, world, port, ret):
    """DistAPPNP (differentiable K-hop over a 2x2 grid) with the CPU oracle standing in for the
    kernels: forward rows and the gradient w.r.t. the input equal single-process autograd."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import rgb_experiment_b200.partition as PT
        from oracle import pyg_restated as R
        torch.manual_seed(0)
        N, F, K, alpha = 97, 12, 3, 0.1
        ei = torch.randint(N, (2, 800))                       # directed: the transpose really differs
        ed, w = R.gcn_norm(ei, None, N, dtype=torch.float64)
        h_full = torch.randn(N, F, dtype=torch.float64)
        wgt = torch.randn(N, F, dtype=torch.float64)           # loss = sum(z * wgt)
        grid = PT.Grid(rank, world, 2)
        Rr = PT.rows_per_rank(N, grid.Pr)
        lo, hi = PT.row_range(N, grid.rp, grid.Pr)

        def make_runner(by_src):
            tgt, oth = (ed[0], ed[1]) if by_src else (ed[1], ed[0])
            m = (tgt >= lo) & (tgt < hi)
            key, col, wl = (tgt[m] - lo), oth[m], w[m]

            def spmm(x_full, z0_local, a, b):
                out = torch.zeros(Rr, x_full.size(1), dtype=x_full.dtype).index_add_(0, key, wl.view(-1, 1) * x_full[col])
                return out * a + b * z0_local

            drv = PT.PartitionedPropagator(N, grid.rp, grid.Pr, spmm, group=grid.row_group)
            return lambda t, K_, al: drv.run(t, K_, 1 - al, al)

        a, b = grid.feature_slice(F)
        mod = PT.DistAPPNP(grid, F, K, alpha, make_runner(False), make_runner(True), col_group=grid.col_group)
        assert mod.ld == 8
        h_loc = torch.zeros(Rr, F, dtype=torch.float64)
        h_loc[: hi - lo] = h_full[lo:hi]
        h_loc.requires_grad_(True)
        z = mod(h_loc)
        wl_ = torch.zeros(Rr, F, dtype=torch.float64)
        wl_[: hi - lo] = wgt[lo:hi]
        # every rank of a row block sees the same full-width rows; its backward only feeds its own slice
        (z * wl_).sum().backward()
        g = h_loc.grad.clone()
        dist.all_reduce(g, group=grid.col_group)               # sum the slice-wise gradients of the row block
        ho = h_full.clone().requires_grad_(True)
        zo = R.appnp_propagate(ho, ei, K, alpha)
        (zo * wgt).sum().backward()
        ok_f = torch.allclose(z[: hi - lo].detach(), zo[lo:hi].detach(), rtol=1e-12, atol=1e-12)
        ok_b = torch.allclose(g[: hi - lo], ho.grad[lo:hi], rtol=1e-12, atol=1e-12)
        # f2: under no_grad the second identical forward is served from the memo on EVERY rank; when one rank's rows
        # change, every rank recomputes (the decision is all-reduced, nobody skips a collective)
        with torch.no_grad():
            e1 = mod(h_loc.detach())
            e2 = mod(h_loc.detach().clone())
            hits_after_repeat = mod.memo_hits
            h2 = h_loc.detach().clone()
            if rank == 0:
                h2[0, a] += 1.0
            e3 = mod(h2)
        ok_m = (hits_after_repeat == 1 and mod.memo_hits == 1 and torch.equal(e1, e2)
                and torch.equal(e1, z.detach()))
        changed = torch.tensor([0.0 if torch.equal(e3, e1) else 1.0])
        dist.all_reduce(changed)
        ret[rank] = bool(ok_f and ok_b and ok_m and changed.item() >= 1.0)
    finally:
        dist.destroy_process_group()


def test_dist_appnp_autograd_on_a_grid_matches_single_process():
    world = 4
    port = 33500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_dist_autograd_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert all(ret.get(r) is True for r in range(world))


def _relabel_worker(rank, world, port, ret):
    """Community row blocks: every rank renames the nodes from the same (replicated) group vector, buckets the renamed
    edges, propagates its block; the gathered result read back through inv equals the single-process oracle on the
    ORIGINAL graph (same operator under another node naming; rounding differs only by the order of a row's edges,
    which the renaming keeps)."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import rgb_experiment_b200.partition as PT
        from oracle import pyg_restated as R
        torch.manual_seed(0)
        N, F, K, alpha, S = 101, 6, 5, 0.1, 7
        ei = torch.randint(N, (2, 900))
        group = torch.randint(S, (N,), dtype=torch.int32)
        z0 = torch.randn(N, F)
        perm, inv = PT.community_naming(group, N)
        ed, w = R.gcn_norm(inv[ei], None, N, dtype=torch.float32)        # the renamed graph: edge order unchanged
        Rr = PT.rows_per_rank(N, world)
        lo, hi = PT.row_range(N, rank, world)
        g_new = group[perm]
        if rank + 1 < world and hi < N:                                 # a block boundary never splits ids out of group order
            assert int(g_new[hi - 1]) <= int(g_new[hi])
        key, src = PT.local_edges(ed[0], ed[1], lo, hi)
        wl = w[(ed[1] >= lo) & (ed[1] < hi)]

        def spmm(x_full, z0_local, a, b):
            out = torch.zeros(Rr, x_full.size(1)).index_add_(0, key.long(), wl.view(-1, 1) * x_full[src.long()])
            return out * a + b * z0_local

        z0_local = torch.zeros(Rr, F)
        z0_local[: hi - lo] = z0[perm[lo:hi]]                           # my rows are the nodes perm[lo:hi] of the caller
        drv = PT.PartitionedPropagator(N, rank, world, spmm)
        out_local = drv.run(z0_local, K, 1 - alpha, alpha)
        full = torch.empty(Rr * world, F)
        dist.all_gather_into_tensor(full, out_local)
        ref = R.appnp_propagate(z0, ei, K, alpha)
        ret[rank] = bool(torch.equal(full[:N][inv], ref))
    finally:
        dist.destroy_process_group()


def test_community_row_blocks_match_the_single_process_oracle():
    world = 2
    port = 31500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_relabel_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert ret.get(0) is True and ret.get(1) is True
