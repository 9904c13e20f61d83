"""Row-generated graphs (config C5): the generator is a pure function of (seed, row, k) -- the same
graph for every partition count -- and the CUDA path on such a CSR-by-construction matches the
oracle (bf16 storage, fp32 accumulation: <= 2^-8 norm-wise; fp32: <= 1e-5)."""
import pytest
import torch

from helpers import relerr


def test_rowgen_is_partition_invariant_and_has_the_named_shape():
    import rgb_experiment_b200.synth as S
    N, E = 50_000, 1_450_000
    rp, col = S.rowgen_block(N, E, 0, N)
    assert rp.dtype == torch.int64 and col.dtype == torch.int32 and rp[0] == 0 and rp[-1] == col.numel()
    assert abs(col.numel() / (E + N) - 1) < 0.05                    # mean in-degree ~ E/N, plus the self loops
    assert int(col.min()) >= 0 and int(col.max()) < N
    assert torch.equal(col[rp[:-1]].long(), torch.arange(N))         # every row starts with its self loop
    cuts = [0, 12_345, 30_000, N]
    parts = [S.rowgen_block(N, E, a, b) for a, b in zip(cuts[:-1], cuts[1:])]
    assert torch.equal(torch.cat([c for _, c in parts]), col)
    deg = torch.cat([r[1:] - r[:-1] for r, _ in parts])
    assert torch.equal(deg, rp[1:] - rp[:-1])
    rp2, col2 = S.rowgen_block(N, E, 0, N, locality=0.9)
    near = ((col2.long() - torch.repeat_interleave(torch.arange(N), rp2[1:] - rp2[:-1])).abs() % N)
    near = torch.minimum(near, N - near) <= N // 64
    assert near.float().mean() > 0.85


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 2 ** -8), (torch.float32, 1e-5)])
def test_rowgen_block_propagation_matches_oracle(dtype, tol):
    import rgb_experiment_b200 as P
    import rgb_experiment_b200.partition as PT
    import rgb_experiment_b200.synth as S
    from oracle import pyg_restated as R
    dev = torch.device("cuda:0")
    N, E, F, K = 20_000, 580_000, 128, 2
    blk = PT.LocalBlock.from_rowgen(N, E, 0, 1, device=dev)
    rp, col = blk.csr.rowptr.cpu(), blk.csr.col.cpu().long()
    rp_c, col_c = S.rowgen_block(N, E, 0, N)                           # CPU generation = GPU generation
    assert torch.equal(rp, rp_c) and torch.equal(col, col_c.long())
    dst = torch.repeat_interleave(torch.arange(N), rp[1:] - rp[:-1])
    ei = torch.stack([col, dst])                                       # self loops are already in the list
    _, w = R.gcn_norm(ei, None, N, add_self_loops=False, dtype=torch.float32)
    x = torch.randn(N, F, generator=torch.Generator().manual_seed(0))
    xq = x.to(dtype)
    ref = xq.double()
    for _ in range(K):
        ref = R.propagate(ei, ref, w.double(), "add", N)
    prop = PT.PartitionedAPPNP(blk, F, dtype=dtype)
    z0 = torch.zeros((blk.R, prop.ld), dtype=dtype, device=dev)
    z0[:, :F] = xq.to(dev)
    out = prop.run(z0, K, 0.0)[:N, :F]
    assert relerr(out.float(), ref) <= tol * K


@pytest.mark.gpu
@pytest.mark.parametrize("relabel", [False, True])
def test_local_block_with_community_naming_is_the_same_operator(relabel, monkeypatch):
    """LocalBlock(relabel=True) renames the nodes by (locality group, id) before bucketing the rows: row i of every
    operand / result is node perm[i] of the caller's numbering, inv is the inverse, and the propagation is the same
    operator -- K hops on the block (world size 1: the whole renamed graph) equal ops.appnp on the original graph,
    read through perm."""
    import rgb_experiment_b200 as P
    import rgb_experiment_b200.graph as G_
    import rgb_experiment_b200.partition as PT
    import rgb_experiment_b200.synth as S
    monkeypatch.setattr(G_, "CLUSTER", "1")                # the test graph is far below the size where grouping pays
    dev = torch.device("cuda:0")
    sg = S.make_graph(30_000, 600_000, 8, 16, seed=5, device=dev, features=False)
    N, F, K, alpha = sg.num_nodes, 10, 4, 0.1
    blk = PT.LocalBlock(sg.edge_index, N, P.LOOP_ADD_REMAINING, 0, 1, relabel=relabel)
    x = torch.randn(N, F, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
    want = P.ops.appnp(x, P.Graph(sg.edge_index, N, P.LOOP_ADD_REMAINING), K, alpha)
    if relabel:
        assert blk.perm is not None and torch.equal(blk.inv[blk.perm], torch.arange(N, device=dev))
        grp = blk.groups[0].long()
        assert bool((grp[1:] >= grp[:-1]).all())           # a row range is a run of whole groups
        x_in, want = x[blk.perm], want[blk.perm]
    else:
        assert blk.perm is None and blk.inv is None
        x_in = x
    prop = PT.PartitionedAPPNP(blk, F)
    z0 = torch.zeros((blk.R, prop.ld), device=dev)
    z0[:, :F] = x_in
    got = prop.run(z0, K, alpha)[:N, :F]
    err = float((got - want).abs().max() / want.abs().max())
    assert err <= 1e-6, err
