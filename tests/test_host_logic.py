"""Host-side logic that needs no GPU: the chaining of locality groups, the long-row thresholds, the non-view
output buffers, head padding of the fused attention family, the byte-capped weakref graph cache key."""
import numpy as np
import torch


def test_chain_groups_puts_strongly_connected_groups_next_to_each_other():
    from rgb_experiment_b200.graph import chain_groups
    rng = np.random.default_rng(0)
    S, B = 240, 12
    block = np.arange(S) % B                                  # 12 planted communities of 20 groups each, interleaved ids
    W = rng.random((S, S)) * 5 + (block[:, None] == block[None, :]) * 400.0
    rank, intra = chain_groups(W)
    assert sorted(rank.tolist()) == list(range(S))            # a permutation
    seq = block[np.argsort(rank)]
    assert int((seq[1:] != seq[:-1]).sum()) == B - 1          # every community is one contiguous run of the chain
    assert 0.0 < intra < 0.2                                  # share of the weight on the diagonal (each group alone)


def test_chain_groups_handles_empty_and_single_group():
    from rgb_experiment_b200.graph import chain_groups
    rank, intra = chain_groups(np.zeros((3, 3)))
    assert sorted(rank.tolist()) == [0, 1, 2] and intra == 0.0
    rank, intra = chain_groups(np.array([[7.0]]))
    assert rank.tolist() == [0] and intra == 1.0


def test_small_graphs_split_long_rows_early():
    from rgb_experiment_b200 import graph as G
    assert G.default_chunks(2708) == (G.SMALL_CHUNK, G.SMALL_LONG_CHUNK)
    assert G.default_chunks(2_449_029) == (G.DEFAULT_CHUNK, G.DEFAULT_LONG_CHUNK)
    assert G.SMALL_CHUNK < G.DEFAULT_CHUNK and G.SMALL_LONG_CHUNK % G.SMALL_CHUNK == 0


def test_function_outputs_are_strided_tensors_not_views():
    """graphsage.py:60 updates the propagate result in place; autograd forbids that on a view made inside a Function."""
    from rgb_experiment_b200 import ops
    t, ld = ops.alloc_rows(5, 10, torch.float32, "cpu")
    assert ld == 12 and t.shape == (5, 10) and t.stride() == (12, 1) and not t._is_view()
    t.zero_()
    t += 1.0                                                  # in-place is fine
    assert float(t.sum()) == 50.0
    t2, ld2 = ops.alloc_rows(5, 8, torch.float32, "cpu")
    assert ld2 == 8 and t2.is_contiguous() and not t2._is_view()
    tb, ldb = ops.alloc_rows(3, 9, torch.bfloat16, "cpu")
    assert ldb == 16 and tb.stride() == (16, 1)


def test_head_padding_of_the_fused_attention_family():
    from rgb_experiment_b200 import ops
    assert [ops._head_pad(1, c) for c in (1, 7, 41, 128)] == [1, 7, 41, 128]        # a single head keeps its width
    assert [ops._head_pad(8, c) for c in (4, 8, 10, 16, 47, 64, 100)] == [8, 8, 16, 16, 64, 64, 128]


def test_graph_cache_key_distinguishes_what_it_must():
    from rgb_experiment_b200.graph import _graph_key
    ei = torch.zeros((2, 6), dtype=torch.long)
    k = _graph_key(ei, 4, 2, False)
    assert k != _graph_key(ei, 4, 2, True) and k != _graph_key(ei, 5, 2, False) and k != _graph_key(ei, 4, 3, False)
    assert k != _graph_key(ei.t().contiguous().t(), 4, 2, False)          # other storage / strides
    ei[0, 0] = 1                                                           # in-place edit bumps _version
    assert k != _graph_key(ei, 4, 2, False)


def test_community_naming_makes_row_blocks_runs_of_whole_groups_and_keeps_the_operator():
    """partition.community_naming: perm / inv are inverse permutations, the groups are non-decreasing under the new
    names (equal-sized row blocks are runs of whole groups), ids keep their order inside a group, and propagating on
    the renamed edge list is the same operator read through perm (checked with the CPU oracle)."""
    import rgb_experiment_b200.partition as PT
    from oracle import pyg_restated as R
    gen = torch.Generator().manual_seed(0)
    N, S = 97, 5
    group = torch.randint(S, (N,), generator=gen, dtype=torch.int32)
    perm, inv = PT.community_naming(group, N)
    ids = torch.arange(N)
    assert torch.equal(inv[perm], ids) and torch.equal(perm[inv], ids)
    g_new = group[perm].long()
    assert bool((g_new[1:] >= g_new[:-1]).all())
    same = g_new[1:] == g_new[:-1]
    assert bool((perm[1:][same] > perm[:-1][same]).all())
    ei = torch.randint(N, (2, 400), generator=gen)
    x = torch.randn(N, 6, generator=gen, dtype=torch.float64)
    z = R.appnp_propagate(x, ei, 3, 0.1)
    z_new = R.appnp_propagate(x[perm], inv[ei], 3, 0.1)
    assert torch.allclose(z_new, z[perm], rtol=0, atol=1e-12)


def test_block_schedule_rule():
    import rgb_experiment_b200.partition as PT
    assert PT.block_keeps_groups(1_224_515, 47 * 4) and PT.block_keeps_groups(1_224_515, 24 * 4)      # 2x1, 2x2
    assert PT.block_keeps_groups(612_258, 47 * 4)                                                      # 4x1
    assert not PT.block_keeps_groups(612_258, 24 * 4) and not PT.block_keeps_groups(306_129, 47 * 4)   # 4x2, 8x1
    assert PT.block_keeps_groups(1_000_000, None) and not PT.block_keeps_groups(999_999, None)
