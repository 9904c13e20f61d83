"""f2 on the device: the repeated no-grad forward of the K-hop / single-hop / GAT ops is served
from the exact memo and stays bit-identical to a fresh kernel run."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture()
def setup():
    import rgb_experiment_b200 as P
    import rgb_experiment_b200.memo as M
    import rgb_experiment_b200.synth as S
    M.clear()
    for k in M.stats:
        M.stats[k] = 0
    sg = S.make_graph(200_000, 3_400_000, 16, 5, seed=11, device=DEV)
    yield P, M, sg
    M.clear()


def test_appnp_eval_forward_twice_runs_the_kernels_once(setup):
    P, M, sg = setup
    g = P.Graph(sg.edge_index, sg.num_nodes, P.LOOP_ADD_REMAINING)
    x = sg.x
    fresh = P.ops.appnp(x, g, 10, 0.1)                     # autograd enabled: never memoised
    assert M.stats["hits"] == 0 and M.stats["misses"] == 0
    with torch.no_grad():
        a = P.ops.appnp(x, g, 10, 0.1)
        b = P.ops.appnp(x.clone(), g, 10, 0.1)
        assert M.stats["misses"] == 1 and M.stats["hits"] == 1
        assert torch.equal(a, fresh) and torch.equal(b, fresh) and a.data_ptr() != b.data_ptr()
        x2 = x.clone()
        x2[123, 4] += 0.5                                  # "the optimiser stepped": values differ
        c = P.ops.appnp(x2, g, 10, 0.1)
        assert M.stats["misses"] == 2
    assert torch.equal(c, P.ops.appnp(x2, g, 10, 0.1)) and not torch.equal(c, a)


def test_propagate_and_gat_memo(setup):
    P, M, sg = setup
    n = sg.num_nodes
    g = P.Graph(sg.edge_index, n, P.LOOP_REMOVE_THEN_ADD)
    x = torch.randn(n, 128, device=DEV)
    a_s, a_d = torch.randn(n, 8, device=DEV), torch.randn(n, 8, device=DEV)
    xp = x[:, :64].contiguous()
    ref_p = P.ops.propagate(x, g, "mean")
    ref_g = P.ops.gat(xp, a_s, a_d, g, 8, 8, 0.2)
    with torch.no_grad():
        for _ in range(2):
            assert torch.equal(P.ops.propagate(x, g, "mean"), ref_p)
            assert torch.equal(P.ops.gat(xp, a_s, a_d, g, 8, 8, 0.2), ref_g)
        assert M.stats["hits"] == 2 and M.stats["misses"] == 2
        # same inputs, different static argument: its own entry
        assert not torch.equal(P.ops.propagate(x, g, "sum"), ref_p)
        assert M.stats["misses"] == 3


def test_small_graphs_skip_the_memo(setup):
    P, M, _ = setup
    import rgb_experiment_b200.synth as S
    sg = S.make_graph(2708, 10556, 64, 7, seed=3, device=DEV)          # Cora-shaped: bookkeeping would dominate
    g = P.Graph(sg.edge_index, sg.num_nodes, P.LOOP_ADD_REMAINING)
    with torch.no_grad():
        P.ops.propagate(sg.x, g, "gcn")
        P.ops.propagate(sg.x, g, "gcn")
    assert M.stats["hits"] == 0 and M.stats["misses"] == 0 and M.stats["skipped"] == 2
