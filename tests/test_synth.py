"""The synthetic inputs of SURVEY.md 8d (rgb-experiment_b200/synth.py): BASELINE shapes, symmetric, loop-free,
deterministic from (shape, seed), planted classes with homophily -- checked on CPU at the small shapes."""
import torch

import rgb_experiment_b200.synth as S


def test_named_shapes_match_baseline_json():
    assert S.SHAPES["cora"] == (2_708, 10_556, 1_433, 7)
    assert S.SHAPES["arxiv"][:3] == (169_343, 2_315_598, 128)
    assert S.SHAPES["reddit"][:3] == (232_965, 114_615_892, 602)
    assert S.SHAPES["products"][:3] == (2_449_029, 123_718_280, 100)


def test_cora_shaped_graph_properties():
    g = S.make_named("cora")
    n, e, f, c = S.SHAPES["cora"]
    ei = g.edge_index
    assert g.num_nodes == n and ei.shape == (2, e) and ei.dtype == torch.int64
    assert g.x.shape == (n, f) and g.x.dtype == torch.float32 and g.y.shape == (n,) and g.num_classes == c
    assert int(ei.min()) >= 0 and int(ei.max()) < n
    assert not bool((ei[0] == ei[1]).any())                                   # no self loops
    fwd = set(zip(ei[0].tolist(), ei[1].tolist()))
    assert all((b, a) in fwd for a, b in fwd)                                 # symmetric
    assert int(g.y.min()) == 0 and int(g.y.max()) == c - 1


def test_same_seed_same_graph_other_seed_other_graph():
    a = S.make_graph(500, 4000, 8, 4, seed=5)
    b = S.make_graph(500, 4000, 8, 4, seed=5)
    c = S.make_graph(500, 4000, 8, 4, seed=6)
    assert torch.equal(a.edge_index, b.edge_index) and torch.equal(a.x, b.x) and torch.equal(a.y, b.y)
    assert not torch.equal(a.edge_index, c.edge_index)


def test_power_law_is_skewed_and_homophily_is_planted():
    g = S.make_graph(5000, 100_000, 8, 5, seed=2)
    u = S.make_graph(5000, 100_000, 8, 5, seed=2, power_law=False)
    deg_g = torch.bincount(g.edge_index[1], minlength=5000)
    deg_u = torch.bincount(u.edge_index[1], minlength=5000)
    assert int(deg_g.max()) > 4 * int(deg_u.max())                            # R-MAT hubs vs Erdos-Renyi-like
    same = (g.y[g.edge_index[0]] == g.y[g.edge_index[1]]).float().mean().item()
    assert same > 0.6                                                         # 0.8 planted + chance, minus collisions


def test_features_carry_the_class_signal():
    g = S.make_graph(3000, 20_000, 16, 4, seed=1)
    means = torch.stack([g.x[g.y == c].mean(0) for c in range(4)])
    # class means differ by much more than the standard error of unit-variance noise
    assert (means[0] - means[1]).abs().max().item() > 0.3


def test_launch_summary_tool_parses_the_committed_launch_list(capsys):
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("launch_summary", os.path.join(root, "tools", "launch_summary.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.summarise(os.path.join(root, "profiles", "r01_launches_v5.csv"))
    out = capsys.readouterr().out
    assert "spmm_rows_kernel" in out and "300 launches" in out
