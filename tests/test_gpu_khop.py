"""Fused K-hop propagation families (APPNP, SGC, DAGNN hops, LabelPropagation / C&S, PTA) against
the oracle and against the reference's own in-tree functions (tests/golden).  Through the C ABI."""
import pytest
import torch

from helpers import CASES, GOLDEN, load_golden, relerr
from oracle import pyg_restated as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5


def P():
    import rgb_experiment_b200 as P_
    return P_


@pytest.mark.parametrize("case", ["loops_dups", "isolated", "hub", "medium"])
@pytest.mark.parametrize("fold", [False, True])
@pytest.mark.parametrize("K,F", [(1, 7), (2, 47), (10, 47), (5, 64)])
def test_appnp_forward_backward(case, fold, K, F):
    p = P()
    ei, n = CASES[case]()
    gen = torch.Generator().manual_seed(K * 100 + F)
    x = torch.randn(n, F, generator=gen)
    dz = torch.randn(n, F, generator=gen)
    xo = x.double().requires_grad_(True)                 # fp64 arbiter
    zo = R.appnp_propagate(xo, ei, K, 0.1)
    zo.backward(dz.double())
    g = p.Graph(ei.to(DEV), n, p.LOOP_ADD_REMAINING)
    xg = x.to(DEV).requires_grad_(True)
    zg = p.ops.appnp(xg, g, K, 0.1, fold)
    zg.backward(dz.to(DEV))
    assert relerr(zg.detach(), zo.detach()) <= TOL
    assert relerr(xg.grad, xo.grad) <= TOL
    if not fold:
        z32 = R.appnp_propagate(x, ei, K, 0.1)
        assert relerr(zg.detach(), z32) <= TOL


def test_appnp_k_to_infinity_is_the_ppr_fixed_point():
    p = P()
    ei, n = CASES["loops_dups"]()
    ei = R.to_undirected(ei, n)
    z0 = torch.randn(n, 6, generator=torch.Generator().manual_seed(0))
    alpha = 0.2
    ei2, w = R.gcn_norm(ei, None, n, dtype=torch.float64)
    A = torch.zeros(n, n, dtype=torch.float64).index_put_((ei2[1], ei2[0]), w, accumulate=True)
    exact = alpha * torch.linalg.solve(torch.eye(n, dtype=torch.float64) - (1 - alpha) * A, z0.double())
    g = p.Graph(ei.to(DEV), n, p.LOOP_ADD_REMAINING)
    z = p.ops.appnp(z0.to(DEV), g, 150, alpha)
    assert relerr(z, exact) <= TOL


@pytest.mark.parametrize("case", ["loops_dups", "hub"])
def test_sgc_power_and_dagnn_hops(case):
    p = P()
    ei, n = CASES[case]()
    x = torch.randn(n, 100, generator=torch.Generator().manual_seed(1))
    g = p.Graph(ei.to(DEV), n, p.LOOP_ADD_REMAINING)
    xg = x.to(DEV).requires_grad_(True)
    y = p.ops.gcn_power(xg, g, 2)
    xo = x.double().requires_grad_(True)
    yo = R.sgc_propagate(xo, ei, 2)
    assert relerr(y.detach(), yo.detach()) <= TOL
    dy = torch.randn(n, 100, generator=torch.Generator().manual_seed(2))
    y.backward(dy.to(DEV))
    yo.backward(dy.double())
    assert relerr(xg.grad, xo.grad) <= TOL
    xs = x[:, :9].contiguous()
    hops = p.ops.dagnn_hops(xs.to(DEV), g, 4)
    assert hops.shape == (n, 5, 9)
    assert relerr(hops, R.dagnn_hops(xs.double(), ei, 4)) <= TOL


@pytest.mark.parametrize("path", GOLDEN)
def test_pta_ops_match_reference_golden(path):
    """itexperiments.py:671-719 and pta.py:79-84 executed verbatim in the build container."""
    p = P()
    gold = load_golden(path)
    ei, n = gold["edge_index"], gold["num_nodes"]
    g = p.ops.pta_graph(ei.to(DEV), n)
    y = p.ops.pta_label_propagation(g, gold["pta_labels"].to(DEV), gold["pta_idx"].to(DEV), gold["pta_K"],
                                    gold["pta_alpha"])
    assert relerr(y, gold["pta_lp"]) <= TOL
    out = p.ops.pta_inference(gold["pta_h"].to(DEV), g, gold["pta_K"], gold["pta_alpha"])
    assert relerr(out, gold["pta_inference"]) <= TOL


@pytest.mark.parametrize("path", GOLDEN)
def test_dagnn_prop_and_sage_match_reference_golden(path):
    """dagnn.py:34-65 and graphsage.py:36-62 outputs recorded from the reference."""
    p = P()
    gold = load_golden(path)
    ei, n = gold["edge_index"].to(DEV), gold["num_nodes"]
    g = p.Graph(ei, n, p.LOOP_ADD_REMAINING)
    pps = p.ops.dagnn_hops(gold["prop_x"].to(DEV), g, gold["prop_K"])
    score = torch.sigmoid(pps @ gold["prop_proj_w"].to(DEV).t() + gold["prop_proj_b"].to(DEV))
    out = torch.matmul(score.squeeze(-1).unsqueeze(1), pps).squeeze(1)
    assert relerr(out, gold["prop_out"]) <= TOL
    x = gold["sage_x"].to(DEV)
    xl = x @ gold["sage_wl"].to(DEV).t() + gold["sage_bl"].to(DEV)
    xr = x @ gold["sage_wr"].to(DEV).t() + gold["sage_br"].to(DEV)
    gs = p.Graph(ei, n, p.LOOP_REMOVE_THEN_ADD)
    out = p.ops.propagate(xl, gs, "mean") + xr
    assert relerr(out, gold["sage_out"]) <= TOL


@pytest.mark.parametrize("case", ["loops_dups", "isolated", "hub"])
@pytest.mark.parametrize("autoscale", [True, False])
def test_correct_and_smooth(case, autoscale):
    p = P()
    from rgb_experiment_b200.shim import nn as PL
    from oracle import layers as OL
    ei, n = CASES[case]()
    ei = R.to_undirected(ei, n)
    C = 5
    gen = torch.Generator().manual_seed(3)
    y_soft = torch.softmax(torch.randn(n, C, generator=gen), -1)
    y = torch.randint(C, (n,), generator=gen)
    mask = torch.rand(n, generator=gen) < 0.4
    kw = dict(num_correction_layers=50, correction_alpha=0.8, num_smoothing_layers=50, smoothing_alpha=0.8,
              autoscale=autoscale, scale=0.7)
    o, g = OL.CorrectAndSmooth(**kw), PL.CorrectAndSmooth(**kw)
    c_o = o.correct(y_soft.clone(), y[mask], mask, ei)
    c_g = g.correct(y_soft.to(DEV), y[mask].to(DEV), mask.to(DEV), ei.to(DEV))
    assert relerr(c_g, c_o) <= 5e-5            # 50 hops, then a division by a small L1 norm (autoscale)
    s_o = o.smooth(c_o, y[mask], mask, ei)
    s_g = g.smooth(c_o.to(DEV), y[mask].to(DEV), mask.to(DEV), ei.to(DEV))
    assert relerr(s_g, s_o) <= TOL
    assert float(s_g.min()) >= 0.0 and float(s_g.max()) <= 1.0


def test_cs_smooth_with_alpha_zero_returns_the_clamped_input():
    p = P()
    from rgb_experiment_b200.shim import nn as PL
    ei, n = CASES["loops_dups"]()
    C = 4
    gen = torch.Generator().manual_seed(4)
    y_soft = torch.softmax(torch.randn(n, C, generator=gen), -1)
    y = torch.randint(C, (n,), generator=gen)
    mask = torch.zeros(n, dtype=torch.bool)
    mask[:10] = True
    cs = PL.CorrectAndSmooth(3, 0.5, 5, 0.0)
    out = cs.smooth(y_soft.to(DEV), y[mask].to(DEV), mask.to(DEV), ei.to(DEV))
    exp = y_soft.clone()
    exp[mask] = torch.nn.functional.one_hot(y[mask], C).float()
    assert torch.equal(out.cpu(), exp.clamp(0, 1))


def test_host_buffer_entry_point_matches_device_path():
    """rgbmp_appnp_host: H2D -> K hops -> D2H (the e2e form bench.py times)."""
    p = P()
    ei, n = CASES["medium"]()
    F, K, alpha = 47, 10, 0.1
    g = p.Graph(ei.to(DEV), n, p.LOOP_ADD_REMAINING)
    z0 = torch.randn(n, F, generator=torch.Generator().manual_seed(5)).pin_memory()
    out = torch.empty(n, F).pin_memory()
    p.ops.appnp_host(g, z0, out, K, alpha)
    ref = R.appnp_propagate(z0.double(), ei, K, alpha)
    assert relerr(out, ref) <= TOL



@pytest.mark.parametrize("case", ["loops_dups", "hub"])
def test_dagnn_hops_backward(case):
    """ops.dagnn_hops keeps every hop in one K-hop call and is differentiable (backward = K transposed SpMMs)."""
    p = P()
    ei, n = CASES[case]()
    g = p.Graph(ei.to(DEV), n, p.LOOP_ADD_REMAINING)
    x = torch.randn(n, 5, generator=torch.Generator().manual_seed(2))
    dS = torch.randn(n, 4, 5, generator=torch.Generator().manual_seed(3))
    xo = x.double().requires_grad_(True)
    R.dagnn_hops(xo, ei, 3).backward(dS.double())
    xg = x.to(DEV).requires_grad_(True)
    out = p.ops.dagnn_hops(xg, g, 3)
    out.backward(dS.to(DEV))
    assert relerr(out.detach(), R.dagnn_hops(x.double(), ei, 3)) <= TOL
    assert relerr(xg.grad, xo.grad) <= TOL
