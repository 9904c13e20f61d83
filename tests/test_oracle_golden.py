"""Pin the oracle against outputs of the reference's own in-tree functions (tests/golden/*.npz,
made by tests/golden/make_golden.py) and against closed forms."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import pyg_restated as R

FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def load(path):
    z = np.load(path)
    return {k: (torch.from_numpy(z[k]) if z[k].ndim else z[k].item()) for k in z.files}


def test_golden_present():
    assert len(FILES) >= 4


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p) for p in FILES])
def test_gcn_norm_matches_reference_intree_copy(path):
    g = load(path)
    ei, n = g["edge_index"], g["num_nodes"]
    ei2, w = R.gcn_norm(ei, None, n, dtype=torch.float32)
    assert torch.equal(ei2, g["gcn_norm_edge_index"])          # integer work: bit-exact
    assert torch.equal(w, g["gcn_norm_weight"])
    _, w3 = R.gcn_norm(ei, None, n, add_self_loops=False, dtype=torch.float32)
    assert torch.equal(w3, g["gcn_norm_noloop_weight"])
    assert torch.equal(R.edit_loops(ei, n, R.LOOP_ADD_REMAINING), g["gcn_norm_edge_index"])


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p) for p in FILES])
def test_pta_ops_match_reference(path):
    g = load(path)
    ei, n = g["edge_index"], g["num_nodes"]
    r, c, v = R.pta_norm_adj(ei, n, torch.float64)
    dense = torch.zeros(n, n, dtype=torch.float64).index_put_((r, c), v, accumulate=True)
    assert torch.allclose(dense, g["pta_adj_dense"], rtol=0, atol=1e-15)
    y = R.pta_label_propagation(ei, n, g["pta_labels"], g["pta_idx"], g["pta_K"], g["pta_alpha"])
    assert (y - g["pta_lp"]).abs().max() <= 1e-6 * g["pta_lp"].abs().max()
    out = R.pta_inference(g["pta_h"], ei, n, g["pta_K"], g["pta_alpha"])
    assert (out - g["pta_inference"]).abs().max() <= 1e-6 * g["pta_inference"].abs().max()


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p) for p in FILES])
def test_dagnn_prop_matches_reference(path):
    g = load(path)
    ei = g["edge_index"]
    pps = R.dagnn_hops(g["prop_x"], ei, g["prop_K"])
    score = torch.sigmoid(pps @ g["prop_proj_w"].t() + g["prop_proj_b"]).squeeze(-1).unsqueeze(1)
    out = torch.matmul(score, pps).squeeze(1)
    assert torch.allclose(out, g["prop_out"], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p) for p in FILES])
def test_sage_mean_matches_reference(path):
    g = load(path)
    x = g["sage_x"]
    xl = x @ g["sage_wl"].t() + g["sage_bl"]
    xr = x @ g["sage_wr"].t() + g["sage_br"]
    out = R.sage_mean(xl, g["edge_index"]) + xr
    assert torch.allclose(out, g["sage_out"], rtol=1e-6, atol=1e-6)


def test_pta_equals_reversed_gcn_when_no_loops():
    """normalize_adj operator == gcn_norm+propagate on edge_index.flip(0)."""
    g = load(FILES[0])
    ei, n = g["edge_index"], g["num_nodes"]
    x = torch.randn(n, 5, dtype=torch.float64)
    r, c, v = R.pta_norm_adj(ei, n, torch.float64)
    a = R.pta_spmm(r, c, v, x)
    b = R.gcn_propagate(x, ei.flip(0))
    assert torch.allclose(a, b, atol=1e-12)


# ---------------- closed forms (K2) ----------------

def complete_graph(n):
    i, j = torch.meshgrid(torch.arange(n), torch.arange(n), indexing="ij")
    m = i != j
    return torch.stack([i[m], j[m]])


def test_gcn_on_complete_graph_weights_are_1_over_n():
    n = 9
    ei, w = R.gcn_norm(complete_graph(n), None, n, dtype=torch.float64)
    assert torch.allclose(w, torch.full_like(w, 1.0 / n))


def test_mean_of_constant_is_constant_and_isolated_rows_are_zero():
    ei = torch.tensor([[0, 1, 2], [1, 2, 0]])
    x = torch.full((5, 3), 2.5)
    out = R.propagate(ei, x, None, "mean", 5)
    assert torch.equal(out[:3], x[:3]) and torch.equal(out[3:], torch.zeros(2, 3))


def test_appnp_converges_to_ppr_fixed_point():
    torch.manual_seed(0)
    n = 30
    ei = R.to_undirected(torch.randint(n, (2, 120)), n)
    z0 = torch.randn(n, 4, dtype=torch.float64)
    alpha = 0.2
    ei2, w = R.gcn_norm(ei, None, n, dtype=torch.float64)
    A = torch.zeros(n, n, dtype=torch.float64).index_put_((ei2[1], ei2[0]), w, accumulate=True)
    exact = alpha * torch.linalg.solve(torch.eye(n, dtype=torch.float64) - (1 - alpha) * A, z0)
    z = R.appnp_propagate(z0, ei, 200, alpha)
    assert torch.allclose(z, exact, atol=1e-10)


def test_softmax_rows_sum_to_one():
    torch.manual_seed(1)
    idx = torch.randint(7, (50,))
    s = R.softmax(torch.randn(50, 3, dtype=torch.float64), idx, num_nodes=9)
    tot = R.scatter(s, idx, 0, 9, "sum")
    present = torch.bincount(idx, minlength=9) > 0
    assert torch.allclose(tot[present], torch.ones_like(tot[present]), atol=1e-12)
    assert torch.equal(tot[~present], torch.zeros_like(tot[~present]))


def test_cs_smooth_alpha0_returns_clamped_input():
    torch.manual_seed(2)
    n, C = 12, 3
    ei = R.to_undirected(torch.randint(n, (2, 30)), n)
    y = torch.softmax(torch.randn(n, C), -1)
    mask = torch.zeros(n, dtype=torch.bool)
    mask[:4] = True
    yt = torch.randint(C, (4,))
    out = R.cs_smooth(y, yt, mask, ei, 5, 0.0)
    exp = y.clone()
    exp[mask] = torch.nn.functional.one_hot(yt, C).float()
    assert torch.allclose(out, exp.clamp(0, 1))


def test_csr_build_is_stable_and_consistent():
    torch.manual_seed(3)
    n = 17
    ei = torch.randint(n, (2, 200))
    ed = R.edit_loops(ei, n, R.LOOP_REMOVE_THEN_ADD)
    rowptr, col, eid = R.csr_build(ed, n, "dst")
    assert rowptr[-1] == ed.size(1)
    for i in range(n):
        seg = eid[rowptr[i]:rowptr[i + 1]]
        assert torch.all(ed[1, seg] == i)
        assert torch.all(seg[1:] > seg[:-1])            # stable = increasing edge id
        assert torch.equal(col[rowptr[i]:rowptr[i + 1]], ed[0, seg])
    rp_t, col_t, eid_t = R.csr_build(ed, n, "src")
    assert torch.equal(torch.sort(eid_t)[0], torch.arange(ed.size(1)))


def test_gat_backward_identities_fp64():
    """SURVEY A10 identities used by the fused CUDA backward, checked against autograd."""
    torch.manual_seed(4)
    n, H, C = 11, 2, 3
    ei = torch.randint(n, (2, 40))
    xp = torch.randn(n, H, C, dtype=torch.float64, requires_grad=True)
    a_s = torch.randn(n, H, dtype=torch.float64, requires_grad=True)
    a_d = torch.randn(n, H, dtype=torch.float64, requires_grad=True)
    out, alpha, ed = R.gat_aggregate(xp, a_s, a_d, ei, 0.2)
    dout = torch.randn_like(out)
    gx, gs, gd = torch.autograd.grad(out, (xp, a_s, a_d), dout)
    row, col = ed[0], ed[1]
    with torch.no_grad():
        dalpha = (dout[col] * xp[row]).sum(-1)
        S = (dout * out).sum(-1)
        de = alpha * (dalpha - S[col])
        raw = a_s[row] + a_d[col]
        dlogit = de * torch.where(raw > 0, torch.ones_like(raw), torch.full_like(raw, 0.2))
        gs2 = torch.zeros_like(a_s).index_add_(0, row, dlogit)
        gd2 = torch.zeros_like(a_d).index_add_(0, col, dlogit)
        gx2 = torch.zeros_like(xp).index_add_(0, row, alpha.unsqueeze(-1) * dout[col])
    assert torch.allclose(gx, gx2, atol=1e-12)
    assert torch.allclose(gs, gs2, atol=1e-12)
    assert torch.allclose(gd, gd2, atol=1e-12)
