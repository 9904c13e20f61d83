"""One-cluster shared-memory K-hop path (csrc/khop_cta.cu) for small graphs: same operator as the K-launch path.
Rows that neither path splits accumulate their edges in the same order -> bit-identical; longer rows are split
differently (a warp's 32/G edge slots here, CTA-sized items there) and are held to 5e-6 of the K-launch result.
The oracle comparisons of tests/test_gpu_khop.py run through this path too (their graphs are small); here the two
CUDA paths are compared with each other, family by family, incl. the epilogue variants."""
import pytest
import torch

from helpers import CASES, relerr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NO_CTA = 1 << 30          # RGBMP_TUNE_NO_CTA


def P():
    import rgb_experiment_b200 as P_
    return P_


def families(p, ei, n, F, seed):
    g = p.Graph(ei, n, 2)
    gn = p.Graph(ei, n, 0)
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(n, F, generator=gen).to(DEV)
    y0 = torch.rand(n, F, generator=gen).to(DEV)
    mask = (torch.rand(n, generator=gen) < 0.3).to(DEV)
    out = {
        "appnp": p.ops.appnp(x, g, 10, 0.1),
        "appnp_edge_weights": p.ops.appnp(x, g, 3, 0.2, fold=False),
        "sgc": p.ops.gcn_power(x, g, 2),
        "sgc_edge_weights": p.ops.gcn_power(x, g, 2, fold=False),
        "lp_clamp": p.ops.label_propagation(gn, y0, 50, 0.8),
        "lp_reset": p.ops.label_propagation(gn, y0, 7, 0.8, reset_mask=mask, reset_val=y0),
        "lp_noclamp": p.ops.label_propagation(gn, x, 5, 0.9, clamp=None),
        "dagnn_hops": p.ops.dagnn_hops(x, g, 4),
    }
    xg = x.clone().requires_grad_(True)
    z = p.ops.appnp(xg, g, 4, 0.15)
    z.square().sum().backward()
    out["appnp_grad"] = xg.grad
    return out


@pytest.mark.parametrize("case", ["tiny", "loops_dups", "isolated", "hub", "empty", "single_node", "medium", "cora"])
@pytest.mark.parametrize("F", [1, 3, 7, 10, 16, 47])
def test_one_cta_khop_equals_k_launch_path(case, F, monkeypatch):
    p = P()
    import rgb_experiment_b200.memo as memo
    monkeypatch.setattr(memo, "MIN_WORK", float("inf"))
    L = p._lib.lib()
    if case == "cora":
        import rgb_experiment_b200.synth as S
        sg = S.make_named("cora", features=False)
        ei, n = sg.edge_index, sg.num_nodes
    else:
        ei, n = CASES[case]()
    ei = ei.to(DEV)
    before = L.rgbmp_khop_cta_calls()
    fast = families(p, ei, n, F, 3)
    took = L.rgbmp_khop_cta_calls() - before
    ld = (F + 3) // 4 * 4

    def eligible(nnz):          # csrc/khop_cta.cu: khop_cta_try
        return ld <= 16 and nnz * ld * 4 <= 8e6 and 1 <= n <= 200000

    nnz_hi = p.Graph(ei, n, 2).nnz
    nnz_lo = p.Graph(ei, n, 0).nnz
    if eligible(nnz_hi):
        assert took >= 9, took           # every family of families() (all K >= 2) went through the cluster kernel
    if not eligible(nnz_lo):
        assert took == 0, took
    monkeypatch.setattr(p.ops, "TUNE_OVERRIDE", NO_CTA)
    before = L.rgbmp_khop_cta_calls()
    slow = families(p, ei, n, F, 3)
    assert L.rgbmp_khop_cta_calls() == before
    deg_max = int(torch.bincount(ei[1], minlength=n).max()) + 1 if ei.numel() else 1
    for k in fast:
        a, b = fast[k], slow[k]
        assert a.shape == b.shape
        if deg_max <= 8:                     # no row leaves its lane group in either path
            assert torch.equal(a, b), k
        else:
            assert relerr(a, b) <= 5e-6, (k, relerr(a, b))      # both sit within 1e-5 of the fp64 oracle (test_gpu_khop.py)


def test_switch_and_environment_knob():
    p = P()
    L = p._lib.lib()
    old = L.rgbmp_set_khop_cta(0)
    try:
        ei, n = CASES["loops_dups"]()
        g = p.Graph(ei.to(DEV), n, 2)
        before = L.rgbmp_khop_cta_calls()
        p.ops.appnp(torch.randn(n, 7, device=DEV), g, 3, 0.1)
        assert L.rgbmp_khop_cta_calls() == before
        assert L.rgbmp_set_khop_cta(1) == 0
        p.ops.appnp(torch.randn(n, 7, device=DEV), g, 3, 0.1)
        assert L.rgbmp_khop_cta_calls() == before + 1
        assert L.rgbmp_set_khop_cta(-1) == 1          # query only
    finally:
        L.rgbmp_set_khop_cta(old)
