"""Size-independent properties at BASELINE.json's FULL sizes (the oracle cannot run there in seconds):
products-shaped K-hop propagation (C4: 2.45 M nodes / 126 M edges) and the Reddit-shaped GAT layer
(C3: 233 K nodes / 115 M edges).  Through the C ABI, like every GPU parity test."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def relmax(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def products():
    import rgb_experiment_b200 as P
    import rgb_experiment_b200.synth as S
    sg = S.make_named("products", device=DEV, features=False)
    g = P.Graph(sg.edge_index, sg.num_nodes, P.LOOP_ADD_REMAINING)
    yield P, g, sg.num_nodes
    del g, sg
    torch.cuda.empty_cache()


def test_sqrt_degree_is_a_fixed_point_of_appnp_at_products_size(products):
    """A_hat = D^-1/2 (A+I) D^-1/2 has the eigenvector sqrt(deg) with eigenvalue 1 on a symmetric graph, so
    z0 = sqrt(deg) * c is reproduced by every hop and therefore by APPNP(K, alpha) for any K, alpha --
    in both normalisation forms (per-edge weights / folded row scalings), forward and (transpose graph) backward."""
    P, g, n = products
    deg = g.fwd.degree().to(torch.float32)
    c = torch.linspace(-2.0, 3.0, 47, device=DEV)
    z0 = deg.sqrt().unsqueeze(1) * c
    for fold in (False, True):
        z = P.ops.appnp(z0, g, 10, 0.1, fold)
        assert relmax(z, z0) <= 1e-5, fold
    zt = P.ops._appnp_khop(g.bwd, g, z0, 10, 0.1, True, False)      # the backward operator (A_hat is symmetric here)
    assert relmax(zt, z0) <= 1e-5
    # SGC power: A_hat^2 z0 = z0 ; C&S-style LP without clamp effect: values stay inside [0, 1] after scaling
    assert relmax(P.ops.gcn_power(z0, g, 2), z0) <= 1e-5


def test_appnp_is_linear_and_self_adjoint_at_products_size(products):
    P, g, n = products
    gen = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn(n, 47, device=DEV, generator=gen)
    y = torch.randn(n, 47, device=DEV, generator=gen)
    M = lambda t: P.ops.appnp(t, g, 10, 0.1, True)
    assert relmax(M(2.0 * x - 0.5 * y), 2.0 * M(x) - 0.5 * M(y)) <= 1e-5
    # <M x, y> = <x, M^T y>, M^T = the same recursion on the transpose CSR (what autograd's backward runs)
    xr = x.clone().requires_grad_(True)
    (P.ops.appnp(xr, g, 10, 0.1, True) * y).sum().backward()
    Mty = P.ops._appnp_khop(g.bwd, g, y, 10, 0.1, True, True)
    assert relmax(xr.grad, Mty) <= 1e-5
    d1 = (M(x).double() * y.double()).sum().item()
    d2 = (x.double() * Mty.double()).sum().item()
    assert abs(d1 - d2) <= 1e-6 * max(abs(d1), 1.0)


def test_label_propagation_keeps_a_distribution_at_products_size(products):
    """C&S smoothing on the loop-free graph: entries stay in [0, 1] (the clamp) and a one-hot start with alpha = 0
    is returned unchanged."""
    P, _, n = products
    import rgb_experiment_b200.synth as S
    sg = S.make_named("products", device=DEV, features=False)
    g0 = P.Graph(sg.edge_index, n, P.LOOP_NONE)
    y = torch.nn.functional.one_hot(torch.arange(n, device=DEV) % 47, 47).float()
    out = P.ops.label_propagation(g0, y, 5, 0.8)
    assert float(out.min()) >= 0.0 and float(out.max()) <= 1.0 and bool(torch.isfinite(out).all())
    assert torch.equal(P.ops.label_propagation(g0, y, 3, 0.0), y)


def test_gat_with_constant_logits_is_the_mean_at_reddit_size():
    """a_src = a_dst = 0 makes every alpha_ij = 1 / deg_i: the fused edge-softmax + aggregate must equal the mean
    aggregation kernel on the same (self-loop-edited) graph; with Xp = 1 the rows sum to one; the backward w.r.t.
    Xp is then the transpose mean."""
    import rgb_experiment_b200 as P
    import rgb_experiment_b200.synth as S
    sg = S.make_named("reddit", device=DEV, features=False)
    n = sg.num_nodes
    g = P.Graph(sg.edge_index, n, P.LOOP_REMOVE_THEN_ADD)
    assert g.nnz == sg.edge_index.size(1) + n
    gen = torch.Generator(device=DEV).manual_seed(9)
    for H, C in ((8, 8), (1, 41)):
        xp = torch.randn(n, H * C, device=DEV, generator=gen).requires_grad_(True)
        zeros = torch.zeros(n, H, device=DEV)
        out = P.ops.gat(xp, zeros, zeros, g, H, C, 0.2)
        ref = P.ops.propagate(xp.detach(), g, "mean")
        assert relmax(out.detach(), ref) <= 1e-5, (H, C)
        w = torch.randn(n, H * C, device=DEV, generator=gen)
        (out * w).sum().backward()
        xm = xp.detach().clone().requires_grad_(True)
        (P.ops.propagate(xm, g, "mean") * w).sum().backward()
        assert relmax(xp.grad, xm.grad) <= 1e-5, (H, C)
    a_s = torch.randn(n, 8, device=DEV, generator=gen) * 3
    a_d = torch.randn(n, 8, device=DEV, generator=gen) * 3
    ones = P.ops.gat(torch.ones(n, 64, device=DEV), a_s, a_d, g, 8, 8, 0.2)
    assert float((ones - 1).abs().max()) <= 1e-5
