"""Second, independent statement of the operators whose parity is "unpinned" (no runnable PyG):
each one is written again from its paper's DENSE matrix equations in numpy fp64 -- adjacency
multiplicity matrices, explicit loops over nodes, no scatter / edge lists -- and the edge-list
oracle (oracle/pyg_restated.py) has to agree on random multigraphs with self loops, duplicates
and isolated nodes.  This does not replace golden vectors of torch_geometric itself (absent from
the image); it removes the risk that oracle and CUDA path share one mistake
in the scatter/softmax plumbing.

  GAT       Velickovic et al. 2018, eq. 1-4   (PyG GATConv: loops removed, one loop added per node)
  SuperGAT  Kim & Oh 2021, eq. 2 (MX)         e = a^T[Wh_i || Wh_j] * sigmoid(Wh_i . Wh_j)
  FAGCN     Bo et al. 2021, eq. 3-4           alpha = tanh(g^T[h_i || h_j]), 1/sqrt(d_i d_j)
  GCN/APPNP/SGC  Kipf & Welling 2017; Klicpera et al. 2019 eq. 4; Wu et al. 2019
  C&S       Huang et al. 2020, eq. 4-7        E <- (1-a) E0 + a S E, S = D^-1/2 A D^-1/2
"""
import numpy as np
import pytest
import torch

from helpers import CASES
from oracle import pyg_restated as R

SMALL = ["tiny", "loops_dups", "isolated"]


def counts(ei, n, *, drop_loops=False, add_loop=None):
    """M[i, j] = number of edges j -> i (row = target).  add_loop: 'all' | 'remaining' | None."""
    M = np.zeros((n, n))
    for s, d in zip(ei[0].tolist(), ei[1].tolist()):
        if drop_loops and s == d:
            continue
        M[d, s] += 1
    if add_loop == "all":
        M += np.eye(n)
    elif add_loop == "remaining":                       # add_remaining_self_loops(fill=1): existing loops are REPLACED by one
        np.fill_diagonal(M, 1.0)
    return M


def sym_norm(M):
    deg = M.sum(1)                                      # in-degree (row = target)
    with np.errstate(divide="ignore"):
        dinv = np.where(deg > 0, deg ** -0.5, 0.0)
    return dinv[:, None] * M * dinv[None, :]


def leaky(x, s=0.2):
    return np.where(x > 0, x, s * x)


def masked_softmax_rows(E, M):
    """alpha[i, j] = M_ij exp(E_ij - max_i) / (sum_k M_ik exp(E_ik - max_i) + 1e-16); rows without edges stay 0."""
    A = np.zeros_like(E)
    for i in range(E.shape[0]):
        nz = M[i] > 0
        if not nz.any():
            continue
        m = E[i, nz].max()
        w = M[i] * np.exp(np.where(nz, E[i] - m, -np.inf))
        A[i] = w / (w.sum() + 1e-16)
    return A


@pytest.mark.parametrize("case", SMALL)
def test_gcn_appnp_sgc_dense(case):
    ei, n = CASES[case]()
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n, 5))
    # the existing-loop quirk of add_remaining_self_loops: remove loops, then ONE loop per node
    A_hat = sym_norm(counts(ei, n, drop_loops=True, add_loop="all"))
    xt = torch.from_numpy(X)
    assert np.allclose(R.gcn_propagate(xt, ei).numpy(), A_hat @ X, rtol=0, atol=1e-12)
    Z = X.copy()
    for _ in range(7):
        Z = 0.9 * A_hat @ Z + 0.1 * X
    assert np.allclose(R.appnp_propagate(xt, ei, 7, 0.1).numpy(), Z, rtol=0, atol=1e-12)
    assert np.allclose(R.sgc_propagate(xt, ei, 3).numpy(), np.linalg.matrix_power(A_hat, 3) @ X, rtol=0, atol=1e-12)


@pytest.mark.parametrize("case", SMALL)
def test_sage_mean_dense(case):
    ei, n = CASES[case]()
    X = np.random.default_rng(1).standard_normal((n, 4))
    M = counts(ei, n, drop_loops=True, add_loop="all")          # graphsage.py:55-56
    ref = (M @ X) / np.maximum(M.sum(1), 1.0)[:, None]
    assert np.allclose(R.sage_mean(torch.from_numpy(X), ei).numpy(), ref, rtol=0, atol=1e-12)


@pytest.mark.parametrize("case", SMALL)
@pytest.mark.parametrize("H,C", [(1, 5), (3, 4)])
def test_gat_dense(case, H, C):
    ei, n = CASES[case]()
    rng = np.random.default_rng(2)
    Xp = rng.standard_normal((n, H, C))
    att_s, att_d = rng.standard_normal((H, C)), rng.standard_normal((H, C))
    a_s, a_d = (Xp * att_s).sum(-1), (Xp * att_d).sum(-1)       # a^T [W h_i || W h_j] split into two halves
    M = counts(ei, n, drop_loops=True, add_loop="all")
    out = np.zeros((n, H, C))
    for h in range(H):
        E = leaky(a_d[:, h][:, None] + a_s[:, h][None, :])     # E[i, j] for edge j -> i
        out[:, h, :] = masked_softmax_rows(E, M) @ Xp[:, h, :]
    got, alpha, ed = R.gat_aggregate(torch.from_numpy(Xp), torch.from_numpy(a_s), torch.from_numpy(a_d), ei, 0.2)
    assert np.allclose(got.numpy(), out, rtol=0, atol=1e-12)
    # every node has its loop, so every softmax row sums to 1 (up to the 1e-16 guard)
    assert np.allclose(R.scatter_add(alpha, ed[1], 0, dim_size=n).numpy(), 1.0, atol=1e-12)


@pytest.mark.parametrize("case", SMALL)
def test_supergat_mx_dense(case):
    ei, n = CASES[case]()
    H, C = 2, 3
    rng = np.random.default_rng(3)
    Xp = rng.standard_normal((n, H, C))
    att_l, att_r = rng.standard_normal((1, H, C)), rng.standard_normal((1, H, C))
    M = counts(ei, n, drop_loops=True, add_loop="all")
    out = np.zeros((n, H, C))
    for h in range(H):
        x = Xp[:, h, :]
        l, r = x @ att_l[0, h], x @ att_r[0, h]                # att_l acts on the source x_j, att_r on the target x_i
        E = leaky((r[:, None] + l[None, :]) / (1.0 + np.exp(-(x @ x.T))))
        out[:, h, :] = masked_softmax_rows(E, M) @ x
    ed = R.edit_loops(ei, n, R.LOOP_REMOVE_THEN_ADD)
    xt = torch.from_numpy(Xp)
    a, _ = R.supergat_mx_alpha(xt, torch.from_numpy(att_l), torch.from_numpy(att_r), ed, 0.2)
    alpha = R.softmax(a, ed[1], num_nodes=n)
    got = R.scatter_add(xt[ed[0]] * alpha.unsqueeze(-1), ed[1], dim=0, dim_size=n)
    assert np.allclose(got.numpy(), out, rtol=0, atol=1e-12)


@pytest.mark.parametrize("case", SMALL)
def test_faconv_dense(case):
    ei, n = CASES[case]()
    rng = np.random.default_rng(4)
    X, X0 = rng.standard_normal((n, 6)), rng.standard_normal((n, 6))
    gl, gr = rng.standard_normal(6), rng.standard_normal(6)
    a_l, a_r = X @ gl, X @ gr
    S = sym_norm(counts(ei, n, drop_loops=True, add_loop="all"))
    T = np.tanh(a_r[:, None] + a_l[None, :])                    # target half + source half
    ref = (T * S) @ X + 0.3 * X0
    got = R.faconv_aggregate(torch.from_numpy(X), torch.from_numpy(X0), torch.from_numpy(a_l), torch.from_numpy(a_r),
                             ei, 0.3)
    assert np.allclose(got.numpy(), ref, rtol=0, atol=1e-12)


@pytest.mark.parametrize("case", SMALL)
@pytest.mark.parametrize("autoscale", [True, False])
def test_correct_and_smooth_dense(case, autoscale):
    ei, n = CASES[case]()
    Cn = 4
    rng = np.random.default_rng(5)
    P = rng.random((n, Cn)) + 0.1
    P /= P.sum(1, keepdims=True)
    y = rng.integers(0, Cn, n)
    train = np.arange(n) % 3 == 0
    Y = np.eye(Cn)[y]
    S = sym_norm(counts(ei, n))                                 # gcn_norm(add_self_loops=False): loops and duplicates stay
    # correct (eq. 4-6): E0 = (Y - P) on the training rows; 6 steps; clamp(-1, 1) or training rows re-fixed
    E0 = np.zeros((n, Cn))
    E0[train] = Y[train] - P[train]
    E = E0.copy()
    a1 = 0.8
    for _ in range(6):
        E = a1 * (S @ E) + (1 - a1) * E0
        if autoscale:
            E = np.clip(E, -1.0, 1.0)
        else:
            E[train] = E0[train]
    if autoscale:
        sigma = np.abs(E0[train]).sum() / train.sum()
        with np.errstate(divide="ignore"):
            sc = sigma / np.abs(E).sum(1, keepdims=True)
        sc[np.isinf(sc) | (sc > 1000)] = 1.0
        Z = P + sc * E
    else:
        Z = P + 1.5 * E
    mask = torch.from_numpy(train)
    yt = torch.from_numpy(y[train])                             # the caller passes the TRAINING labels (itexperiments.py:525)
    got = R.cs_correct(torch.from_numpy(P), yt, mask, ei, 6, a1, autoscale, 1.5)
    assert np.allclose(got.numpy(), Z, rtol=0, atol=1e-12)
    # smooth (eq. 7): G0 = Z with the true labels on the training rows; clamp(0, 1)
    G0 = Z.copy()
    G0[train] = Y[train]
    G = G0.copy()
    for _ in range(5):
        G = np.clip(0.7 * (S @ G) + 0.3 * G0, 0.0, 1.0)
    got2 = R.cs_smooth(torch.from_numpy(Z), yt, mask, ei, 5, 0.7)
    assert np.allclose(got2.numpy(), G, rtol=0, atol=1e-12)


@pytest.mark.parametrize("case", SMALL + ["empty", "single_node"])
def test_to_undirected_and_coalesce_as_sets(case):
    ei, n = CASES[case]()
    pairs = set(zip(ei[0].tolist(), ei[1].tolist()))
    und = sorted(pairs | {(b, a) for a, b in pairs})
    got = R.to_undirected(ei, n)
    assert list(zip(got[0].tolist(), got[1].tolist())) == und
    idx, _ = R.coalesce(ei, None, n, n)
    assert list(zip(idx[0].tolist(), idx[1].tolist())) == sorted(pairs)
