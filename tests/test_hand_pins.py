"""The CPU oracle against hand-computed known answers (tests/hand_vectors.py): pins for the operators whose
reference implementation (PyG 1.7-2.0) cannot be run here."""
import torch

import hand_vectors as H
from oracle import layers as OL
from oracle import pyg_restated as R

TOL = 1e-6


def close(a, b, tol=TOL):
    return float((a.double() - b.double()).abs().max()) <= tol


def test_add_remaining_self_loops_with_existing_loops():
    ei, w = R.add_remaining_self_loops(H.ARSL_EI, H.ARSL_W.float(), 1.0, 3)
    assert torch.equal(ei, H.ARSL_OUT_EI) and close(w, H.ARSL_OUT_W.float(), 0)


def test_gcn_norm_propagate_appnp_mean():
    ei, w = R.gcn_norm(H.G_EI, None, H.G_N, False, True, dtype=torch.float64)
    assert torch.equal(ei, H.G_NORM_EI) and close(w, H.G_NORM_W)
    assert close(R.gcn_propagate(H.G_X.double(), H.G_EI), H.G_PROP)
    assert close(R.appnp_propagate(H.G_X.double(), H.G_EI, 1, 0.1), H.G_APPNP1)
    assert close(R.sage_mean(H.G_X.double(), H.G_EI), H.G_MEAN)


def test_gat_edge_softmax():
    xp = H.G_X.double().view(3, 1, 1)
    out, alpha, ei = R.gat_aggregate(xp, H.GAT_AS.double(), H.GAT_AD.double(), H.G_EI, 0.2)
    assert close(out.view(3, 1), H.GAT_OUT, 1e-12)
    # every target's attention sums to one; target 1's three weights are 1/7, 2/7, 4/7
    t1 = sorted(float(a) for a, c in zip(alpha.view(-1), ei[1]) if int(c) == 1)
    assert all(abs(a - b) < 1e-12 for a, b in zip(t1, [1 / 7, 2 / 7, 4 / 7]))
    out, _, _ = R.gat_aggregate(xp, H.GAT_AS.double(), H.GAT_AD_NEG.double(), H.G_EI, 0.2)
    assert close(out.view(3, 1), H.GAT_OUT_NEG, 1e-12)


def test_correct_and_smooth():
    auto = OL.CorrectAndSmooth(1, 0.5, 1, 0.5, autoscale=True)
    assert close(auto.correct(H.CS_YSOFT.clone(), H.CS_YTRUE, H.CS_MASK, H.CS_EI), H.CS_CORRECT_AUTO)
    fixed = OL.CorrectAndSmooth(1, 0.5, 1, 0.5, autoscale=False, scale=1.0)
    assert close(fixed.correct(H.CS_YSOFT.clone(), H.CS_YTRUE, H.CS_MASK, H.CS_EI), H.CS_CORRECT_FIXED)
    assert close(auto.smooth(H.CS_YSOFT.clone(), H.CS_YTRUE, H.CS_MASK, H.CS_EI), H.CS_SMOOTH)


def test_coalesce_and_to_undirected():
    assert torch.equal(R.coalesce(H.CO_EI, None, 3, 3)[0], H.CO_OUT)
    assert torch.equal(R.to_undirected(H.UND_EI, 3), H.UND_OUT)
