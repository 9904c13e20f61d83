"""The CUDA path against the same hand-computed known answers (tests/hand_vectors.py) -- directly, not through
the oracle."""
import pytest
import torch

import hand_vectors as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-6


def close(a, b, tol=TOL):
    return float((a.double().cpu() - b.double()).abs().max()) <= tol


def test_loop_edit_and_gcn_norm_weights():
    import rgb_experiment_b200 as P
    import rgb_experiment_b200.shim.utils as U
    ei, w = U.add_remaining_self_loops(H.ARSL_EI.to(DEV), H.ARSL_W.float().to(DEV), 1.0, 3)
    assert torch.equal(ei.cpu(), H.ARSL_OUT_EI) and close(w, H.ARSL_OUT_W.float(), 0)
    g = P.Graph(H.G_EI.to(DEV), H.G_N, P.LOOP_ADD_REMAINING)
    assert torch.equal(g.edge_index().cpu(), H.G_NORM_EI)
    assert close(g.to_edge_order(g.gcn_val(False)), H.G_NORM_W)


def test_propagate_appnp_mean():
    import rgb_experiment_b200 as P
    x = H.G_X.float().to(DEV)
    g = P.Graph(H.G_EI.to(DEV), H.G_N, P.LOOP_ADD_REMAINING)
    assert close(P.ops.propagate(x, g, "gcn"), H.G_PROP)
    for fold in (False, True):
        assert close(P.ops.appnp(x, g, 1, 0.1, fold), H.G_APPNP1)
    gm = P.Graph(H.G_EI.to(DEV), H.G_N, P.LOOP_REMOVE_THEN_ADD)
    assert close(P.ops.propagate(x, gm, "mean"), H.G_MEAN)


def test_gat_edge_softmax():
    import rgb_experiment_b200 as P
    g = P.Graph(H.G_EI.to(DEV), H.G_N, P.LOOP_REMOVE_THEN_ADD)
    assert close(P.ops.gat(H.G_X.float().to(DEV), H.GAT_AS.float().to(DEV), H.GAT_AD.float().to(DEV), g, 1, 1, 0.2), H.GAT_OUT)
    assert close(P.ops.gat(H.G_X.float().to(DEV), H.GAT_AS.float().to(DEV), H.GAT_AD_NEG.float().to(DEV), g, 1, 1, 0.2), H.GAT_OUT_NEG)


def test_correct_and_smooth():
    import importlib
    PL = importlib.import_module("rgb_experiment_b200.shim.nn")
    args = (H.CS_YTRUE.to(DEV), H.CS_MASK.to(DEV), H.CS_EI.to(DEV))
    auto = PL.CorrectAndSmooth(1, 0.5, 1, 0.5, autoscale=True)
    assert close(auto.correct(H.CS_YSOFT.float().to(DEV), *args), H.CS_CORRECT_AUTO)
    fixed = PL.CorrectAndSmooth(1, 0.5, 1, 0.5, autoscale=False, scale=1.0)
    assert close(fixed.correct(H.CS_YSOFT.float().to(DEV), *args), H.CS_CORRECT_FIXED)
    assert close(auto.smooth(H.CS_YSOFT.float().to(DEV), *args), H.CS_SMOOTH)


def test_coalesce_and_to_undirected():
    import rgb_experiment_b200.shim.utils as U
    assert torch.equal(U.coalesce(H.CO_EI.to(DEV), None, 3, 3)[0].cpu(), H.CO_OUT)
    assert torch.equal(U.to_undirected(H.UND_EI.to(DEV), 3).cpu(), H.UND_OUT)
