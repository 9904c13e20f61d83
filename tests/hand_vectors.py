"""Hand-computed known answers on graphs of three nodes -- every expected number below was derived on paper from the
published operator definitions and is written as a literal or as scalar `math` arithmetic,
never through torch, oracle/ or the CUDA library.  They pin the operators that no reference-held vector pins (PyG is
absent): loop edits with pre-existing loops, gcn_norm, mean aggregation, the GAT edge softmax, Correct&Smooth
(autoscale on and off), coalesce / to_undirected.  Used by tests/test_hand_pins.py (oracle, CPU) and
tests/test_gpu_hand_pins.py (CUDA)."""
import math

import torch


def _f(v):
    return torch.tensor(v, dtype=torch.float64)


S2, S3 = math.sqrt(2.0), math.sqrt(3.0)

# ---- A3 add_remaining_self_loops with pre-existing (and duplicated) loops ----------------------------------------
# edges 0->1 (.5), 1->1 (2.), 1->2 (.7), 2->2 (3.), 1->1 (4.): the non-loops keep their order, then one loop per node;
# node 0 gets the fill value, node 1 the LAST of its two loop weights, node 2 its own
ARSL_EI = torch.tensor([[0, 1, 1, 2, 1], [1, 1, 2, 2, 1]])
ARSL_W = _f([0.5, 2.0, 0.7, 3.0, 4.0])
ARSL_OUT_EI = torch.tensor([[0, 1, 0, 1, 2], [1, 2, 0, 1, 2]])
ARSL_OUT_W = _f([0.5, 0.7, 1.0, 4.0, 3.0])

# ---- A4 gcn_norm, propagation, APPNP, mean on the directed graph 0->1, 1->2, 2->1 --------------------------------
G_EI = torch.tensor([[0, 1, 2], [1, 2, 1]])
G_N = 3
# in-degrees with the added loops: node 0: 1, node 1: 3 (from 0, from 2, loop), node 2: 2  ->  dinv = 1, 1/sqrt3, 1/sqrt2
G_NORM_EI = torch.tensor([[0, 1, 2, 0, 1, 2], [1, 2, 1, 0, 1, 2]])
G_NORM_W = _f([1 / S3, 1 / (S3 * S2), 1 / (S2 * S3), 1.0, 1 / 3, 1 / 2])
G_X = _f([[1.0], [2.0], [4.0]])
G_PROP = _f([[1.0], [1 / S3 * 1 + 1 / (S2 * S3) * 4 + 2 / 3], [1 / (S3 * S2) * 2 + 0.5 * 4]])      # A_hat x
G_APPNP1 = 0.9 * G_PROP + 0.1 * G_X                                                                             # K=1, alpha=.1
G_MEAN = _f([[1.0], [7 / 3], [3.0]])        # remove + add loops, mean: (1+4+2)/3, (2+4)/2

# ---- A10/A11 GAT edge softmax on the same graph (loops added), one head, one channel ------------------------------
# a_dst = 0, a_src = ln 1, ln 2, ln 4: target 1 sees e = (0, ln 4, ln 2) from sources (0, 2, 1) -> alpha = 1/7, 4/7, 2/7
GAT_AS = _f([[0.0], [math.log(2.0)], [math.log(4.0)]])
GAT_AD = torch.zeros(3, 1, dtype=torch.float64)
GAT_OUT = _f([[1.0], [(1 * 1 + 4 * 4 + 2 * 2) / 7], [(1 * 2 + 2 * 4) / 3]])
# a_dst[1] = -3 makes all three logits of target 1 negative: leaky_relu multiplies them by 0.2 before the softmax
GAT_AD_NEG = _f([[0.0], [-3.0], [0.0]])
_e = [0.2 * (0.0 - 3.0), 0.2 * (math.log(4.0) - 3.0), 0.2 * (math.log(2.0) - 3.0)]      # sources 0, 2, 1
_p = [math.exp(v - max(_e)) for v in _e]
GAT_OUT_NEG = _f([[1.0], [(_p[0] * 1 + _p[1] * 4 + _p[2] * 2) / (sum(_p) + 1e-16)], [(1 * 2 + 2 * 4) / 3]])

# ---- A15 Correct & Smooth on the undirected path 0 - 1 - 2, one layer each, alpha = .5 ----------------------------
CS_EI = torch.tensor([[0, 1, 1, 2], [1, 0, 2, 1]])
CS_YSOFT = _f([[0.6, 0.4], [0.5, 0.5], [0.2, 0.8]])
CS_MASK = torch.tensor([True, False, False])
CS_YTRUE = torch.tensor([1])                            # node 0 is class 1
# gcn_norm without loops: deg = 1, 2, 1 -> every edge weighs 1/sqrt2.  E0 = [(-.6, .6), 0, 0]
# correct: out = .5 * A_hat E0 + .5 * E0 = [(-.3, .3), (-.3/sqrt2 , .3/sqrt2), 0]; sigma = 1.2; scale = 2, 1.2/(.6/sqrt2), inf -> 1
CS_CORRECT_AUTO = _f([[0.0, 1.0], [0.5 - 0.6, 0.5 + 0.6], [0.2, 0.8]])
# fixed scale 1: the training row is reset to its error after the layer: smoothed = [(-.6, .6), (-.3/sqrt2, .3/sqrt2), 0]
CS_CORRECT_FIXED = _f([[0.0, 1.0], [0.5 - 0.3 / S2, 0.5 + 0.3 / S2], [0.2, 0.8]])
# smooth: y0 = [(0,1), (.5,.5), (.2,.8)]; out = .5 * A_hat y0 + .5 * y0
CS_SMOOTH = _f([[0.25 / S2, 0.25 / S2 + 0.5], [0.1 / S2 + 0.25, 0.9 / S2 + 0.25], [0.25 / S2 + 0.1, 0.25 / S2 + 0.4]])

# ---- A16 coalesce / to_undirected ---------------------------------------------------------------------------------
CO_EI = torch.tensor([[2, 0, 2, 1, 0], [1, 1, 1, 0, 1]])
CO_OUT = torch.tensor([[0, 1, 2], [1, 0, 1]])
UND_EI = torch.tensor([[0, 2], [1, 1]])
UND_OUT = torch.tensor([[0, 1, 1, 2], [1, 0, 2, 1]])
