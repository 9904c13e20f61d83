#!/usr/bin/env python
"""A few K-hop calls on the Cora-shaped graph (the command profiled by ncu for khop_cluster_kernel):
    python tools/run_khop_small.py [K] [F]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rgb_experiment_b200 as P            # noqa: E402
import rgb_experiment_b200.synth as S      # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 50
F = int(sys.argv[2]) if len(sys.argv) > 2 else 7
dev = torch.device("cuda:0")
sg = S.make_named("cora", device=dev, features=False)
g0 = P.Graph(sg.edge_index, sg.num_nodes, P.LOOP_NONE)
y0 = torch.rand(sg.num_nodes, F, device=dev)
for _ in range(3):
    out = P.ops.label_propagation(g0, y0, K, 0.8)
torch.cuda.synchronize()
print("ok", float(out.sum()), P._lib.lib().rgbmp_khop_cta_calls())
