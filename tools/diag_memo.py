#!/usr/bin/env python
"""Why does (or doesn't) the eval memo hit inside the APPNPStack epoch?  Captures the propagation inputs of two
consecutive eval forwards and reports bitwise equality, fingerprints and memo statistics."""
import json
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as Fn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import rgb_experiment_b200 as P
    import rgb_experiment_b200.memo as M
    import rgb_experiment_b200.synth as S
    dev = torch.device("cuda:0")
    wl = sys.argv[1] if len(sys.argv) > 1 else "products"
    sg = S.make_named(wl, device=dev)
    N, Fin, C = sg.num_nodes, sg.x.size(1), sg.num_classes
    g = P.Graph(sg.edge_index, N, P.LOOP_ADD_REMAINING)
    torch.manual_seed(0)
    lin1, bn, lin2 = nn.Linear(Fin, 64).to(dev), nn.BatchNorm1d(64).to(dev), nn.Linear(64, C).to(dev)
    seen = []

    def fwd():
        h = lin2(bn(lin1(sg.x)))
        seen.append(h)
        return Fn.log_softmax(P.ops.appnp(h, g, 10, 0.1, True), dim=1)

    for m in (lin1, bn, lin2):
        m.train()
    fwd().sum().backward()                  # one training forward so that BN has running statistics
    for m in (lin1, bn, lin2):
        m.eval()
    seen.clear()
    res = {}
    with torch.no_grad():
        t = []
        for i in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fwd()
            e1.record()
            torch.cuda.synchronize()
            t.append(round(e0.elapsed_time(e1), 2))
            res[f"stats_after_{i}"] = dict(M.stats)
        res["eval_ms"] = t
        res["inputs_bitwise_equal"] = [bool(torch.equal(seen[0], seen[i])) for i in (1, 2)]
        res["max_abs_diff"] = [float((seen[0] - seen[i]).abs().max()) for i in (1, 2)]
        res["fingerprints_equal"] = [M.fingerprint((seen[0],)) == M.fingerprint((seen[i],)) for i in (1, 2)]
        res["fingerprint_repeatable_on_one_tensor"] = M.fingerprint((seen[0],)) == M.fingerprint((seen[0],))
        res["held_MB"] = M.held_bytes() >> 20
        a = lin1(sg.x)
        res["lin1_repeatable"] = bool(torch.equal(a, lin1(sg.x)))
        b = bn(a)
        res["bn_repeatable"] = bool(torch.equal(b, bn(a)))
        res["lin2_repeatable"] = bool(torch.equal(lin2(b), lin2(b)))
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
