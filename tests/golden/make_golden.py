"""Generate golden vectors by EXECUTING the reference's own in-tree, PyG-free functions.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

Reference code executed verbatim (imported from /root/reference, never copied):
  * rgb_experiment/itexperiments.py:671-696  edge_index2sparse_matrix, normalize_adj,
                                             sparse_mx_to_torch_sparse_tensor
  * rgb_experiment/itexperiments.py:698-719  label_propagation
  * rgb_experiment/models/pta.py:79-84       PTA.inference
  * rgb_experiment/models/dagnn.py:12-31     gcn_norm  (its two helpers, add_remaining_self_loops
                                             and scatter_add -- PyG/torch_scatter functions absent
                                             here -- come from tests/golden/independent_stubs.py:
                                             plain Python loops that share NO code with oracle/)
  * rgb_experiment/models/dagnn.py:34-65     Prop.forward / message (loop MessagePassing base from the stubs)
  * rgb_experiment/models/graphsage.py:36-62 my_SAGEConv.forward (loop MessagePassing base from the stubs)
  * rgb_experiment/utils/mask.py:10-21       get_whole_mask (decides the split)

Outputs: tests/golden/*.npz (small; committed).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import shim  # noqa: E402  (only its matplotlib stub is used)
sys.path.insert(0, HERE)
import independent_stubs  # noqa: E402


def small_graph(seed, n, e, self_loops=0, dups=0):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(n, (e,), generator=g)
    dst = torch.randint(n, (e,), generator=g)
    keep = src != dst
    src, dst = src[keep], dst[keep]
    if self_loops:
        l = torch.randint(n, (self_loops,), generator=g)
        pos = torch.randint(src.numel() + 1, (1,), generator=g).item()
        src = torch.cat([src[:pos], l, src[pos:]])
        dst = torch.cat([dst[:pos], l, dst[pos:]])
    if dups:
        src = torch.cat([src, src[:dups]])
        dst = torch.cat([dst, dst[:dups]])
    return torch.stack([src, dst]).long()


def main():
    shim.install_matplotlib_stub()
    independent_stubs.install()
    sys.path.insert(0, "/root/reference")
    import scipy.sparse as sp
    from rgb_experiment import itexperiments as IT
    from rgb_experiment.models import dagnn as DAG
    from rgb_experiment.models import graphsage as SAGE
    from rgb_experiment.models.pta import PTA
    from rgb_experiment.utils import get_whole_mask

    cases = [
        dict(name="g0", seed=1, n=23, e=90, self_loops=0, dups=0, C=4),
        dict(name="g1_loops", seed=2, n=40, e=300, self_loops=5, dups=0, C=5),
        dict(name="g2_dups", seed=3, n=64, e=700, self_loops=3, dups=40, C=7),
        dict(name="g3_isolated", seed=4, n=97, e=150, self_loops=0, dups=0, C=3),
    ]
    for c in cases:
        torch.manual_seed(100 + c["seed"])
        n, C = c["n"], c["C"]
        ei = small_graph(c["seed"], n, c["e"], c["self_loops"], c["dups"])
        out = {"edge_index": ei.numpy(), "num_nodes": np.int64(n)}

        # --- PTA: scipy normalisation, verbatim (itexperiments.py:354-357) ---
        adj = IT.edge_index2sparse_matrix(ei, n)
        adj = adj + sp.eye(adj.shape[0])
        adj = IT.normalize_adj(adj)
        out["pta_adj_dense"] = np.asarray(adj.todense(), dtype=np.float64)
        adj_t = IT.sparse_mx_to_torch_sparse_tensor(adj)

        labels = torch.randint(C, (n,))
        labels[0] = C - 1                                 # make sure max label present
        idx = torch.randperm(n)[: max(2, n // 3)]
        K, alpha = 10, 0.1
        y = IT.label_propagation(adj_t, labels, idx, K, alpha, torch.device("cpu"))
        out.update(pta_labels=labels.numpy(), pta_idx=idx.numpy(), pta_K=np.int64(K),
                   pta_alpha=np.float64(alpha), pta_lp=y.numpy())

        h = torch.randn(n, C)
        model = PTA(nfeat=3, nhid=4, nclass=C, dropout=0.0, epsilon=100, K=K, alpha=alpha)
        out.update(pta_h=h.numpy(), pta_inference=model.inference(h, adj_t).numpy())

        # --- dagnn.gcn_norm verbatim (dagnn.py:12-31) ---
        ei2, w = DAG.gcn_norm(ei, None, n, dtype=torch.float32)
        out.update(gcn_norm_edge_index=ei2.numpy(), gcn_norm_weight=w.numpy())
        ei3, w3 = DAG.gcn_norm(ei, None, n, add_self_loops=False, dtype=torch.float32)
        out.update(gcn_norm_noloop_weight=w3.numpy())

        # --- dagnn.Prop verbatim (dagnn.py:34-65) ---
        prop = DAG.Prop(C, 4)
        with torch.no_grad():
            x = torch.randn(n, C)
            out.update(prop_x=x.numpy(), prop_proj_w=prop.proj.weight.numpy().copy(),
                       prop_proj_b=prop.proj.bias.numpy().copy(), prop_K=np.int64(4),
                       prop_out=prop(x, ei).numpy())

        # --- graphsage.my_SAGEConv verbatim (graphsage.py:36-62) ---
        conv = SAGE.my_SAGEConv(6, 5)
        with torch.no_grad():
            x = torch.randn(n, 6)
            out.update(sage_x=x.numpy(), sage_out=conv(x, ei).numpy(),
                       sage_wl=conv.lin_l.weight.numpy().copy(), sage_bl=conv.lin_l.bias.numpy().copy(),
                       sage_wr=conv.lin_r.weight.numpy().copy(), sage_br=conv.lin_r.bias.numpy().copy())

        # --- split masks (utils/mask.py:10-21) ---
        tr, va, te = get_whole_mask(labels, "6-2-2", 123456789)
        out.update(mask_train=tr.numpy(), mask_val=va.numpy(), mask_test=te.numpy())

        path = os.path.join(HERE, c["name"] + ".npz")
        np.savez_compressed(path, **out)
        print("wrote", path, {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
