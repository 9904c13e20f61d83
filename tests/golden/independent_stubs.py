"""Stand-ins for the few torch_geometric / torch_scatter names the reference's IN-TREE code needs when
tests/golden/make_golden.py executes it -- written as plain Python / numpy LOOPS from the published semantics
(SURVEY.md Appendix A1-A6), deliberately sharing NO code with oracle/ (round-1 review: three of the four reference
goldens ran reference code on top of the oracle's own helpers, so they could not pin those helpers).

    add_remaining_self_loops / remove_self_loops / add_self_loops   (A1-A3; dagnn.py:22-23, graphsage.py:55-56)
    scatter_add                                                    (A5; dagnn.py:28)
    MessagePassing(aggr='add'|'mean').propagate                    (A6; dagnn.py:36,46,57-59, graphsage.py:39,58)

Everything else the reference imports at module load (GCNConv, GATConv, ... -- never called by the code the golden
script runs) is a placeholder class that raises when constructed.
"""
import inspect
import sys
import types

import numpy as np
import torch


def remove_self_loops(edge_index, edge_attr=None):
    keep = [k for k in range(edge_index.size(1)) if int(edge_index[0, k]) != int(edge_index[1, k])]
    idx = torch.tensor(keep, dtype=torch.long)
    return edge_index[:, idx], (None if edge_attr is None else edge_attr[idx])


def add_self_loops(edge_index, edge_weight=None, fill_value=1.0, num_nodes=None):
    n = int(num_nodes)
    rows = [int(v) for v in edge_index[0]] + list(range(n))
    cols = [int(v) for v in edge_index[1]] + list(range(n))
    w = None
    if edge_weight is not None:
        w = torch.tensor([float(v) for v in edge_weight] + [float(fill_value)] * n, dtype=edge_weight.dtype)
    return torch.tensor([rows, cols], dtype=torch.long), w


def add_remaining_self_loops(edge_index, edge_weight=None, fill_value=1.0, num_nodes=None):
    n = int(num_nodes)
    rows, cols, ws = [], [], []
    loop_w = [float(fill_value)] * n
    for k in range(edge_index.size(1)):
        r, c = int(edge_index[0, k]), int(edge_index[1, k])
        if r != c:
            rows.append(r)
            cols.append(c)
            if edge_weight is not None:
                ws.append(float(edge_weight[k]))
        elif edge_weight is not None:
            loop_w[r] = float(edge_weight[k])              # an existing loop keeps its weight (last one wins)
    rows += list(range(n))
    cols += list(range(n))
    w = None
    if edge_weight is not None:
        w = torch.tensor(np.asarray(ws + loop_w, dtype=np.float64), dtype=edge_weight.dtype)
    return torch.tensor([rows, cols], dtype=torch.long), w


def scatter_add(src, index, dim=0, out=None, dim_size=None):
    """Sequential accumulation in input order, in the dtype of `src` (what torch_scatter does on the CPU)."""
    assert dim in (0, -1) and (src.dim() == 1 or dim == 0)
    n = int(dim_size) if dim_size is not None else int(index.max()) + 1
    res = np.zeros((n,) + tuple(src.shape[1:]), dtype=src.detach().numpy().dtype)
    s = src.detach().numpy()
    for k in range(s.shape[0]):
        res[int(index[k])] = res[int(index[k])] + s[k]
    return torch.from_numpy(res)


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2):
        super().__init__()
        assert flow == "source_to_target"
        self.aggr = aggr

    def propagate(self, edge_index, size=None, **kwargs):
        names = list(inspect.signature(self.message).parameters)
        x = kwargs["x"]
        n = x.size(0)
        args = {}
        for nm in names:
            if nm.endswith("_j"):
                args[nm] = torch.stack([kwargs[nm[:-2]][int(j)] for j in edge_index[0]])
            elif nm.endswith("_i"):
                args[nm] = torch.stack([kwargs[nm[:-2]][int(i)] for i in edge_index[1]])
            else:
                args[nm] = kwargs[nm]
        msg = self.message(**args).detach().numpy()
        out = np.zeros((n,) + msg.shape[1:], dtype=msg.dtype)
        cnt = np.zeros(n, dtype=msg.dtype)
        for k in range(msg.shape[0]):
            i = int(edge_index[1, k])
            out[i] = out[i] + msg[k]
            cnt[i] += 1
        if self.aggr == "mean":
            cnt[cnt < 1] = 1
            out = out / cnt.reshape((-1,) + (1,) * (out.ndim - 1))
        return torch.from_numpy(out)

    def message(self, x_j):
        return x_j


def _placeholder(name):
    def __init__(self, *a, **k):
        raise RuntimeError(f"{name} is a placeholder: the golden script never constructs it")
    return type(name, (torch.nn.Module,), {"__init__": __init__})


def install():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        return m

    conv_names = ("GCNConv", "SAGEConv", "GATConv", "SuperGATConv", "APPNP", "SGConv", "FAConv", "GINConv", "GatedGraphConv")
    ph = {n: _placeholder(n) for n in conv_names}
    conv = mod("torch_geometric.nn.conv", MessagePassing=MessagePassing, **ph)
    nn_ = mod("torch_geometric.nn", conv=conv, MessagePassing=MessagePassing, CorrectAndSmooth=_placeholder("CorrectAndSmooth"), **ph)
    utils = mod("torch_geometric.utils", remove_self_loops=remove_self_loops, add_self_loops=add_self_loops,
                add_remaining_self_loops=add_remaining_self_loops, to_undirected=None, to_networkx=None)

    class Data:                                           # itexperiments.py:23 only needs the name at import time
        pass

    data = mod("torch_geometric.data", Data=Data)
    tg = mod("torch_geometric", nn=nn_, utils=utils, data=data, __version__="independent-stubs")
    ts = mod("torch_scatter", scatter_add=scatter_add)
    tsp = mod("torch_sparse", coalesce=None)
    sys.modules.update({"torch_geometric": tg, "torch_geometric.nn": nn_, "torch_geometric.nn.conv": conv,
                        "torch_geometric.utils": utils, "torch_geometric.data": data, "torch_scatter": ts,
                        "torch_sparse": tsp})
