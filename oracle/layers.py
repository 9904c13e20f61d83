"""nn.Module restatement of the PyG layer classes the reference imports (TEST INFRASTRUCTURE).

Import sites in the reference: models/gcn.py:3, graphsage.py:3-4, graphsage2.py:5, gat.py:3,
supergat.py:6, appnp_stack.py:3, sgc.py:4, dagnn.py:7-10, fagcn.py:4, ggnn.py:3, gin.py:6,
itexperiments.py:21-23.  Semantics: SURVEY.md Appendix A (PyG 1.7-2.0.x).  Everything here
is plain CPU torch built on ``oracle.pyg_restated``; the product layers in
``rgb-experiment_b200/shim`` create their parameters in the same order with the same
initialisers so that a fixed seed gives both stacks identical weights.
"""
from __future__ import annotations

import inspect
import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from . import pyg_restated as R


class Data:
    """A17: attribute bag (itexperiments.py:188-189,199,258,315; rd2pd.py:127)."""

    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs):
        self.x = x
        self.edge_index = edge_index
        self.edge_attr = edge_attr
        self.y = y
        self.pos = pos
        for k, v in kwargs.items():
            setattr(self, k, v)

    def __getattr__(self, name):            # only reached for missing attributes
        raise AttributeError(name)

    @property
    def num_nodes(self):
        if self.__dict__.get("x") is not None:
            return self.x.size(0)
        ei = self.__dict__.get("edge_index")
        return int(ei.max()) + 1 if ei is not None and ei.numel() else 0

    @property
    def num_node_features(self):
        x = self.__dict__.get("x")
        if x is None:
            return 0
        return 1 if x.dim() == 1 else x.size(1)

    num_features = num_node_features

    @property
    def num_edges(self):
        ei = self.__dict__.get("edge_index")
        return 0 if ei is None else ei.size(1)

    @property
    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None]

    def clone(self):
        out = self.__class__.__new__(self.__class__)
        for k, v in self.__dict__.items():
            out.__dict__[k] = v.clone() if torch.is_tensor(v) else v
        return out

    def to(self, device, *args, **kwargs):
        for k, v in self.__dict__.items():
            if torch.is_tensor(v):
                self.__dict__[k] = v.to(device, *args, **kwargs)
        return self

    def cpu(self):
        return self.to("cpu")

    def __repr__(self):
        parts = [f"{k}={list(v.shape) if torch.is_tensor(v) else v}" for k, v in self.__dict__.items()
                 if v is not None]
        return f"Data({', '.join(parts)})"


class MessagePassing(nn.Module):
    """A6.  Generic gather -> message -> scatter (graphsage.py:36-62, dagnn.py:34-65)."""

    def __init__(self, aggr: Optional[str] = "add", flow: str = "source_to_target", node_dim: int = -2):
        super().__init__()
        assert flow == "source_to_target"
        self.aggr = aggr
        self.flow = flow
        self.node_dim = node_dim

    def propagate(self, edge_index: Tensor, size=None, **kwargs):
        row, col = edge_index[0], edge_index[1]
        N = None
        for v in kwargs.values():
            if torch.is_tensor(v) and v.dim() >= 2:
                N = v.size(0)
                break
        if size is not None:
            N = size[1] if isinstance(size, (tuple, list)) else size
        params = inspect.signature(self.message).parameters
        args = {}
        for name in params:
            if name.endswith("_j"):
                args[name] = kwargs[name[:-2]].index_select(0, row)
            elif name.endswith("_i"):
                args[name] = kwargs[name[:-2]].index_select(0, col)
            elif name == "index" or name == "edge_index_i":
                args[name] = col
            elif name == "edge_index_j":
                args[name] = row
            elif name == "ptr":
                args[name] = None
            elif name == "size_i":
                args[name] = N
            elif name in kwargs:
                args[name] = kwargs[name]
        out = self.message(**args)
        out = R.scatter(out, col, dim=0, dim_size=N, reduce=self.aggr)
        return self.update(out)

    def message(self, x_j):
        return x_j

    def update(self, inputs):
        return inputs


class GCNConv(MessagePassing):
    """A7 (models/gcn.py:18-21)."""

    def __init__(self, in_channels, out_channels, improved=False, cached=False,
                 add_self_loops=True, normalize=True, bias=True, **kwargs):
        kwargs.setdefault("aggr", "add")
        super().__init__(**kwargs)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.improved, self.cached = improved, cached
        self.add_self_loops, self.normalize = add_self_loops, normalize
        self.lin = nn.Linear(in_channels, out_channels, bias=False)
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        R.glorot_(self.lin.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, x, edge_index, edge_weight=None):
        if self.normalize:
            edge_index, edge_weight = R.gcn_norm(edge_index, edge_weight, x.size(0), self.improved,
                                                 self.add_self_loops, dtype=x.dtype)
        x = self.lin(x)
        out = R.propagate(edge_index, x, edge_weight, "add")
        if self.bias is not None:
            out = out + self.bias
        return out


class SAGEConv(MessagePassing):
    """A14 (models/graphsage2.py:20-23): mean at in_channels width, lin_l(bias) + lin_r."""

    def __init__(self, in_channels, out_channels, normalize=False, root_weight=True, bias=True, **kwargs):
        kwargs.setdefault("aggr", "mean")
        super().__init__(**kwargs)
        self.normalize, self.root_weight = normalize, root_weight
        self.lin_l = nn.Linear(in_channels, out_channels, bias=bias)
        if root_weight:
            self.lin_r = nn.Linear(in_channels, out_channels, bias=False)

    def forward(self, x, edge_index):
        out = R.propagate(edge_index, x, None, "mean")
        out = self.lin_l(out)
        if self.root_weight:
            out = out + self.lin_r(x)
        if self.normalize:
            out = F.normalize(out, p=2.0, dim=-1)
        return out


class GATConv(MessagePassing):
    """A10 (models/gat.py:18-21)."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2,
                 dropout=0.0, add_self_loops=True, bias=True, **kwargs):
        kwargs.setdefault("aggr", "add")
        super().__init__(**kwargs)
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.negative_slope, self.dropout = concat, negative_slope, dropout
        self.add_self_loops = add_self_loops
        self.lin = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(heads * out_channels if concat else out_channels))
        else:
            self.bias = None
        self.reset_parameters()

    def reset_parameters(self):
        R.glorot_(self.lin.weight)
        R.glorot_(self.att_src)
        R.glorot_(self.att_dst)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, x, edge_index):
        H, C = self.heads, self.out_channels
        xp = self.lin(x).view(-1, H, C)
        a_s = (xp * self.att_src).sum(-1)
        a_d = (xp * self.att_dst).sum(-1)
        N = xp.size(0)
        ei = R.edit_loops(edge_index, N, R.LOOP_REMOVE_THEN_ADD) if self.add_self_loops else edge_index
        row, col = ei[0], ei[1]
        e = F.leaky_relu(a_s[row] + a_d[col], self.negative_slope)
        alpha = R.softmax(e, col, num_nodes=N)
        alpha = F.dropout(alpha, p=self.dropout, training=self.training)
        out = R.scatter_add(xp[row] * alpha.unsqueeze(-1), col, dim=0, dim_size=N)
        out = out.view(-1, H * C) if self.concat else out.mean(dim=1)
        if self.bias is not None:
            out = out + self.bias
        return out


def dropout_adj(edge_index, p=0.5, training=True):
    """PyG utils.dropout_adj (no edge_attr, not force_undirected): Bernoulli(1-p) keep mask."""
    if not training or p == 0.0:
        return edge_index, None
    mask = torch.full((edge_index.size(1),), 1 - p, dtype=torch.float, device=edge_index.device)
    mask = torch.bernoulli(mask).to(torch.bool)
    return edge_index[:, mask], None


def negative_sampling(edge_index, num_nodes, num_neg_samples):
    """Restated PyG negative_sampling (sparse method): draw candidate (i,j) pairs uniformly,
    reject existing edges, up to 3 rounds.  Self pairs are not excluded upstream either."""
    N = num_nodes
    idx = edge_index[0] * N + edge_index[1]
    size = N * N
    num_neg = min(int(num_neg_samples), size - idx.numel())
    if num_neg <= 0:
        return edge_index.new_empty((2, 0))
    alpha = abs(1 / (1 - 1.1 * (edge_index.size(1) / size)))
    sample_size = int(alpha * num_neg)
    neg = None
    for _ in range(3):
        rnd = torch.randint(size, (sample_size,), dtype=torch.long, device=edge_index.device)
        mask = ~torch.isin(rnd, idx)
        rnd = rnd[mask]
        neg = rnd if neg is None else torch.cat([neg, rnd])
        if neg.numel() >= num_neg:
            neg = neg[:num_neg]
            break
    return torch.stack([neg // N, neg % N], dim=0)


class SuperGATConv(MessagePassing):
    """A12, MX attention (models/supergat.py:15-21,26,29)."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2,
                 dropout=0.0, add_self_loops=True, bias=True, attention_type="MX",
                 neg_sample_ratio=0.5, edge_sample_ratio=1.0, is_undirected=False, **kwargs):
        kwargs.setdefault("aggr", "add")
        super().__init__(**kwargs)
        assert attention_type in ("MX", "SD")
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.concat, self.negative_slope, self.dropout = concat, negative_slope, dropout
        self.add_self_loops, self.attention_type = add_self_loops, attention_type
        self.neg_sample_ratio, self.edge_sample_ratio = neg_sample_ratio, edge_sample_ratio
        self.is_undirected = is_undirected
        self.lin = nn.Linear(in_channels, heads * out_channels, bias=False)
        if attention_type == "MX":
            self.att_l = nn.Parameter(torch.empty(1, heads, out_channels))
            self.att_r = nn.Parameter(torch.empty(1, heads, out_channels))
        else:
            self.register_parameter("att_l", None)
            self.register_parameter("att_r", None)
        self.att_x = self.att_y = None
        if bias:
            self.bias = nn.Parameter(torch.empty(heads * out_channels if concat else out_channels))
        else:
            self.bias = None
        self.reset_parameters()

    def reset_parameters(self):
        R.glorot_(self.lin.weight)
        if self.att_l is not None:
            R.glorot_(self.att_l)
            R.glorot_(self.att_r)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def get_attention(self, x_i, x_j, return_logits=False):
        if self.attention_type == "MX":
            logits = (x_i * x_j).sum(dim=-1)
            if return_logits:
                return logits
            alpha = (x_j * self.att_l).sum(-1) + (x_i * self.att_r).sum(-1)
            alpha = alpha * logits.sigmoid()
        else:
            alpha = (x_i * x_j).sum(dim=-1) / math.sqrt(self.out_channels)
            if return_logits:
                return alpha
        return F.leaky_relu(alpha, self.negative_slope)

    def forward(self, x, edge_index, neg_edge_index=None):
        N, H, C = x.size(0), self.heads, self.out_channels
        if self.add_self_loops:
            edge_index = R.edit_loops(edge_index, N, R.LOOP_REMOVE_THEN_ADD)
        xp = self.lin(x).view(-1, H, C)
        row, col = edge_index[0], edge_index[1]
        alpha = self.get_attention(xp[col], xp[row])
        alpha = R.softmax(alpha, col, num_nodes=N)
        alpha = F.dropout(alpha, p=self.dropout, training=self.training)
        out = R.scatter_add(xp[row] * alpha.view(-1, H, 1), col, dim=0, dim_size=N)
        if self.training:
            pos_ei, _ = dropout_adj(edge_index, p=1.0 - self.edge_sample_ratio, training=True)
            ei_for_neg = R.to_undirected(edge_index, N) if not self.is_undirected else edge_index
            if neg_edge_index is None:
                num_neg = int(self.neg_sample_ratio * self.edge_sample_ratio * edge_index.size(1))
                neg_edge_index = negative_sampling(ei_for_neg, N, num_neg)
            pos_att = self.get_attention(xp[pos_ei[1]], xp[pos_ei[0]], return_logits=True)
            neg_att = self.get_attention(xp[neg_edge_index[1]], xp[neg_edge_index[0]], return_logits=True)
            self.att_x = torch.cat([pos_att, neg_att], dim=0)
            self.att_y = self.att_x.new_zeros(self.att_x.size(0))
            self.att_y[:pos_ei.size(1)] = 1.0
        out = out.view(-1, H * C) if self.concat else out.mean(dim=1)
        if self.bias is not None:
            out = out + self.bias
        return out

    def get_attention_loss(self):
        if not self.training:
            return torch.tensor([0], device=self.lin.weight.device)
        return F.binary_cross_entropy_with_logits(self.att_x.mean(dim=-1), self.att_y)


class APPNP(MessagePassing):
    """A8 (models/appnp_stack.py:22)."""

    def __init__(self, K, alpha, dropout=0.0, cached=False, add_self_loops=True, normalize=True, **kwargs):
        kwargs.setdefault("aggr", "add")
        super().__init__(**kwargs)
        self.K, self.alpha, self.dropout = K, alpha, dropout
        self.add_self_loops, self.normalize = add_self_loops, normalize

    def forward(self, x, edge_index, edge_weight=None):
        if self.normalize:
            edge_index, edge_weight = R.gcn_norm(edge_index, edge_weight, x.size(0), False,
                                                 self.add_self_loops, dtype=x.dtype)
        h = x
        for _ in range(self.K):
            w = edge_weight
            if self.dropout > 0 and self.training:
                w = F.dropout(w, p=self.dropout)
            x = R.propagate(edge_index, x, w, "add")
            x = x * (1 - self.alpha)
            x = x + self.alpha * h
        return x


class SGConv(MessagePassing):
    """A9 (models/sgc.py:9-10)."""

    def __init__(self, in_channels, out_channels, K=1, cached=False, add_self_loops=True, bias=True, **kwargs):
        kwargs.setdefault("aggr", "add")
        super().__init__(**kwargs)
        self.K, self.cached, self.add_self_loops = K, cached, add_self_loops
        self._cached_x = None
        self.lin = nn.Linear(in_channels, out_channels, bias=bias)

    def forward(self, x, edge_index, edge_weight=None):
        cache = self._cached_x
        if cache is None:
            edge_index, edge_weight = R.gcn_norm(edge_index, edge_weight, x.size(0), False,
                                                 self.add_self_loops, dtype=x.dtype)
            for _ in range(self.K):
                x = R.propagate(edge_index, x, edge_weight, "add")
            if self.cached:
                self._cached_x = x
        else:
            x = cache
        return self.lin(x)


class FAConv(MessagePassing):
    """A13 (models/fagcn.py:15,31)."""

    def __init__(self, channels, eps=0.1, dropout=0.0, cached=False, add_self_loops=True,
                 normalize=True, **kwargs):
        kwargs.setdefault("aggr", "add")
        super().__init__(**kwargs)
        self.channels, self.eps, self.dropout = channels, eps, dropout
        self.add_self_loops, self.normalize = add_self_loops, normalize
        self.att_l = nn.Linear(channels, 1, bias=False)
        self.att_r = nn.Linear(channels, 1, bias=False)

    def forward(self, x, x_0, edge_index, edge_weight=None):
        if self.normalize:
            edge_index, edge_weight = R.gcn_norm(edge_index, None, x.size(0), False,
                                                 self.add_self_loops, dtype=x.dtype)
        row, col = edge_index[0], edge_index[1]
        a_l, a_r = self.att_l(x), self.att_r(x)
        c = (a_l.view(-1)[row] + a_r.view(-1)[col]).tanh()
        c = F.dropout(c, p=self.dropout, training=self.training)
        out = R.scatter_add(x[row] * (c * edge_weight).view(-1, 1), col, dim=0, dim_size=x.size(0))
        if self.eps != 0.0:
            out = out + self.eps * x_0
        return out


class GINConv(MessagePassing):
    """A14 (models/gin.py:14-32): out = nn((1+eps)*x + sum_j x_j)."""

    def __init__(self, nn_module, eps=0.0, train_eps=False, **kwargs):
        kwargs.setdefault("aggr", "add")
        super().__init__(**kwargs)
        self.nn = nn_module
        self.initial_eps = eps
        if train_eps:
            self.eps = nn.Parameter(torch.tensor([float(eps)]))
        else:
            self.register_buffer("eps", torch.tensor([float(eps)]))

    def reset_parameters(self):
        for m in self.nn.modules():
            if m is not self.nn and hasattr(m, "reset_parameters"):
                m.reset_parameters()
        self.eps.data.fill_(self.initial_eps)

    def forward(self, x, edge_index):
        out = R.propagate(edge_index, x, None, "add")
        out = out + (1 + self.eps) * x
        return self.nn(out)


class GatedGraphConv(MessagePassing):
    """A14 (models/ggnn.py:20)."""

    def __init__(self, out_channels, num_layers, aggr="add", bias=True, **kwargs):
        super().__init__(aggr=aggr, **kwargs)
        self.out_channels, self.num_layers = out_channels, num_layers
        self.weight = nn.Parameter(torch.empty(num_layers, out_channels, out_channels))
        self.rnn = nn.GRUCell(out_channels, out_channels, bias=bias)
        bound = 1.0 / math.sqrt(out_channels)
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)

    def forward(self, x, edge_index):
        if x.size(-1) > self.out_channels:
            raise ValueError("input width larger than out_channels")
        if x.size(-1) < self.out_channels:
            x = torch.cat([x, x.new_zeros(x.size(0), self.out_channels - x.size(-1))], dim=1)
        for i in range(self.num_layers):
            m = torch.matmul(x, self.weight[i])
            m = R.propagate(edge_index, m, None, self.aggr)
            x = self.rnn(m, x)
        return x


class LabelPropagation(nn.Module):
    """A15."""

    def __init__(self, num_layers, alpha):
        super().__init__()
        self.num_layers, self.alpha = num_layers, alpha

    @torch.no_grad()
    def forward(self, y, edge_index, mask=None, edge_weight=None, post_step=None):
        if y.dtype == torch.long:
            y = F.one_hot(y.view(-1)).to(torch.float)
        out = y
        if mask is not None:
            out = torch.zeros_like(y)
            out[mask] = y[mask]
        return R.lp_propagate(out, edge_index, self.num_layers, self.alpha, post_step)


class CorrectAndSmooth(nn.Module):
    """A15 (itexperiments.py:520-526)."""

    def __init__(self, num_correction_layers, correction_alpha, num_smoothing_layers,
                 smoothing_alpha, autoscale=True, scale=1.0):
        super().__init__()
        self.autoscale, self.scale = autoscale, scale
        self.prop1 = LabelPropagation(num_correction_layers, correction_alpha)
        self.prop2 = LabelPropagation(num_smoothing_layers, smoothing_alpha)

    @torch.no_grad()
    def correct(self, y_soft, y_true, mask, edge_index, edge_weight=None):
        return R.cs_correct(y_soft, y_true, mask, edge_index, self.prop1.num_layers,
                            self.prop1.alpha, self.autoscale, self.scale)

    @torch.no_grad()
    def smooth(self, y_soft, y_true, mask, edge_index, edge_weight=None):
        return R.cs_smooth(y_soft, y_true, mask, edge_index, self.prop2.num_layers, self.prop2.alpha)


def to_networkx(data, *args, **kwargs):          # imported at dagnn.py:7, never called
    raise NotImplementedError("to_networkx is imported by the reference but never used")
