"""Functional CPU restatement of the PyG operators on the hot path (TEST INFRASTRUCTURE).

Every function cites the reference call site it serves and the SURVEY.md
Appendix-A entry that states the upstream (torch_geometric / torch_scatter /
torch_sparse, 2021-era) semantics.  Conventions:
``edge_index`` is int64 [2, E]; ``row = edge_index[0]`` is the SOURCE j,
``col = edge_index[1]`` is the TARGET i; node i aggregates over edges whose
``col == i``.

All functions are dtype-generic (run them in float64 for the accuracy arbiter)
and use only sequential-in-edge-order CPU ``scatter_add_`` / ``index_add_``, which
is what PyG-on-CPU executes.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
from torch import Tensor

# --------------------------------------------------------------------------------------
# A1-A3: self-loop edits.  Call sites: rgb_experiment/models/graphsage.py:55-56,
# rgb_experiment/models/dagnn.py:22-23.
# --------------------------------------------------------------------------------------


def remove_self_loops(edge_index: Tensor, edge_attr: Optional[Tensor] = None):
    """A1.  Keep edges with row != col, order preserved (graphsage.py:55)."""
    mask = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, mask]
    if edge_attr is None:
        return edge_index, None
    return edge_index, edge_attr[mask]


def add_self_loops(edge_index: Tensor, edge_weight: Optional[Tensor] = None,
                   fill_value: float = 1.0, num_nodes: Optional[int] = None):
    """A2.  Append arange(N) loops unconditionally (graphsage.py:56)."""
    N = _num_nodes(edge_index, num_nodes)
    loop = torch.arange(N, dtype=edge_index.dtype, device=edge_index.device)
    loop = loop.unsqueeze(0).repeat(2, 1)
    if edge_weight is not None:
        lw = edge_weight.new_full((N,), fill_value)
        edge_weight = torch.cat([edge_weight, lw], dim=0)
    return torch.cat([edge_index, loop], dim=1), edge_weight


def add_remaining_self_loops(edge_index: Tensor, edge_weight: Optional[Tensor] = None,
                             fill_value: float = 1.0, num_nodes: Optional[int] = None):
    """A3.  Drop existing loops, append one loop per node; an existing loop's weight is
    kept (last write wins).  Call site: dagnn.py:22-23."""
    N = _num_nodes(edge_index, num_nodes)
    row, col = edge_index[0], edge_index[1]
    mask = row != col
    loop = torch.arange(N, dtype=edge_index.dtype, device=edge_index.device)
    loop = loop.unsqueeze(0).repeat(2, 1)
    if edge_weight is not None:
        inv = ~mask
        lw = edge_weight.new_full((N,), fill_value)
        rem = edge_weight[inv]
        if rem.numel() > 0:
            lw[row[inv]] = rem
        edge_weight = torch.cat([edge_weight[mask], lw], dim=0)
    return torch.cat([edge_index[:, mask], loop], dim=1), edge_weight


def _num_nodes(edge_index: Tensor, num_nodes: Optional[int]) -> int:
    if num_nodes is not None:
        return int(num_nodes)
    return int(edge_index.max()) + 1 if edge_index.numel() > 0 else 0


# --------------------------------------------------------------------------------------
# A5: torch_scatter.  Call site: dagnn.py:28 (scatter_add for the degree).
# --------------------------------------------------------------------------------------


def _bcast_index(index: Tensor, src: Tensor, dim: int) -> Tensor:
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(dim):
            index = index.unsqueeze(0)
    for _ in range(index.dim(), src.dim()):
        index = index.unsqueeze(-1)
    return index.expand_as(src)


def scatter_add(src: Tensor, index: Tensor, dim: int = -1, out: Optional[Tensor] = None,
                dim_size: Optional[int] = None) -> Tensor:
    """A5 ``sum``: zeros(...).scatter_add_(dim, broadcast(index), src)."""
    index = _bcast_index(index, src, dim)
    if out is None:
        size = list(src.size())
        if dim_size is not None:
            size[dim] = dim_size
        elif index.numel() == 0:
            size[dim] = 0
        else:
            size[dim] = int(index.max()) + 1
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)


def scatter(src: Tensor, index: Tensor, dim: int = 0, dim_size: Optional[int] = None,
            reduce: str = "sum") -> Tensor:
    """A5 ``scatter(..., reduce=)`` for sum/add, mean, max (dim=0 is all the path uses)."""
    if reduce in ("sum", "add"):
        return scatter_add(src, index, dim, None, dim_size)
    if reduce == "mean":
        out = scatter_add(src, index, dim, None, dim_size)
        n = out.size(dim)
        ones = torch.ones(index.size(0), dtype=src.dtype, device=src.device)
        count = scatter_add(ones, index, 0, None, n)
        count[count < 1] = 1
        shape = [1] * out.dim()
        shape[dim] = n
        return out / count.view(shape)
    if reduce == "max":
        size = list(src.size())
        size[dim] = dim_size if dim_size is not None else (int(index.max()) + 1 if index.numel() else 0)
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
        idx = _bcast_index(index, src, dim)
        # empty groups stay 0 (torch_scatter semantics); include_self=False ignores the zeros
        return out.scatter_reduce_(dim, idx, src, reduce="amax", include_self=False)
    raise ValueError(reduce)


# --------------------------------------------------------------------------------------
# A4: gcn_norm -- identical to rgb_experiment/models/dagnn.py:12-31 (in-tree copy).
# --------------------------------------------------------------------------------------


def gcn_norm(edge_index: Tensor, edge_weight: Optional[Tensor] = None,
             num_nodes: Optional[int] = None, improved: bool = False,
             add_self_loops: bool = True, dtype=None) -> Tuple[Tensor, Tensor]:
    fill_value = 2.0 if improved else 1.0
    N = _num_nodes(edge_index, num_nodes)
    if edge_weight is None:
        edge_weight = torch.ones((edge_index.size(1),), dtype=dtype, device=edge_index.device)
    if add_self_loops:
        edge_index, edge_weight = add_remaining_self_loops(edge_index, edge_weight, fill_value, N)
    row, col = edge_index[0], edge_index[1]
    deg = scatter_add(edge_weight, col, dim=0, dim_size=N)          # in-degree by TARGET
    dinv = deg.pow_(-0.5)
    dinv.masked_fill_(dinv == float("inf"), 0)
    return edge_index, dinv[row] * edge_weight * dinv[col]


# --------------------------------------------------------------------------------------
# A6: MessagePassing.propagate for the two in-tree message forms
# (graphsage.py:39,58 mean of x_j ; dagnn.py:36,46,57-59 add of norm*x_j).
# --------------------------------------------------------------------------------------


def propagate(edge_index: Tensor, x: Tensor, edge_weight: Optional[Tensor] = None,
              aggr: str = "add", num_nodes: Optional[int] = None) -> Tensor:
    """gather x[row] -> (optional) scale by edge_weight -> scatter over col."""
    N = x.size(0) if num_nodes is None else num_nodes
    row, col = edge_index[0], edge_index[1]
    msg = x.index_select(0, row)
    if edge_weight is not None:
        msg = edge_weight.view(-1, *([1] * (msg.dim() - 1))) * msg
    return scatter(msg, col, dim=0, dim_size=N, reduce=aggr)


# --------------------------------------------------------------------------------------
# bit-exact integer definitions (SURVEY.md 8c "bit-exact definitions")
# --------------------------------------------------------------------------------------

LOOP_NONE, LOOP_ADD, LOOP_ADD_REMAINING, LOOP_REMOVE_THEN_ADD = 0, 1, 2, 3


def edit_loops(edge_index: Tensor, num_nodes: int, loop_mode: int) -> Tensor:
    """The edge list a layer actually aggregates over (kept edges in original order,
    then loops 0..N-1).  Modes: none (LabelPropagation, A15), add (A2),
    add_remaining (A3/A4; GCNConv/APPNP/SGConv/FAConv/dagnn), remove_then_add
    (A1+A2; graphsage.py:55-56, GATConv, SuperGATConv)."""
    if loop_mode == LOOP_NONE:
        return edge_index
    if loop_mode == LOOP_ADD:
        return add_self_loops(edge_index, num_nodes=num_nodes)[0]
    if loop_mode == LOOP_ADD_REMAINING:
        return add_remaining_self_loops(edge_index, None, 1.0, num_nodes)[0]
    if loop_mode == LOOP_REMOVE_THEN_ADD:
        ei, _ = remove_self_loops(edge_index)
        return add_self_loops(ei, num_nodes=num_nodes)[0]
    raise ValueError(loop_mode)


def csr_build(edge_index: Tensor, num_nodes: int, by: str = "dst"):
    """Stable CSR of an (already edited) edge list.
    by='dst': rows = targets, col = sources (forward aggregation order);
    by='src': rows = sources, col = targets (transpose, for the backward).
    Returns (rowptr int64 [N+1], col int64 [nnz], eid int64 [nnz]) with
    perm = argsort(key, stable=True); col = other[perm]; eid = perm."""
    row, col = edge_index[0], edge_index[1]
    key, other = (col, row) if by == "dst" else (row, col)
    perm = torch.argsort(key, stable=True)
    deg = torch.bincount(key, minlength=num_nodes)
    rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(deg, 0)
    return rowptr, other[perm], perm


def degree(edge_index: Tensor, num_nodes: int, by: str = "dst") -> Tensor:
    key = edge_index[1] if by == "dst" else edge_index[0]
    return torch.bincount(key, minlength=num_nodes)


# --------------------------------------------------------------------------------------
# A16: to_undirected / coalesce.  Call sites: itexperiments.py:235-238, rd2pd.py:92-93.
# --------------------------------------------------------------------------------------


def coalesce(index: Tensor, value: Optional[Tensor], m: int, n: int, op: str = "add"):
    """torch_sparse.coalesce: sort by row*n+col, unique; values reduced with ``op``."""
    row, col = index[0], index[1]
    key = row * n + col
    uniq, inv = torch.unique(key, sorted=True, return_inverse=True)
    out_index = torch.stack([uniq // n, uniq % n], dim=0)
    if value is None:
        return out_index, None
    out_val = scatter(value, inv, dim=0, dim_size=uniq.numel(),
                      reduce="sum" if op == "add" else op)
    return out_index, out_val


def to_undirected(edge_index: Tensor, num_nodes: Optional[int] = None) -> Tensor:
    N = _num_nodes(edge_index, num_nodes)
    row, col = edge_index[0], edge_index[1]
    row, col = torch.cat([row, col], 0), torch.cat([col, row], 0)
    return coalesce(torch.stack([row, col], 0), None, N, N)[0]


# --------------------------------------------------------------------------------------
# A11: edge softmax.
# --------------------------------------------------------------------------------------


def softmax(src: Tensor, index: Tensor, ptr=None, num_nodes: Optional[int] = None) -> Tensor:
    N = (int(index.max()) + 1 if index.numel() else 0) if num_nodes is None else num_nodes
    m = scatter(src, index, 0, N, "max")[index]
    out = (src - m).exp()
    s = scatter(out, index, 0, N, "sum")[index]
    return out / (s + 1e-16)


# --------------------------------------------------------------------------------------
# A7 GCNConv, A8 APPNP, A9 SGConv propagation parts (functional; weights applied by caller)
# gcn.py:18-31, appnp_stack.py:22,30, sgc.py:7-13
# --------------------------------------------------------------------------------------


def gcn_propagate(x: Tensor, edge_index: Tensor, add_self_loops_: bool = True) -> Tensor:
    """out[i] = sum_{j->i} dinv[j]*dinv[i] * x[j] over the add_remaining-looped graph."""
    ei, w = gcn_norm(edge_index, None, x.size(0), False, add_self_loops_, dtype=x.dtype)
    return propagate(ei, x, w, "add")


def appnp_propagate(x: Tensor, edge_index: Tensor, K: int, alpha: float) -> Tensor:
    """A8: h = x; K x { x = A_hat x ; x = x*(1-alpha) ; x += alpha*h }."""
    ei, w = gcn_norm(edge_index, None, x.size(0), False, True, dtype=x.dtype)
    h = x
    for _ in range(K):
        x = propagate(ei, x, w, "add")
        x = x * (1 - alpha)
        x = x + alpha * h
    return x


def sgc_propagate(x: Tensor, edge_index: Tensor, K: int, add_self_loops_: bool = True) -> Tensor:
    """A9: K x { x = A_hat x } on the raw features."""
    ei, w = gcn_norm(edge_index, None, x.size(0), False, add_self_loops_, dtype=x.dtype)
    for _ in range(K):
        x = propagate(ei, x, w, "add")
    return x


def dagnn_hops(x: Tensor, edge_index: Tensor, K: int) -> Tensor:
    """dagnn.py:41-49: K hops keeping every hop, stacked [N, K+1, C]."""
    ei, w = gcn_norm(edge_index, None, x.size(0), dtype=x.dtype)
    preds = [x]
    for _ in range(K):
        x = propagate(ei, x, w, "add")
        preds.append(x)
    return torch.stack(preds, dim=1)


def sage_mean(x: Tensor, edge_index: Tensor, loops: bool = True) -> Tensor:
    """graphsage.py:53-58: remove+add self loops, mean of x_j (count clamp(min=1))."""
    if loops:
        ei = edit_loops(edge_index, x.size(0), LOOP_REMOVE_THEN_ADD)
    else:
        ei = edge_index
    return propagate(ei, x, None, "mean")


# --------------------------------------------------------------------------------------
# A10 GATConv attention + aggregate (gat.py:18-21), A12 SuperGAT MX, A13 FAConv
# --------------------------------------------------------------------------------------


def gat_aggregate(xp: Tensor, a_src: Tensor, a_dst: Tensor, edge_index: Tensor,
                  negative_slope: float = 0.2, add_self_loops_: bool = True,
                  alpha_dropout_mask: Optional[Tensor] = None):
    """xp [N,H,C]; a_src,a_dst [N,H] -> out [N,H,C] and alpha [nnz,H] (edge order of the
    remove_then_add edge list)."""
    N = xp.size(0)
    ei = edit_loops(edge_index, N, LOOP_REMOVE_THEN_ADD) if add_self_loops_ else edge_index
    row, col = ei[0], ei[1]
    e = torch.nn.functional.leaky_relu(a_src[row] + a_dst[col], negative_slope)
    alpha = softmax(e, col, num_nodes=N)
    if alpha_dropout_mask is not None:
        alpha = alpha * alpha_dropout_mask
    out = scatter_add(xp[row] * alpha.unsqueeze(-1), col, dim=0, dim_size=N)
    return out, alpha, ei


def supergat_mx_alpha(xp: Tensor, att_l: Tensor, att_r: Tensor, ei: Tensor,
                      negative_slope: float = 0.2):
    """A12 (MX attention): returns (alpha_raw_after_leaky [nnz,H], logits [nnz,H])."""
    row, col = ei[0], ei[1]
    x_j, x_i = xp[row], xp[col]
    logits = (x_i * x_j).sum(dim=-1)
    alpha = (x_j * att_l).sum(-1) + (x_i * att_r).sum(-1)
    alpha = alpha * logits.sigmoid()
    return torch.nn.functional.leaky_relu(alpha, negative_slope), logits


def faconv_aggregate(x: Tensor, x0: Tensor, a_l: Tensor, a_r: Tensor, edge_index: Tensor,
                     eps: float, drop_mask: Optional[Tensor] = None) -> Tensor:
    """A13: gcn_norm; c = tanh(a_l[row] + a_r[col]); out = sum x_j * (c*w) ; += eps*x0."""
    ei, w = gcn_norm(edge_index, None, x.size(0), False, True, dtype=x.dtype)
    row, col = ei[0], ei[1]
    c = (a_l.view(-1)[row] + a_r.view(-1)[col]).tanh()
    if drop_mask is not None:
        c = c * drop_mask
    out = scatter_add(x[row] * (c * w).view(-1, 1), col, dim=0, dim_size=x.size(0))
    return out + eps * x0


# --------------------------------------------------------------------------------------
# A15 LabelPropagation / CorrectAndSmooth (itexperiments.py:520-526)
# --------------------------------------------------------------------------------------


def lp_propagate(out0: Tensor, edge_index: Tensor, num_layers: int, alpha: float,
                 post_step=None) -> Tensor:
    """LP core on an already prepared start matrix: gcn_norm(add_self_loops=False);
    res=(1-a)*out ; L x { out = A_hat out ; out = a*out + res ; out = post_step(out) }."""
    if post_step is None:
        post_step = lambda y: y.clamp_(0.0, 1.0)  # noqa: E731
    N = out0.size(0)
    ei, w = gcn_norm(edge_index, None, N, add_self_loops=False, dtype=out0.dtype)
    out = out0
    res = (1 - alpha) * out
    for _ in range(num_layers):
        out = propagate(ei, out, w, "add")
        out = out * alpha
        out = out + res
        out = post_step(out)
    return out


def cs_correct(y_soft: Tensor, y_true: Tensor, mask: Tensor, edge_index: Tensor,
               num_layers: int, alpha: float, autoscale: bool = True, scale: float = 1.0):
    assert abs(float(y_soft.sum()) / y_soft.size(0) - 1.0) < 1e-2
    numel = int(mask.sum()) if mask.dtype == torch.bool else mask.size(0)
    if y_true.dtype == torch.long:
        y_true = torch.nn.functional.one_hot(y_true.view(-1), y_soft.size(-1)).to(y_soft.dtype)
    error = torch.zeros_like(y_soft)
    error[mask] = y_true - y_soft[mask]
    if autoscale:
        smoothed = lp_propagate(error, edge_index, num_layers, alpha,
                                post_step=lambda x: x.clamp_(-1.0, 1.0))
        sigma = error[mask].abs().sum() / numel
        sc = sigma / smoothed.abs().sum(dim=1, keepdim=True)
        sc[sc.isinf() | (sc > 1000)] = 1.0
        return y_soft + sc * smoothed
    def fix_input(x):
        x[mask] = error[mask]
        return x
    smoothed = lp_propagate(error, edge_index, num_layers, alpha, post_step=fix_input)
    return y_soft + scale * smoothed


def cs_smooth(y_soft: Tensor, y_true: Tensor, mask: Tensor, edge_index: Tensor,
              num_layers: int, alpha: float):
    if y_true.dtype == torch.long:
        y_true = torch.nn.functional.one_hot(y_true.view(-1), y_soft.size(-1)).to(y_soft.dtype)
    y_soft = y_soft.clone()
    y_soft[mask] = y_true
    return lp_propagate(y_soft, edge_index, num_layers, alpha)


# --------------------------------------------------------------------------------------
# a10 PTA graph ops restated in torch (itexperiments.py:671-719, pta.py:79-84).
# Orientation: adjacency ROW = edge_index[0]; duplicates summed; A + I doubles an
# existing loop; degree = row sums.
# --------------------------------------------------------------------------------------


def pta_norm_adj(edge_index: Tensor, num_nodes: int, dtype=torch.float32):
    """Returns (row, col, val) COO triplets of D^-1/2 (A+I) D^-1/2 with A[row=src, col=dst]
    (uncoalesced, like the scipy COO sum keeps duplicates until tocoo()).  Values are
    computed in float64 then cast, as scipy does (itexperiments.py:677-691)."""
    N = num_nodes
    src, dst = edge_index[0], edge_index[1]
    loop = torch.arange(N, dtype=src.dtype)
    r = torch.cat([src, loop])
    c = torch.cat([dst, loop])
    v = torch.ones(r.numel(), dtype=torch.float64)
    rowsum = torch.zeros(N, dtype=torch.float64).index_add_(0, r, v)
    rinv = rowsum.pow(-0.5)
    rinv[torch.isinf(rinv)] = 0.0
    val = (rinv[r] * v * rinv[c]).to(dtype)
    return r, c, val


def pta_spmm(r: Tensor, c: Tensor, val: Tensor, y: Tensor) -> Tensor:
    """torch.matmul(sparse_coo(adj), y): out[r] += val * y[c]."""
    out = torch.zeros_like(y)
    return out.index_add_(0, r, val.to(y.dtype).view(-1, 1) * y[c])


def pta_label_propagation(edge_index: Tensor, num_nodes: int, labels: Tensor, idx: Tensor,
                          K: int, alpha: float) -> Tensor:
    """itexperiments.py:698-719 with the per-node Python loops vectorised."""
    C = int(labels.max()) + 1
    r, c, val = pta_norm_adj(edge_index, num_nodes)
    y0 = torch.zeros(labels.shape[0], C)
    y0[idx, labels[idx]] = 1.0
    lab = labels.clone()
    lab[lab < 0] = 0
    onehot = torch.nn.functional.one_hot(lab).to(y0.dtype)
    y = y0
    for _ in range(K):
        y = pta_spmm(r, c, val, y)
        y[idx] = onehot[idx]
        y = (1 - alpha) * y + alpha * y0
    return y


def pta_inference(h: Tensor, edge_index: Tensor, num_nodes: int, K: int, alpha: float) -> Tensor:
    """pta.py:79-84."""
    r, c, val = pta_norm_adj(edge_index, num_nodes, h.dtype)
    y0 = torch.softmax(h, dim=-1)
    y = y0
    for _ in range(K):
        y = (1 - alpha) * pta_spmm(r, c, val, y) + alpha * y0
    return y


def glorot_(t: Tensor) -> Tensor:
    """PyG inits.glorot: uniform(-a, a), a = sqrt(6 / (size(-2) + size(-1)))."""
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        return t.uniform_(-a, a)
