"""torch.autograd.Function drop-ins over the C ABI (include/rgbmp.h).

Every function here enqueues hand-written sm_100a kernels from librgbmp.so on the current
CUDA stream.  Nothing falls back to PyTorch arithmetic for the aggregation itself; PyTorch only
provides the tensors (device memory), the stream and the autograd graph.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib, memo
from ._lib import Epilogue, check, dtype_code, lib, ptr, stream_of
from .graph import CSR, Graph, NORM_COUNT, NORM_INV_SQRT, oom_retry

RESET_NONE, RESET_BEFORE_TELEPORT, RESET_AFTER_CLAMP = 0, 1, 2
TUNE_OVERRIDE = 0      # experiments (tools/): launch-shape word applied to every SpMM / K-hop call that passes tune=0


# ------------------------------------------------------------------------------------------
# buffer helpers: the vector path wants 16-byte aligned rows (ld % 4 == 0 fp32, % 8 bf16)
# ------------------------------------------------------------------------------------------
def _vec(dtype) -> int:
    return 8 if dtype == torch.bfloat16 else 4


def padded_width(F: int, dtype=torch.float32) -> int:
    v = _vec(dtype)
    return (F + v - 1) // v * v


def as_rows(x: torch.Tensor) -> Tuple[torch.Tensor, int]:
    """Return (tensor usable by the kernels, leading dimension).  No copy when `x` is already
    row-major with an aligned row stride; otherwise one copy into a zero-padded buffer."""
    _lib.require_cuda(x, "x")
    if x.dim() != 2:
        raise RuntimeError("feature matrix must be 2-D [rows, F]")
    N, F = x.shape
    v = _vec(x.dtype)
    if (x.stride(1) == 1 or F == 1) and x.stride(0) % v == 0 and x.stride(0) >= padded_width(F, x.dtype) \
            and x.data_ptr() % 16 == 0:
        return x, x.stride(0)
    Fp = padded_width(F, x.dtype)
    if Fp == F:
        buf = x.contiguous()
        if buf.data_ptr() % 16 == 0:
            return buf, F
    buf = torch.zeros((N, Fp), dtype=x.dtype, device=x.device)
    buf[:, :F].copy_(x)
    return buf[:, :F], Fp


def alloc_rows(N: int, F: int, dtype, device) -> Tuple[torch.Tensor, int]:
    """Fresh [N, F] tensor whose rows are padded(F) elements apart (16-byte aligned rows for the vector kernels).
    NOT a view: the result of an autograd.Function must tolerate the caller's in-place update (`out += x_r`,
    graphsage.py:60), which autograd forbids on a view created inside a custom Function -- so the
    strided tensor is bound to the padded storage directly instead of slicing a [N, padded(F)] tensor."""
    Fp = padded_width(F, dtype)
    buf = torch.empty((N, Fp), dtype=dtype, device=device)
    if Fp == F:
        return buf, Fp
    out = torch.empty(0, dtype=dtype, device=device).set_(buf.untyped_storage(), buf.storage_offset(), (N, F), (Fp, 1))
    return out, Fp


def make_epilogue(*, row_scale=None, row_div=False, a=1.0, b=0.0, T=None, ldt=0, clamp=None,
                  reset_mask=None, reset_val=None, ld_reset=0, reset_when=RESET_NONE,
                  out2_scale=None, acc_in=None, ld_acc=0, skip_empty=False, peers=(), peer_row0=0,
                  ld_peer=0) -> Epilogue:
    e = Epilogue()
    e.row_scale, e.row_div = ptr(row_scale), int(bool(row_div))
    e.reset_mask, e.reset_val, e.ld_reset, e.reset_when = ptr(reset_mask), ptr(reset_val), ld_reset, reset_when
    e.a, e.b, e.T, e.ldt = float(a), float(b), ptr(T), ldt
    if clamp is not None:
        e.clamp, e.lo, e.hi = 1, float(clamp[0]), float(clamp[1])
    e.out2_scale = ptr(out2_scale)
    e.acc_in, e.ld_acc, e.skip_empty = ptr(acc_in), ld_acc, int(bool(skip_empty))
    if len(peers) > 8:
        raise RuntimeError("at most 8 peer buffers (one NVSwitch box)")
    for q, pp in enumerate(peers):          # raw device pointers (ints) of peer-mapped buffers
        e.peer_out[q] = int(pp)
    e.n_peers, e.peer_row0, e.ld_peer = len(peers), int(peer_row0), int(ld_peer)
    return e


def _graph_ref(csr, row_bytes: int, hot: bool = True):
    if hot and hasattr(csr, "hot_ref"):
        return csr.hot_ref(row_bytes)
    return csr.ref


def spmm_raw(csr: CSR, x: torch.Tensor, val: Optional[torch.Tensor] = None, **kw) -> Optional[torch.Tensor]:
    """y[i] = epilogue(sum_k val[k] * x[col[k]]); see _spmm_raw.  A CUDA OOM empties the graph / memo caches and
    retries once (graph.oom_retry)."""
    return oom_retry(lambda: _spmm_raw(csr, x, val, **kw))


def _spmm_raw(csr: CSR, x: torch.Tensor, val: Optional[torch.Tensor] = None, *, ep: Optional[Epilogue] = None,
              keep=(), tune: int = 0, out: Optional[torch.Tensor] = None, hot: bool = True,
              store_local: bool = True) -> Optional[torch.Tensor]:
    """`keep` holds tensors referenced by `ep`.
    hot=True lets feature matrices beyond the L2 budget use the hot-tagged column ids (graph.CSR.hot_ref)."""
    tune = tune or TUNE_OVERRIDE
    xb, ldx = as_rows(x)
    F = x.size(1)
    if x.size(0) < csr.n_cols:
        raise RuntimeError(f"x has {x.size(0)} rows but the graph addresses {csr.n_cols}")
    if not store_local:                       # every store goes through ep.peer_out / ep.Y2
        y, ldy = None, 0
    elif out is None:
        y, ldy = alloc_rows(csr.n_rows, F, x.dtype, x.device)
    else:
        y, ldy = out, out.stride(0)
    ws = csr.spmm_workspace(F)
    dev = x.device
    check(lib().rgbmp_spmm(_graph_ref(csr, ldx * xb.element_size(), hot), ptr(val), ptr(xb), ldx, ptr(y), ldy, F, dtype_code(x),
                           C.byref(ep) if ep is not None else None, tune, ptr(ws),
                           0 if ws is None else ws.numel(), dev.index, stream_of(dev)), "spmm")
    return y


def row_scale(x: torch.Tensor, scale: torch.Tensor, divide: bool = False) -> torch.Tensor:
    xb, ldx = as_rows(x)
    y, ldy = alloc_rows(x.size(0), x.size(1), x.dtype, x.device)
    dev = x.device
    check(lib().rgbmp_row_scale(ptr(xb), ldx, ptr(scale), int(divide), ptr(y), ldy, x.size(0), x.size(1),
                                dtype_code(x), dev.index, stream_of(dev)), "row_scale")
    return y


def khop_raw(csr: CSR, x0: torch.Tensor, K: int, **kw):
    """K fused hops.  Returns the final iterate, or (final, [K, N, F] hop outputs) when hops=True."""
    return oom_retry(lambda: _khop_raw(csr, x0, K, **kw))


def _khop_raw(csr: CSR, x0: torch.Tensor, K: int, *, val=None, ep: Optional[Epilogue] = None, keep=(),
              hops: bool = False, tune: int = 0, hot: bool = True):
    tune = tune or TUNE_OVERRIDE
    xb, ldx = as_rows(x0)
    N, F = x0.shape
    dev, dt = x0.device, x0.dtype
    Fp = padded_width(F, dt)
    out, ldo = alloc_rows(N, F, dt, dev)
    ping = pong = None
    if K > 1:
        ping = torch.empty((N, Fp), dtype=dt, device=dev)
        pong = torch.empty((N, Fp), dtype=dt, device=dev)
    hop_buf = torch.empty((K, N, Fp), dtype=dt, device=dev) if hops else None
    ws = csr.spmm_workspace(F)
    check(lib().rgbmp_khop(_graph_ref(csr, Fp * xb.element_size(), hot), ptr(val), ptr(xb), ldx, ptr(ping), ptr(pong), Fp, ptr(out), ldo,
                           ptr(hop_buf), Fp, N * Fp, F, dtype_code(x0), K,
                           C.byref(ep) if ep is not None else None, tune, ptr(ws),
                           0 if ws is None else ws.numel(), dev.index, stream_of(dev)), "khop")
    if hops:
        return out, hop_buf[:, :, :F]
    return out


# ------------------------------------------------------------------------------------------
# single-hop propagate  (MessagePassing.propagate / GCNConv / SAGEConv / GINConv ...)
# ------------------------------------------------------------------------------------------
class _Propagate(torch.autograd.Function):
    """kind: 'sum' (aggr='add', message=x_j), 'mean' (aggr='mean'), 'gcn' (symmetric norm weights).
    Backward = the same kernel on the transpose CSR (linear operator, nothing but the graph saved)."""

    @staticmethod
    def forward(ctx, x, graph: Graph, kind: str):
        ctx.graph, ctx.kind = graph, kind
        csr = graph.fwd
        if kind == "sum":
            return spmm_raw(csr, x)
        if kind == "mean":
            cnt = csr.norm(NORM_COUNT)
            return spmm_raw(csr, x, ep=make_epilogue(row_scale=cnt, row_div=True), keep=(cnt,))
        if kind == "gcn":
            return spmm_raw(csr, x, graph.gcn_val(False))
        raise ValueError(kind)

    @staticmethod
    def backward(ctx, dy):
        g, kind = ctx.graph, ctx.kind
        if kind == "sum":
            return spmm_raw(g.bwd, dy), None, None
        if kind == "mean":
            return spmm_raw(g.bwd, row_scale(dy, g.fwd.norm(NORM_COUNT), divide=True)), None, None
        return spmm_raw(g.bwd, dy, g.gcn_val(True)), None, None


def _work(graph: Graph, x: torch.Tensor) -> float:
    return float(graph.nnz) * float(x.size(-1) if x.dim() > 1 else 1)


def propagate(x: torch.Tensor, graph: Graph, kind: str = "sum") -> torch.Tensor:
    # memo.cached: the second of the caller's two identical eval forwards reuses the first
    return memo.cached(graph, ("propagate", kind), (x,), _work(graph, x), lambda: _Propagate.apply(x, graph, kind))


class _PropagateWeighted(torch.autograd.Function):
    """y[i] = sum_{e: dst=i} w[e] * x[src[e]] with per-edge weights given in EDGE order [nnz]
    (Prop.message, dagnn.py:57-59; generic edge_weight).  dw[e] = <dy[dst], x[src]> (SDDMM)."""

    @staticmethod
    def forward(ctx, x, w, graph: Graph):
        ctx.graph = graph
        val = graph.to_csr_order(w, transpose=False)
        ctx.save_for_backward(x if ctx.needs_input_grad[1] else None, w)
        return spmm_raw(graph.fwd, x, val)

    @staticmethod
    def backward(ctx, dy):
        g = ctx.graph
        x, w = ctx.saved_tensors
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = spmm_raw(g.bwd, dy, g.to_csr_order(w, transpose=True))
        if ctx.needs_input_grad[1]:
            dw = g.to_edge_order(sddmm(g.fwd, dy, x, 1, dy.size(1))).view_as(w)
        return dx, dw, None


def propagate_weighted(x, w, graph: Graph):
    return memo.cached(graph, ("propagate_weighted",), (x, w), _work(graph, x),
                       lambda: _PropagateWeighted.apply(x, w, graph))


# ------------------------------------------------------------------------------------------
# K-hop families
# ------------------------------------------------------------------------------------------
FOLD_KHOP = True     # K-hop families apply D^-1/2 (A+I) D^-1/2 as row scalings around an unweighted sum (no per-edge
                     # weight stream, 4 B/edge less and adds instead of FMAs); False: per-edge gcn_norm weights as PyG multiplies them


def _khop_gcn(csr: CSR, graph: Graph, x0b: torch.Tensor, K: int, transpose: bool, fold: bool, **epkw):
    """K hops of z <- epilogue(A_hat z) on the chosen orientation.  fold: the iterate travels pre-scaled
    (u = D^-1/2 z, what the next hop gathers) and the epilogue scales the row sum by D^-1/2 again; the
    in-degree vector serves both orientations (A_hat^T = D^-1/2 A^T D^-1/2 with the same D)."""
    if fold:
        d = graph.dinv()
        return khop_raw(csr, row_scale(x0b, d), K, ep=make_epilogue(row_scale=d, out2_scale=d, **epkw))
    return khop_raw(csr, x0b, K, val=graph.gcn_val(transpose), ep=make_epilogue(**epkw))


def _staged(x: torch.Tensor, d: torch.Tensor):
    """(z0 padded to an aligned row stride, u0 = D^-1/2 z0) from an UNALIGNED fp32 input in ONE kernel (instead of a zero
    fill, a copy and a row scale): the case of every class count that is not a multiple of four (47, 10, ...)."""
    N, F = x.shape
    Fp = padded_width(F, x.dtype)
    z0 = torch.empty((N, Fp), dtype=x.dtype, device=x.device)
    u0 = torch.empty((N, Fp), dtype=x.dtype, device=x.device)
    xc = x if x.stride(1) == 1 else x.contiguous()
    dev = x.device
    check(lib().rgbmp_stage_rows(ptr(xc), xc.stride(0), ptr(d), ptr(z0), ptr(u0), Fp, N, F, dev.index, stream_of(dev)), "stage_rows")
    return z0[:, :F], u0[:, :F], Fp


def _appnp_khop(csr: CSR, graph: Graph, x: torch.Tensor, K: int, alpha: float, transpose: bool, fold: bool):
    _lib.require_cuda(x, "x")
    if fold and x.dim() == 2 and x.dtype == torch.float32 and x.size(1) % 4 != 0 and x.size(0) > 0:
        d = graph.dinv()
        z0, u0, ld = _staged(x, d)
        ep = make_epilogue(row_scale=d, out2_scale=d, a=1.0 - alpha, b=alpha, T=z0, ldt=ld)
        return khop_raw(csr, u0, K, ep=ep, keep=(z0, d))
    xb, ldx = as_rows(x)
    return _khop_gcn(csr, graph, xb, K, transpose, fold, a=1.0 - alpha, b=alpha, T=xb, ldt=ldx)


class _APPNP(torch.autograd.Function):
    """z = x; K x { z = (1-alpha) * A_hat z + alpha * x }  (A8; appnp_stack.py:22).
    The map is linear, z_K = M x with M = alpha*sum_{k<K} B^k + B^K, B=(1-alpha)A_hat, so the
    backward is the same recursion on the transpose graph -- no hop activations are kept."""

    @staticmethod
    def forward(ctx, x, graph: Graph, K: int, alpha: float, fold: bool):
        ctx.graph, ctx.K, ctx.alpha, ctx.fold = graph, K, alpha, fold
        return _appnp_khop(graph.fwd, graph, x, K, alpha, False, fold)

    @staticmethod
    def backward(ctx, dz):
        g = ctx.graph
        return _appnp_khop(g.bwd, g, dz, ctx.K, ctx.alpha, True, ctx.fold), None, None, None, None


def appnp(x, graph: Graph, K: int, alpha: float, fold: Optional[bool] = None):
    if K == 0:
        return x
    K, alpha, fold = int(K), float(alpha), bool(FOLD_KHOP if fold is None else fold)
    return memo.cached(graph, ("appnp", K, alpha, fold), (x,), K * _work(graph, x),
                       lambda: _APPNP.apply(x, graph, K, alpha, fold))


class _PowerHops(torch.autograd.Function):
    """x <- A_hat^K x (A9 SGConv, sgc.py:9-10)."""

    @staticmethod
    def forward(ctx, x, graph: Graph, K: int, fold: bool):
        ctx.graph, ctx.K, ctx.fold = graph, K, fold
        return _khop_gcn(graph.fwd, graph, as_rows(x)[0], K, False, fold)

    @staticmethod
    def backward(ctx, dy):
        g = ctx.graph
        return _khop_gcn(g.bwd, g, as_rows(dy)[0], ctx.K, True, ctx.fold), None, None, None


def gcn_power(x, graph: Graph, K: int, fold: Optional[bool] = None):
    if K == 0:
        return x
    fold = bool(FOLD_KHOP if fold is None else fold)
    return memo.cached(graph, ("gcn_power", int(K), fold), (x,), K * _work(graph, x),
                       lambda: _PowerHops.apply(x, graph, int(K), fold))


class _DagnnHops(torch.autograd.Function):
    """[N, K+1, C] stack (x, A_hat x, ..., A_hat^K x) in one K-hop call that keeps every hop (dagnn.py:43-49).
    Backward: g_K = dS_K; g_k = dS_k + A_hat^T g_{k+1} -- K transposed SpMMs with the add in the epilogue."""

    @staticmethod
    def forward(ctx, x, graph: Graph, K: int):
        ctx.graph, ctx.K = graph, K
        _, hops = khop_raw(graph.fwd, x, K, val=graph.gcn_val(False), hops=True)
        return torch.cat([x.detach().unsqueeze(0), hops], dim=0).permute(1, 0, 2)

    @staticmethod
    def backward(ctx, dS):
        g, K = ctx.graph, ctx.K
        val = g.gcn_val(True)
        acc = dS[:, K, :].contiguous()
        for k in range(K - 1, -1, -1):
            t, ldt = as_rows(dS[:, k, :])
            acc = spmm_raw(g.bwd, acc, val, ep=make_epilogue(a=1.0, b=1.0, T=t, ldt=ldt), keep=(t,))
        return acc, None, None


def dagnn_hops(x: torch.Tensor, graph: Graph, K: int) -> torch.Tensor:
    """[N, K+1, C] stack of x and its K propagated versions (dagnn.py:43-49), differentiable."""
    return _DagnnHops.apply(x, graph, int(K))


def label_propagation(graph: Graph, out0: torch.Tensor, num_layers: int, alpha: float, *,
                      clamp=(0.0, 1.0), reset_mask=None, reset_val=None, fold: Optional[bool] = None) -> torch.Tensor:
    """A15 LP core: res=(1-a)*out0; L x { out = a * A_hat out + res; post_step }.
    post_step = clamp(lo,hi) (default / autoscale) or `out[mask] = reset_val[mask]` (fixed scale).
    `graph` must be built with LOOP_NONE (gcn_norm(add_self_loops=False))."""
    if num_layers == 0:
        return out0
    xb, ldx = as_rows(out0)
    kw = {}
    keep = [xb]
    if reset_mask is not None:
        m = reset_mask.to(torch.uint8).contiguous()
        rv, ldr = as_rows(reset_val)
        kw = dict(reset_mask=m, reset_val=rv, ld_reset=ldr, reset_when=RESET_AFTER_CLAMP)
        keep += [m, rv]
        clamp = None
    return _khop_gcn(graph.fwd, graph, xb, num_layers, False, bool(FOLD_KHOP if fold is None else fold),
                     a=alpha, b=1.0 - alpha, T=xb, ldt=ldx, clamp=clamp, **kw)


# ------------------------------------------------------------------------------------------
# PTA graph ops (itexperiments.py:671-719, pta.py:79-84).  The PTA adjacency aggregates at
# row = edge_index[0], i.e. it is the gcn propagation of the REVERSED graph, with A+I doubling
# an existing self loop -> build the graph with LOOP_ADD on the flipped
# edge_index: duplicates and pre-existing loops then count exactly like the scipy COO sum.
# ------------------------------------------------------------------------------------------
def pta_graph(edge_index: torch.Tensor, num_nodes: int) -> Graph:
    from .graph import LOOP_ADD, get_graph
    return get_graph(edge_index, num_nodes, LOOP_ADD, reverse=True)


def pta_inference(h: torch.Tensor, graph: Graph, K: int, alpha: float) -> torch.Tensor:
    y0 = torch.softmax(h, dim=-1)
    xb, ldx = as_rows(y0)
    return _khop_gcn(graph.fwd, graph, xb, K, False, FOLD_KHOP, a=1.0 - alpha, b=alpha, T=xb, ldt=ldx)


def pta_label_propagation(graph: Graph, labels: torch.Tensor, idx: torch.Tensor, K: int, alpha: float):
    N = labels.size(0)
    Cn = int(labels.max().item()) + 1
    y0 = torch.zeros((N, Cn), dtype=torch.float32, device=labels.device)
    y0[idx, labels[idx]] = 1.0
    mask = torch.zeros(N, dtype=torch.uint8, device=labels.device)
    mask[idx] = 1
    xb, ldx = as_rows(y0)
    # y <- A y ; y[idx] <- onehot[idx] (= y0[idx]) ; y <- (1-a) y + a y0
    return _khop_gcn(graph.fwd, graph, xb, K, False, FOLD_KHOP, a=1.0 - alpha, b=alpha, T=xb, ldt=ldx, reset_mask=mask,
                     reset_val=xb, ld_reset=ldx, reset_when=RESET_BEFORE_TELEPORT)


# ------------------------------------------------------------------------------------------
# attention / edge-score kernels
# ------------------------------------------------------------------------------------------
def sddmm(csr: CSR, A: torch.Tensor, B: torch.Tensor, H: int, Cc: int) -> torch.Tensor:
    """out[k,h] = <A[row_i,h,:], B[col[k],h,:]> in CSR order."""
    Ab, lda = as_rows(A)
    Bb, ldb = as_rows(B)
    out = torch.empty((csr.nnz, H), dtype=torch.float32, device=A.device)
    dev = A.device
    check(lib().rgbmp_sddmm(csr.ref, ptr(Ab), lda, ptr(Bb), ldb, H, Cc, ptr(out), dev.index, stream_of(dev)), "sddmm")
    return out


def _edge_call(fn_name, csr: CSR, *args):
    dev = csr.device
    check(getattr(lib(), fn_name)(csr.ref, *args, dev.index, stream_of(dev)), fn_name)


def spmm_heads_raw(csr: CSR, w: torch.Tensor, X: torch.Tensor, H: int, Cc: int) -> torch.Tensor:
    Xb, ldx = as_rows(X)
    out, ldo = alloc_rows(csr.n_rows, H * Cc, torch.float32, X.device)
    _edge_call("rgbmp_spmm_heads", csr, ptr(w.contiguous()), ptr(Xb), ldx, H, Cc, ptr(out), ldo)
    return out


def seg_sum_raw(csr: CSR, vals: torch.Tensor, H: int) -> torch.Tensor:
    out = torch.empty((csr.n_rows, H), dtype=torch.float32, device=vals.device)
    _edge_call("rgbmp_seg_sum", csr, ptr(vals.contiguous()), H, ptr(out))
    return out


class _SDDMM(torch.autograd.Function):
    """out[k,h] = <A[i,h,:], B[j,h,:]> for every edge j->i, forward-CSR order (SuperGAT MX logits)."""

    @staticmethod
    def forward(ctx, A, B, graph: Graph, H: int, Cc: int):
        ctx.graph, ctx.H, ctx.Cc = graph, H, Cc
        ctx.save_for_backward(A, B)
        return sddmm(graph.fwd, A, B, H, Cc)

    @staticmethod
    def backward(ctx, g):
        A, B = ctx.saved_tensors
        gr, H, Cc = ctx.graph, ctx.H, ctx.Cc
        g = g.contiguous()
        dA = spmm_heads_raw(gr.fwd, g, B, H, Cc) if ctx.needs_input_grad[0] else None
        dB = spmm_heads_raw(gr.bwd, gr.fwd_to_bwd(g), A, H, Cc) if ctx.needs_input_grad[1] else None
        return dA, dB, None, None, None


class _UAddV(torch.autograd.Function):
    """out[k,h] = u[j,h] + v[i,h] for every edge j->i (forward-CSR order)."""

    @staticmethod
    def forward(ctx, u, v, graph: Graph):
        ctx.graph = graph
        H = u.size(1)
        out = torch.empty((graph.nnz, H), dtype=torch.float32, device=u.device)
        _edge_call("rgbmp_u_add_v", graph.fwd, ptr(u.contiguous()), ptr(v.contiguous()), H, ptr(out))
        return out

    @staticmethod
    def backward(ctx, g):
        gr = ctx.graph
        g = g.contiguous()
        H = g.size(1)
        du = seg_sum_raw(gr.bwd, gr.fwd_to_bwd(g), H) if ctx.needs_input_grad[0] else None
        dv = seg_sum_raw(gr.fwd, g, H) if ctx.needs_input_grad[1] else None
        return du, dv, None


class _SegSoftmax(torch.autograd.Function):
    """Edge softmax over the in-edges of each target (A11), forward-CSR order."""

    @staticmethod
    def forward(ctx, e, graph: Graph):
        ctx.graph = graph
        e = e.contiguous()
        out = torch.empty_like(e)
        _edge_call("rgbmp_seg_softmax", graph.fwd, ptr(e), e.size(1), ptr(out))
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, g):
        (alpha,) = ctx.saved_tensors
        g = g.contiguous()
        out = torch.empty_like(g)
        _edge_call("rgbmp_seg_softmax_backward", ctx.graph.fwd, ptr(alpha), ptr(g), g.size(1), ptr(out))
        return out, None


class _SpmmHeads(torch.autograd.Function):
    """out[i,h,:] = sum_{j->i} w[k,h] * X[j,h,:], w in forward-CSR order."""

    @staticmethod
    def forward(ctx, w, X, graph: Graph, H: int, Cc: int):
        ctx.graph, ctx.H, ctx.Cc = graph, H, Cc
        ctx.save_for_backward(w, X)
        return spmm_heads_raw(graph.fwd, w, X, H, Cc)

    @staticmethod
    def backward(ctx, dout):
        w, X = ctx.saved_tensors
        gr, H, Cc = ctx.graph, ctx.H, ctx.Cc
        dw = sddmm(gr.fwd, dout, X, H, Cc) if ctx.needs_input_grad[0] else None
        dX = spmm_heads_raw(gr.bwd, gr.fwd_to_bwd(w), dout, H, Cc) if ctx.needs_input_grad[1] else None
        return dw, dX, None, None, None


def edge_sddmm(A, B, graph, H, Cc):
    return _SDDMM.apply(A, B, graph, H, Cc)


def edge_u_add_v(u, v, graph):
    return _UAddV.apply(u, v, graph)


def edge_softmax(e, graph):
    return _SegSoftmax.apply(e, graph)


def spmm_heads(w, X, graph, H, Cc):
    return _SpmmHeads.apply(w, X, graph, H, Cc)


ATT_GAT, ATT_MX, ATT_FA = 0, 1, 2


def _head_pad(H: int, Cc: int) -> int:
    """Channels per head the fused kernels run with: C itself for a single head, else the next power of two >= 8
    (a lane's 8 channels must stay inside one head; zero channels change neither dot products nor sums)."""
    if H == 1:
        return Cc
    c = 8
    while c < Cc:
        c <<= 1
    return c


def att_fusable(score: int, H: int, Cc: int) -> bool:
    return bool(lib().rgbmp_att_supported(score, H, _head_pad(H, Cc)))


def gat_fusable(H: int, Cc: int) -> bool:
    return att_fusable(ATT_GAT, H, Cc)


def _rows_zero_padded(x: torch.Tensor) -> Tuple[torch.Tensor, int]:
    """as_rows, but the padding columns up to roundup(F,4) are guaranteed ZERO (the H == 1 kernels take dot
    products over whole float4 vectors)."""
    F = x.size(1)
    if F % 4 == 0:
        return as_rows(x)
    Fp = padded_width(F, x.dtype)
    buf = torch.zeros((x.size(0), Fp), dtype=x.dtype, device=x.device)
    buf[:, :F].copy_(x)
    return buf[:, :F], Fp


class _Att(torch.autograd.Function):
    """Fused edge-score attention + aggregate (csrc/att.cu): GATConv (A10/A11; gat.py:18-21), SuperGATConv-MX
    (A12; supergat.py:15-21) and FAConv (A13; fagcn.py:15,31).  Nothing edge-sized is allocated: the forward saves
    per-node statistics (max, sum and -- in training mode -- the second aggregate that yields the gradient of the
    target-side score term), the backward recomputes the edge weights.  Deterministic (no atomics)."""

    @staticmethod
    def forward(ctx, x, a_nbr, a_own, graph: Graph, score: int, H: int, Cc: int, slope: float, drop_csr):
        _lib.require_cuda(x, "x")
        xb, ldx = _rows_zero_padded(x)
        a_n, a_o = a_nbr.contiguous().view(-1, H), a_own.contiguous().view(-1, H)
        N, dev = graph.N, x.device
        train = any(ctx.needs_input_grad[:3])
        out, ldo = alloc_rows(N, H * Cc, torch.float32, dev)
        rmax = rsum = rowq = out2 = dinv = None
        ldo2 = 0
        if score != ATT_FA:
            rmax = torch.empty((N, H), dtype=torch.float32, device=dev)
            rsum = torch.empty((N, H), dtype=torch.float32, device=dev)
        else:
            dinv = graph.dinv()
        if train and score != ATT_MX:
            out2, ldo2 = alloc_rows(N, H * Cc, torch.float32, dev)
            if score == ATT_GAT:
                rowq = torch.empty((N, H), dtype=torch.float32, device=dev)
        L = lib()
        wsb = L.rgbmp_att_forward_workspace_bytes(graph.fwd.ref, H, Cc)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        check(L.rgbmp_att_forward(graph.fwd.ref, score, ptr(xb), ldx, ptr(a_n), ptr(a_o), ptr(dinv), H, Cc, slope,
                                  ptr(drop_csr), ptr(out), ldo, ptr(rmax), ptr(rsum), ptr(out2), ldo2, ptr(rowq),
                                  ptr(ws), wsb, dev.index, stream_of(dev)), "att_forward")
        ctx.graph, ctx.score, ctx.H, ctx.Cc, ctx.slope = graph, score, H, Cc, slope
        ctx.save_for_backward(xb, a_n, a_o, rmax, rsum, rowq, out, out2, drop_csr)
        return out

    @staticmethod
    def backward(ctx, dout):
        xb, a_n, a_o, rmax, rsum, rowq, out, out2, drop = ctx.saved_tensors
        g, score, H, Cc = ctx.graph, ctx.score, ctx.H, ctx.Cc
        dev, N = dout.device, g.N
        db, ldd = _rows_zero_padded(dout)
        dx, lddx = alloc_rows(N, H * Cc, torch.float32, dev)
        dxf, lddxf = (alloc_rows(N, H * Cc, torch.float32, dev) if score == ATT_MX else (None, 0))
        da_n = torch.empty((N, H), dtype=torch.float32, device=dev)
        da_o = torch.empty((N, H), dtype=torch.float32, device=dev)
        tpos = g.tpos() if drop is not None else None
        dinv = g.dinv() if score == ATT_FA else None
        L = lib()
        wsb = L.rgbmp_att_backward_workspace_bytes(g.fwd.ref, g.bwd.ref, H, Cc)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        check(L.rgbmp_att_backward(g.fwd.ref, g.bwd.ref, score, ptr(xb), xb.stride(0), ptr(a_n), ptr(a_o), ptr(dinv), H, Cc,
                                   ctx.slope, ptr(drop), ptr(tpos), ptr(rmax), ptr(rsum), ptr(rowq), ptr(out),
                                   out.stride(0), ptr(out2), 0 if out2 is None else out2.stride(0), ptr(db), ldd,
                                   ptr(dx), lddx, ptr(dxf), lddxf, ptr(da_n), ptr(da_o), ptr(ws), wsb, dev.index,
                                   stream_of(dev)), "att_backward")
        return dx, da_n, da_o, None, None, None, None, None, None


def _att(score: int, name: str, x, a_nbr, a_own, graph: Graph, H: int, Cc: int, slope: float, drop_edge):
    """Pads heads to the kernel's channel count (torch ops on node-sized tensors, autograd-transparent), memoises
    the no-grad call, applies the fused op."""
    Cp = _head_pad(H, Cc)
    N = x.size(0)
    xk = x if Cp == Cc else torch.nn.functional.pad(x.view(N, H, Cc), (0, Cp - Cc)).reshape(N, H * Cp)
    drop_csr = None if drop_edge is None else graph.to_csr_order(drop_edge)
    if drop_csr is None:
        out = memo.cached(graph, (name, H, Cc, float(slope)), (xk, a_nbr, a_own), _work(graph, xk),
                          lambda: _Att.apply(xk, a_nbr, a_own, graph, score, H, Cp, float(slope), None))
    else:
        out = _Att.apply(xk, a_nbr, a_own, graph, score, H, Cp, float(slope), drop_csr)
    return out if Cp == Cc else out.view(N, H, Cp)[:, :, :Cc].reshape(N, H * Cc)


def gat(xp, a_src, a_dst, graph: Graph, H: int, Cc: int, slope: float = 0.2, drop_edge: Optional[torch.Tensor] = None):
    """xp [N,H*C], a_src/a_dst [N,H] -> [N,H*C].  drop_edge: optional [nnz,H] keep-mask/(1-p) in EDGE order."""
    if att_fusable(ATT_GAT, H, Cc):
        return _att(ATT_GAT, "gat", xp, a_src, a_dst, graph, H, Cc, slope, drop_edge)
    e = torch.nn.functional.leaky_relu(edge_u_add_v(a_src, a_dst, graph), slope)
    alpha = edge_softmax(e, graph)
    if drop_edge is not None:
        alpha = alpha * graph.to_csr_order(drop_edge)
    return spmm_heads(alpha, xp, graph, H, Cc)


def supergat_mx(xp, a_l, a_r, graph: Graph, H: int, Cc: int, slope: float = 0.2,
                drop_edge: Optional[torch.Tensor] = None):
    """SuperGATConv MX attention + aggregate (A12): alpha = softmax(leaky_relu((a_l[j] + a_r[i]) * sigmoid(<x_i, x_j>)))."""
    if att_fusable(ATT_MX, H, Cc):
        return _att(ATT_MX, "supergat_mx", xp, a_l, a_r, graph, H, Cc, slope, drop_edge)
    logits = edge_sddmm(xp, xp, graph, H, Cc)
    alpha = torch.nn.functional.leaky_relu(edge_u_add_v(a_l, a_r, graph) * logits.sigmoid(), slope)
    alpha = edge_softmax(alpha, graph)
    if drop_edge is not None:
        alpha = alpha * graph.to_csr_order(drop_edge)
    return spmm_heads(alpha, xp, graph, H, Cc)


def faconv(x, a_l, a_r, graph: Graph, drop_edge: Optional[torch.Tensor] = None):
    """FAConv aggregate (A13): out[i] = sum_j tanh(a_l[j] + a_r[i]) * gcn_norm_ij * mask_ij * x[j]; a_l, a_r [N,1]."""
    Cc = x.size(1)
    if att_fusable(ATT_FA, 1, Cc):
        return _att(ATT_FA, "faconv", x, a_l, a_r, graph, 1, Cc, 0.0, None if drop_edge is None else drop_edge.view(-1, 1))
    c = edge_u_add_v(a_l.view(-1, 1), a_r.view(-1, 1), graph).tanh()
    if drop_edge is not None:
        c = c * graph.to_csr_order(drop_edge).view(-1, 1)
    return spmm_heads(c * graph.gcn_val(False).view(-1, 1), x, graph, 1, Cc)


class _EidView:
    """The forward CSR re-read with col = eid: row i then lists the EDGE ids of its in-edges, so an
    SpMM over a per-edge message matrix [nnz, F] is the segmented reduction of scatter(msg, dst)."""

    def __init__(self, csr: CSR):
        from ._lib import GraphStruct
        self.n_rows, self.n_cols, self.nnz, self.device = csr.n_rows, csr.nnz, csr.nnz, csr.device
        self._csr = csr
        self.struct = GraphStruct(csr.n_rows, csr.nnz, csr.nnz, ptr(csr.rowptr), ptr(csr.eid), csr.chunk,
                                  csr.long_chunk, csr.n_long, csr.n_items, ptr(csr.long_rows),
                                  ptr(csr.long_item_ptr), ptr(csr.item_long), ptr(csr.item_start),
                                  ptr(csr.row_order))
        self.ref = C.byref(self.struct)

    def spmm_workspace(self, F):
        return self._csr.spmm_workspace(F)

    def norm(self, mode):
        return self._csr.norm(mode)


class _SegmentReduce(torch.autograd.Function):
    """out[i] = sum (or mean) of msg[e] over edges e with dst[e] == i  -- the aggregate step of a
    MessagePassing layer whose user-defined `message` produced msg in edge order (A6)."""

    @staticmethod
    def forward(ctx, msg, graph: Graph, mean: bool):
        view = graph._vals.get("eid_view")
        if view is None:
            view = _EidView(graph.fwd)
            graph._vals["eid_view"] = view
        ctx.graph, ctx.mean = graph, mean
        if mean:
            cnt = graph.fwd.norm(NORM_COUNT)
            return spmm_raw(view, msg, ep=make_epilogue(row_scale=cnt, row_div=True), keep=(cnt,))
        return spmm_raw(view, msg)

    @staticmethod
    def backward(ctx, dy):
        g = ctx.graph
        if ctx.mean:
            dy = row_scale(dy, g.fwd.norm(NORM_COUNT), divide=True)
        return dy.index_select(0, g.e_dst.long()), None, None


def segment_reduce(msg: torch.Tensor, graph: Graph, mean: bool = False) -> torch.Tensor:
    return _SegmentReduce.apply(msg, graph, bool(mean))


class HostAppnpPlan:
    """Device scratch for the host-buffer entry point, allocated once and reused across calls."""

    def __init__(self, graph: Graph, F: int):
        N, dev = graph.N, graph.device
        self.graph, self.F, self.ld = graph, F, padded_width(F)
        self.bufs = [torch.empty((N, self.ld), dtype=torch.float32, device=dev) for _ in range(4)]
        self.ws = graph.fwd.spmm_workspace(F)
        self.dinv = graph.dinv()


def appnp_host(graph: Graph, z0_host: torch.Tensor, out_host: torch.Tensor, K: int, alpha: float,
               plan: Optional[HostAppnpPlan] = None) -> torch.Tensor:
    """End-to-end APPNP propagation with HOST buffers (pinned for full PCIe speed): one C-ABI call
    does H2D, the D^-1/2 pre-scale, K fused hops and D2H, and returns when `out_host` is valid."""
    if z0_host.is_cuda or out_host.is_cuda:
        raise RuntimeError("appnp_host takes host tensors")
    if z0_host.dtype != torch.float32 or not z0_host.is_contiguous() or not out_host.is_contiguous():
        raise RuntimeError("appnp_host needs contiguous float32 host tensors")
    N, F = z0_host.shape
    if plan is None:
        plan = HostAppnpPlan(graph, F)
    dev = graph.device
    z0d, ping, pong, outd = plan.bufs
    check(lib().rgbmp_appnp_host(graph.fwd.hot_ref(plan.ld * 4), ptr(plan.dinv), z0_host.data_ptr(), out_host.data_ptr(), F, K,
                                 float(alpha), ptr(z0d), ptr(ping), ptr(pong), ptr(outd), plan.ld, ptr(plan.ws),
                                 0 if plan.ws is None else plan.ws.numel(), dev.index, stream_of(dev)), "appnp_host")
    return out_host
