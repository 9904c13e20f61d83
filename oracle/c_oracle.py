"""ctypes binding + build of the plain-C oracle (oracle/csrc/rgb_oracle.c) -- TEST INFRASTRUCTURE ONLY.

    python -m oracle.c_oracle          # gcc -O2 -ffp-contract=off -shared -fPIC -> oracle/liboracle_c.so

Takes and returns CPU torch tensors (int64 indices, float32 features).  Used by tests/test_oracle_c.py to pin
the scalar-loop statement against the torch oracle and the reference's golden vectors, bit for bit."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "rgb_oracle.c")
LIB = os.path.join(HERE, "liboracle_c.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        cmd = [os.environ.get("CC", "gcc"), "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c11", "-shared", "-fPIC",
               SRC, "-o", LIB, "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"gcc failed for {SRC}:\n{r.stdout}\n{r.stderr}")
    return LIB


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        i64p, f32p = C.POINTER(C.c_int64), C.POINTER(C.c_float)
        L.orc_edit_loops.restype = C.c_int64
        L.orc_edit_loops.argtypes = [i64p, i64p, C.c_int64, C.c_int64, C.c_int, i64p, i64p]
        L.orc_csr_build.restype = None
        L.orc_csr_build.argtypes = [i64p, i64p, C.c_int64, C.c_int64, i64p, i64p, i64p]
        L.orc_gcn_norm_weights.restype = None
        L.orc_gcn_norm_weights.argtypes = [i64p, i64p, C.c_int64, C.c_int64, f32p, f32p]
        L.orc_propagate_add.restype = None
        L.orc_propagate_add.argtypes = [i64p, i64p, f32p, C.c_int64, f32p, C.c_int64, C.c_int, f32p]
        L.orc_propagate_mean.restype = None
        L.orc_propagate_mean.argtypes = [i64p, i64p, C.c_int64, f32p, C.c_int64, C.c_int, f32p]
        L.orc_appnp.restype = None
        L.orc_appnp.argtypes = [i64p, i64p, f32p, C.c_int64, f32p, C.c_int64, C.c_int, C.c_int, C.c_double, f32p, f32p]
        L.orc_gat_aggregate.restype = None
        L.orc_gat_aggregate.argtypes = [i64p, i64p, C.c_int64, f32p, f32p, f32p, C.c_int64, C.c_int, C.c_int, C.c_double,
                                        f32p, f32p]
        L.orc_version.restype = C.c_int
        _lib = L
    return _lib


def _i(t: torch.Tensor):
    assert t.dtype == torch.int64 and t.is_contiguous() and not t.is_cuda
    return C.cast(t.data_ptr(), C.POINTER(C.c_int64))


def _f(t):
    if t is None:
        return None
    assert t.dtype == torch.float32 and t.is_contiguous() and not t.is_cuda
    return C.cast(t.data_ptr(), C.POINTER(C.c_float))


def edit_loops(edge_index: torch.Tensor, num_nodes: int, loop_mode: int) -> torch.Tensor:
    src, dst = edge_index[0].contiguous(), edge_index[1].contiguous()
    E = src.numel()
    out = torch.empty((2, E + num_nodes), dtype=torch.int64)
    n = lib().orc_edit_loops(_i(src), _i(dst), E, num_nodes, loop_mode, _i(out[0]), _i(out[1]))
    if n < 0:
        raise RuntimeError("node id outside [0, N)")
    return out[:, :n].contiguous()


def csr_build(edge_index: torch.Tensor, num_nodes: int, by: str = "dst"):
    row, col = edge_index[0].contiguous(), edge_index[1].contiguous()
    key, other = (col, row) if by == "dst" else (row, col)
    nnz = key.numel()
    rowptr = torch.empty(num_nodes + 1, dtype=torch.int64)
    c, eid = torch.empty(nnz, dtype=torch.int64), torch.empty(nnz, dtype=torch.int64)
    lib().orc_csr_build(_i(key), _i(other), nnz, num_nodes, _i(rowptr), _i(c), _i(eid))
    return rowptr, c, eid


def gcn_norm_weights(edge_index: torch.Tensor, num_nodes: int):
    """(dinv float32 [N], w float32 [nnz]) for an ALREADY edited edge list with unit weights."""
    row, col = edge_index[0].contiguous(), edge_index[1].contiguous()
    dinv = torch.empty(num_nodes, dtype=torch.float32)
    w = torch.empty(row.numel(), dtype=torch.float32)
    lib().orc_gcn_norm_weights(_i(row), _i(col), row.numel(), num_nodes, _f(dinv), _f(w))
    return dinv, w


def propagate(edge_index: torch.Tensor, x: torch.Tensor, w=None, aggr: str = "add") -> torch.Tensor:
    row, col = edge_index[0].contiguous(), edge_index[1].contiguous()
    x = x.contiguous()
    N, F = x.shape
    out = torch.empty((N, F), dtype=torch.float32)
    if aggr == "mean":
        assert w is None
        lib().orc_propagate_mean(_i(row), _i(col), row.numel(), _f(x), N, F, _f(out))
    else:
        lib().orc_propagate_add(_i(row), _i(col), _f(None if w is None else w.contiguous()), row.numel(), _f(x), N, F, _f(out))
    return out


def appnp(edge_index: torch.Tensor, w: torch.Tensor, h: torch.Tensor, K: int, alpha: float) -> torch.Tensor:
    row, col = edge_index[0].contiguous(), edge_index[1].contiguous()
    h = h.contiguous()
    N, F = h.shape
    tmp, z = torch.empty_like(h), torch.empty_like(h)
    lib().orc_appnp(_i(row), _i(col), _f(w.contiguous()), row.numel(), _f(h), N, F, K, float(alpha), _f(tmp), _f(z))
    return z


def gat_aggregate(edge_index: torch.Tensor, xp: torch.Tensor, a_src: torch.Tensor, a_dst: torch.Tensor, slope: float = 0.2):
    """edge_index: the ALREADY edited list (remove_then_add).  xp [N,H,C] -> (out [N,H,C], alpha [nnz,H])."""
    row, col = edge_index[0].contiguous(), edge_index[1].contiguous()
    xp, a_src, a_dst = xp.contiguous(), a_src.contiguous(), a_dst.contiguous()
    N, H, Cc = xp.shape
    alpha = torch.empty((row.numel(), H), dtype=torch.float32)
    out = torch.empty_like(xp)
    lib().orc_gat_aggregate(_i(row), _i(col), row.numel(), _f(xp), _f(a_src), _f(a_dst), N, H, Cc, float(slope),
                            _f(alpha), _f(out))
    return out, alpha


if __name__ == "__main__":
    print("built", build(force=True))
