"""CSR SpMM (sum / mean / gcn-weighted / edge-weighted), forward and backward, against the
oracle's gather -> scale -> scatter_add_ path.  fp32: exact on rows that are not split (same
accumulation order as PyG-CPU), <= 1e-5 norm-wise otherwise.  Through the C ABI."""
import pytest
import torch

from helpers import CASES, relerr
from oracle import pyg_restated as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5          # north_star: within 1e-5 relative (norm-wise) for fp32


def P():
    import rgb_experiment_b200 as P_
    return P_


def oracle_prop(ei, n, mode, x, kind):
    ed = R.edit_loops(ei, n, mode)
    if kind == "sum":
        return R.propagate(ed, x, None, "add", n)
    if kind == "mean":
        return R.propagate(ed, x, None, "mean", n)
    _, w = R.gcn_norm(ei, None, n, add_self_loops=(mode == R.LOOP_ADD_REMAINING), dtype=x.dtype)
    return R.propagate(ed, x, w, "add", n)


@pytest.mark.parametrize("case", ["tiny", "loops_dups", "isolated", "hub", "empty", "single_node", "medium"])
@pytest.mark.parametrize("kind,mode", [("sum", 0), ("mean", 3), ("gcn", 2), ("gcn", 0)])
@pytest.mark.parametrize("F", [1, 7, 47, 64])
def test_propagate_forward_backward(case, kind, mode, F):
    p = P()
    ei, n = CASES[case]()
    torch.manual_seed(F)
    x = torch.randn(n, F)
    dy = torch.randn(n, F)
    xo = x.clone().requires_grad_(True)
    yo = oracle_prop(ei, n, mode, xo, kind)
    yo.backward(dy)
    g = p.Graph(ei.to(DEV), n, mode)
    xg = x.to(DEV).requires_grad_(True)
    yg = p.ops.propagate(xg, g, kind)
    yg.backward(dy.to(DEV))
    assert yg.shape == (n, F)
    assert relerr(yg.detach(), yo.detach()) <= TOL
    assert relerr(xg.grad, xo.grad) <= TOL


@pytest.mark.parametrize("kind,mode", [("sum", 0), ("mean", 3)])
def test_bitwise_fidelity_to_cpu_scatter_order(kind, mode):
    """Rows that are not split accumulate their edges sequentially in stable edge order -- the
    order PyG-on-CPU scatter_add_ uses -- so unweighted fp32 sums / means are bit-identical to the
    oracle, forward and backward.  (Weighted sums use one fused multiply-add per edge, i.e. one
    rounding instead of two, and are held to the 1e-5 bar instead.)"""
    p = P()
    ei, n = CASES["loops_dups"]()
    F = 20
    gen = torch.Generator().manual_seed(11)
    x = torch.randn(n, F, generator=gen)
    dy = torch.randn(n, F, generator=gen)
    xo = x.clone().requires_grad_(True)
    yo = oracle_prop(ei, n, mode, xo, kind)
    yo.backward(dy)
    g = p.Graph(ei.to(DEV), n, mode)
    assert g.fwd.n_long == 0 and g.bwd.n_long == 0
    xg = x.to(DEV).requires_grad_(True)
    yg = p.ops.propagate(xg, g, kind)
    yg.backward(dy.to(DEV))
    assert torch.equal(yg.detach().cpu(), yo.detach())
    assert torch.equal(xg.grad.cpu(), xo.grad)


@pytest.mark.parametrize("F", [3, 40, 41, 100, 128, 256, 602, 1433])
def test_widths_of_the_baseline_configs(F):
    p = P()
    ei, n = CASES["medium"]()
    x = torch.randn(n, F, generator=torch.Generator().manual_seed(F))
    g = p.Graph(ei.to(DEV), n, p.LOOP_ADD_REMAINING)
    y = p.ops.propagate(x.to(DEV), g, "gcn")
    assert relerr(y, oracle_prop(ei, n, 2, x, "gcn")) <= TOL
    x64 = x.double()
    assert relerr(y, oracle_prop(ei, n, 2, x64, "gcn")) <= TOL          # fp64 arbiter


def test_every_launch_shape_gives_the_same_answer():
    p = P()
    ei, n = CASES["medium"]()
    F = 48
    x = torch.randn(n, F, generator=torch.Generator().manual_seed(0))
    g = p.Graph(ei.to(DEV), n, p.LOOP_ADD_REMAINING)
    ref = oracle_prop(ei, n, 2, x, "gcn")
    xd = x.to(DEV)
    outs = []
    for G in (1, 2, 4, 8, 16, 32):
        for V in (1, 2):
            for U in (2, 4, 8, 18, 20):
                y = p.ops.spmm_raw(g.fwd, xd, g.gcn_val(False), tune=G | (V << 8) | (U << 16))
                assert relerr(y, ref) <= TOL, (G, V, U)
                outs.append(y)
    # rows that are not split accumulate in the same order whatever the launch shape
    short = (g.fwd.degree() <= g.fwd.chunk)
    assert int((~short).sum()) > 0
    for y in outs[1:]:
        assert torch.equal(y[short], outs[0][short])
    # shapes that are no longer instantiated are rejected, not silently replaced
    for bad in (1 | (3 << 8) | (4 << 16), 8 | (1 << 8) | (16 << 16), 3 | (1 << 8) | (4 << 16)):
        with pytest.raises(RuntimeError):
            p.ops.spmm_raw(g.fwd, xd, g.gcn_val(False), tune=bad)


def test_unaligned_input_takes_the_scalar_path_and_noncontiguous_views_work():
    p = P()
    ei, n = CASES["loops_dups"]()
    g = p.Graph(ei.to(DEV), n, p.LOOP_NONE)
    big = torch.randn(n, 23, generator=torch.Generator().manual_seed(1))
    x = big[:, 2:13]                                  # row stride 23, offset 2 floats: not 16-byte aligned
    ref = R.propagate(ei, x, None, "add", n)
    y = p.ops.propagate(big.to(DEV)[:, 2:13], g, "sum")
    assert relerr(y, ref) <= TOL
    # direct C-ABI call on the unaligned view exercises the scalar kernels (EPV=1)
    xd = big.to(DEV)[:, 2:13]
    out = torch.empty(n, 11, device=DEV)
    L = p._lib.lib()
    rc = L.rgbmp_spmm(g.fwd.ref, None, xd.data_ptr(), 23, out.data_ptr(), 11, 11, 0, None, 0, None, 0, 0,
                      torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    assert relerr(out, ref) <= TOL


def test_bf16_features_fp32_accumulate():
    p = P()
    ei, n = CASES["medium"]()
    F = 128
    x = torch.randn(n, F, generator=torch.Generator().manual_seed(2)).bfloat16()
    g = p.Graph(ei.to(DEV), n, p.LOOP_ADD_REMAINING)
    y = p.ops.propagate(x.to(DEV), g, "gcn")
    assert y.dtype == torch.bfloat16
    ref = oracle_prop(ei, n, 2, x.float(), "gcn")
    # stated bf16 tolerance: one bf16 rounding of the fp32-accumulated result (2^-8 relative)
    assert relerr(y.float(), ref) <= 2 ** -8


def test_edge_weighted_propagate_and_weight_gradient():
    p = P()
    ei, n = CASES["loops_dups"]()
    F = 12
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(n, F, generator=gen)
    w = torch.rand(ei.size(1), generator=gen)
    dy = torch.randn(n, F, generator=gen)
    xo, wo = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    R.propagate(ei, xo, wo, "add", n).backward(dy)
    g = p.Graph(ei.to(DEV), n, p.LOOP_NONE)
    xg, wg = x.to(DEV).requires_grad_(True), w.to(DEV).requires_grad_(True)
    y = p.ops.propagate_weighted(xg, wg, g)
    y.backward(dy.to(DEV))
    assert relerr(y.detach(), R.propagate(ei, x, w, "add", n)) <= TOL
    assert relerr(xg.grad, xo.grad) <= TOL
    assert relerr(wg.grad, wo.grad) <= TOL


def test_segment_reduce_generic_message_path():
    p = P()
    ei, n = CASES["hub"]()
    gen = torch.Generator().manual_seed(4)
    msg = torch.randn(ei.size(1), 9, generator=gen)
    g = p.Graph(ei.to(DEV), n, p.LOOP_NONE)
    for mean in (False, True):
        mo = msg.clone().requires_grad_(True)
        yo = R.scatter(mo, ei[1], 0, n, "mean" if mean else "sum")
        yo.sum().backward()
        mg = msg.to(DEV).requires_grad_(True)
        yg = p.ops.segment_reduce(mg, g, mean)
        yg.sum().backward()
        assert relerr(yg.detach(), yo.detach()) <= TOL
        assert relerr(mg.grad, mo.grad) <= TOL


def test_linearity_and_adjointness_at_arxiv_scale():
    """Size-independent properties at a BASELINE shape: A(ax+by) = aAx + bAy and <Ax,y> = <x,A^T y>."""
    p = P()
    import rgb_experiment_b200.synth as S
    sg = S.make_named("arxiv", features=False, device=DEV)
    n, F = sg.num_nodes, 128
    g = p.Graph(sg.edge_index, n, p.LOOP_ADD_REMAINING)
    gen = torch.Generator(device=DEV).manual_seed(5)
    x = torch.randn(n, F, device=DEV, generator=gen)
    y = torch.randn(n, F, device=DEV, generator=gen)
    A = lambda t: p.ops.spmm_raw(g.fwd, t, g.gcn_val(False))
    At = lambda t: p.ops.spmm_raw(g.bwd, t, g.gcn_val(True))
    lhs = A(2.0 * x - 3.0 * y)
    rhs = 2.0 * A(x) - 3.0 * A(y)
    assert relerr(lhs, rhs) <= TOL
    d1 = (A(x).double() * y.double()).sum()
    d2 = (x.double() * At(y).double()).sum()
    assert abs(d1 - d2) <= 1e-6 * max(abs(d1), 1.0)
    # constant vector: mean aggregation of a constant is the constant (rows with >= 1 edge)
    gm = p.Graph(sg.edge_index, n, p.LOOP_REMOVE_THEN_ADD)
    ones = torch.full((n, 8), 2.5, device=DEV)
    assert torch.equal(p.ops.propagate(ones, gm, "mean"), ones)


@pytest.mark.parametrize("case", ["loops_dups", "hub", "medium"])
def test_col_freq_and_hot_tag_bit_exact(case):
    """rgbmp_col_freq = bincount(col); rgbmp_col_tag sets bit 31 exactly on ids whose frequency
    reaches the threshold and leaves the low 31 bits untouched."""
    p = P()
    ei, n = CASES[case]()
    g = p.Graph(ei.to(DEV), n, p.LOOP_ADD_REMAINING)
    csr = g.fwd
    freq = csr.col_freq()[:n].cpu().long()
    assert torch.equal(freq, torch.bincount(csr.col.cpu().long(), minlength=n))
    import rgb_experiment_b200.graph as G_
    old = G_.HOT_L2_BYTES
    try:
        G_.HOT_L2_BYTES = 40 * 64                     # room for 40 rows of 64 bytes
        ref = csr.hot_ref(64)
        tagged, st, _ = csr._tagged[40]
        assert st.col_tagged == 1 and ref is not csr.ref
        t = tagged.cpu().long() & 0xFFFFFFFF
        assert torch.equal(t & 0x7FFFFFFF, csr.col.cpu().long())
        hot = (t >> 31).bool()
        k = 40
        thresh = int(torch.sort(freq, descending=True).values[k - 1]) + 1
        assert torch.equal(hot, freq[csr.col.cpu().long()] >= thresh)
        assert int((freq >= thresh).sum()) <= k
    finally:
        G_.HOT_L2_BYTES = old


@pytest.mark.parametrize("case", ["hub", "medium"])
@pytest.mark.parametrize("F", [7, 47, 100])
def test_hot_tagged_gathers_do_not_change_results(case, F):
    """The hot tag only selects an L2 eviction priority: results are bit-identical with and without it."""
    p = P()
    import rgb_experiment_b200.graph as G_
    ei, n = CASES[case]()
    x = torch.randn(n, F, generator=torch.Generator().manual_seed(F)).to(DEV)
    g = p.Graph(ei.to(DEV), n, p.LOOP_ADD_REMAINING)
    val = g.gcn_val(False)
    plain = p.ops.spmm_raw(g.fwd, x, val, hot=False)
    old = G_.HOT_L2_BYTES
    # a tagged graph never takes the one-cluster K-hop path (which sums long rows in another order): compare like with like
    cta = p._lib.lib().rgbmp_set_khop_cta(0)
    try:
        G_.HOT_L2_BYTES = 4096
        tagged = p.ops.spmm_raw(g.fwd, x, val, hot=True)
        assert len(g.fwd._tagged) == 1
        k10 = p.ops.appnp(x, g, 3, 0.1)
        G_.HOT_L2_BYTES = old
        assert torch.equal(plain, tagged)
        assert torch.equal(k10, p.ops.appnp(x, g, 3, 0.1))
    finally:
        G_.HOT_L2_BYTES = old
        p._lib.lib().rgbmp_set_khop_cta(cta)


@pytest.mark.parametrize("F,dtype", [(7, torch.float32), (24, torch.float32), (47, torch.float32), (64, torch.float32),
                                     (100, torch.float32), (300, torch.float32), (128, torch.bfloat16)])
@pytest.mark.parametrize("case", ["loops_dups", "medium"])
def test_peer_push_epilogue_fills_every_copy_in_both_forms(case, F, dtype):
    """Fused all-gather epilogue on one GPU: every `peer_out` buffer (here three local ones) receives rows
    [peer_row0, peer_row0 + n) = out2_scale * epilogue(row sum), nothing else is touched, and the bulk (TMA) form of the
    push writes the same bytes as the 16-byte stores.  (`medium` has rows for the long-row kernels, which always store.)"""
    p = P()
    L = p._lib.lib()
    ei, n = CASES[case]()
    g = p.Graph(ei.to(DEV), n, p.LOOP_ADD_REMAINING)
    d = g.dinv()
    x = torch.randn(n, F, generator=torch.Generator().manual_seed(F)).to(DEV).to(dtype)
    xb, ldx = p.ops.as_rows(x)
    ld = p.ops.padded_width(F, dtype)
    row0, tail = 3, 2
    tele = dict(a=0.9, b=0.1, T=xb, ldt=ldx)
    y = p.ops.spmm_raw(g.fwd, x, None, ep=p.ops.make_epilogue(row_scale=d, **tele), keep=(xb, d))
    want = (y.float() * d.view(-1, 1)).to(dtype)          # what the next hop gathers: one fp32 multiply, then the store's rounding
    got = {}
    old = L.rgbmp_set_push_bulk(1)
    try:
        for bulk in (1, 0):
            L.rgbmp_set_push_bulk(bulk)
            bufs = [torch.full((row0 + n + tail, ld), float("nan"), dtype=dtype, device=DEV) for _ in range(3)]
            ep = p.ops.make_epilogue(row_scale=d, out2_scale=d, peers=[b.data_ptr() for b in bufs], peer_row0=row0, ld_peer=ld, **tele)
            assert p.ops.spmm_raw(g.fwd, x, None, ep=ep, keep=(xb, d, bufs), store_local=False) is None
            torch.cuda.synchronize()
            for b in bufs:
                if dtype == torch.float32:
                    assert torch.equal(b[row0:row0 + n, :F], want)
                else:                                     # `want` was rounded to bf16 before the scale, the kernel rounds after
                    assert relerr(b[row0:row0 + n, :F].float(), want.float()) <= 2 ** -7
                assert torch.isnan(b[:row0].float()).all() and torch.isnan(b[row0 + n:].float()).all()
            got[bulk] = bufs
    finally:
        L.rgbmp_set_push_bulk(old)
    for a, b in zip(got[0], got[1]):
        # pad columns included (the long-row kernels store element by element and leave a row's pad untouched: NaN in both)
        assert torch.equal(torch.nan_to_num(a[row0:row0 + n].float(), nan=7.0), torch.nan_to_num(b[row0:row0 + n].float(), nan=7.0))
