"""Locality groups for the row schedule (csrc/cluster.cu, graph.locality_groups): integer work, so everything is
checked exactly -- the device label propagation against a plain torch restatement of the same rule, the schedule as a
permutation sorted by (group, -degree), the long-row lists in schedule order, and, above all, that a grouped schedule
changes NO result bit of the aggregation (only the order in which rows are processed)."""
import ctypes as C

import pytest
import torch

from helpers import CASES

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def P():
    import rgb_experiment_b200 as P_
    return P_


def lpa_restated(rowptr, col, S, taus):
    """The rule of rgbmp_cluster_lpa in torch ops on the CPU (int64): seeds = S highest-degree rows (stable), then
    leaves-first plurality voting; unlabelled leftovers get id % S."""
    n = rowptr.numel() - 1
    deg = rowptr[1:] - rowptr[:-1]
    order = torch.argsort(deg, descending=True, stable=True)
    label = torch.full((n,), -1, dtype=torch.int64)
    label[order[:S]] = torch.arange(S)
    row = torch.repeat_interleave(torch.arange(n), deg)
    colL = col.long()
    for tau in taus:
        lj = label[colL]
        m = lj >= 0
        nlab = torch.zeros(n, dtype=torch.int64).index_add_(0, row[m], torch.ones_like(row[m]))
        uk, cnt = torch.unique(row[m] * S + lj[m], return_counts=True)
        best = torch.full((n,), -1, dtype=torch.int64)
        best.scatter_reduce_(0, uk // S, cnt * S + (S - 1 - uk % S), reduce="amax", include_self=True)
        ok = (label < 0) & (best >= 0) & (nlab.double() >= torch.tensor(tau, dtype=torch.float32).double() * deg.double())
        label = torch.where(ok, S - 1 - (best % S), label)
    return torch.where(label < 0, torch.arange(n) % S, label)


@pytest.mark.parametrize("case", ["loops_dups", "isolated", "hub", "medium"])
@pytest.mark.parametrize("S", [4, 37])
def test_label_propagation_matches_the_restated_rule_exactly(case, S):
    p = P()
    from rgb_experiment_b200 import graph as G
    from rgb_experiment_b200._lib import GraphStruct, check, lib, ptr, stream_of
    ei, n = CASES[case]()
    S = min(S, n)
    g = p.Graph(ei.to(DEV), n, p.LOOP_ADD_REMAINING)
    csr = g.fwd
    taus = G.CLUSTER_TAUS
    plain = GraphStruct(n, n, csr.nnz, ptr(csr.rowptr), ptr(csr.col), 0, 0, 0, 0, None, None, None, None, None, 0)
    label = torch.empty(n, dtype=torch.int32, device=DEV)
    ws = G._ws(lib().rgbmp_cluster_workspace_bytes(n), csr.device)
    check(lib().rgbmp_cluster_lpa(C.byref(plain), ptr(csr.degree_order()), S, len(taus), (C.c_float * len(taus))(*taus),
                                  ptr(label), ptr(ws), ws.numel(), 0, stream_of(csr.device)), "lpa")
    want = lpa_restated(csr.rowptr.cpu(), csr.col.cpu(), S, taus)
    assert torch.equal(label.cpu().long(), want)
    W = torch.empty((S, S), dtype=torch.int32, device=DEV)
    check(lib().rgbmp_cluster_connectivity(C.byref(plain), ptr(label), S, 1, ptr(W), 0, stream_of(csr.device)), "conn")
    deg = csr.rowptr[1:] - csr.rowptr[:-1]
    row = torch.repeat_interleave(torch.arange(n, device=DEV), deg)
    Wo = torch.zeros(S * S, dtype=torch.int64, device=DEV).index_add_(
        0, label.long()[row] * S + label.long()[csr.col.long()], torch.ones(csr.nnz, dtype=torch.int64, device=DEV))
    assert torch.equal(W.long().view(-1), Wo)
    # sampled form: every 3rd row only
    check(lib().rgbmp_cluster_connectivity(C.byref(plain), ptr(label), S, 3, ptr(W), 0, stream_of(csr.device)), "conn")
    keep = (row % 3) == 0
    Ws = torch.zeros(S * S, dtype=torch.int64, device=DEV).index_add_(
        0, (label.long()[row] * S + label.long()[csr.col.long()])[keep], torch.ones(int(keep.sum()), dtype=torch.int64, device=DEV))
    assert torch.equal(W.long().view(-1), Ws)


def test_grouped_schedule_is_sorted_by_group_then_degree_and_changes_no_result_bit(monkeypatch):
    p = P()
    from rgb_experiment_b200 import graph as G
    import rgb_experiment_b200.synth as S_
    sg = S_.make_graph(60_000, 3_000_000, 8, 12, seed=9, features=False, device=DEV)     # 12 planted classes, hubs > chunk
    n = sg.num_nodes
    monkeypatch.setattr(G, "CLUSTER", "0")
    g0 = p.Graph(sg.edge_index, n, p.LOOP_ADD_REMAINING)
    monkeypatch.setattr(G, "CLUSTER", "1")
    monkeypatch.setattr(G, "CLUSTER_SEEDS", 64)
    g1 = p.Graph(sg.edge_index, n, p.LOOP_ADD_REMAINING)
    assert not g0.fwd.clustered and g1.fwd.clustered and g1.groups is not None
    assert G.cluster_stats["last"]["intra_group_edge_share"] > 0.3       # the planted communities were found from the edges
    grp, ngrp = g1.groups
    assert int(grp.min()) >= 0 and int(grp.max()) < ngrp
    for csr in (g1.fwd, g1.bwd):
        order = csr.row_order.long()
        assert torch.equal(torch.sort(order).values, torch.arange(n, device=DEV))           # a permutation
        deg = (csr.rowptr[1:] - csr.rowptr[:-1])[order].clamp(max=65535)
        key = grp.long()[order] * 65536 + (65535 - deg)
        assert bool((key[1:] >= key[:-1]).all())                                            # (group, -degree) order
        assert csr.n_long > 0
        pos = torch.empty(n, dtype=torch.long, device=DEV)
        pos[order] = torch.arange(n, device=DEV)
        lp = pos[csr.long_rows.long()]
        assert bool((lp[1:] > lp[:-1]).all())                                               # long rows listed in schedule order
        items = ((csr.rowptr[1:] - csr.rowptr[:-1])[csr.long_rows.long()] + csr.long_chunk - 1) // csr.long_chunk
        assert torch.equal(csr.long_item_ptr.long()[1:] - csr.long_item_ptr.long()[:-1], items)
    x = torch.randn(n, 47, device=DEV)
    for fold in (False, True):
        assert torch.equal(p.ops.appnp(x, g0, 3, 0.1, fold), p.ops.appnp(x, g1, 3, 0.1, fold))
    xg0, xg1 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    p.ops.propagate(xg0, g0, "gcn").pow(2).sum().backward()
    p.ops.propagate(xg1, g1, "gcn").pow(2).sum().backward()
    assert torch.equal(xg0.grad, xg1.grad)
    a = torch.randn(n, 8, device=DEV)
    xp = torch.randn(n, 64, device=DEV)
    gg0 = p.ops.gat(xp, a, a.flip(1), g0, 8, 8, 0.2)
    gg1 = p.ops.gat(xp, a, a.flip(1), g1, 8, 8, 0.2)
    assert torch.equal(gg0, gg1)


def test_graph_without_communities_keeps_the_degree_schedule(monkeypatch):
    p = P()
    from rgb_experiment_b200 import graph as G
    import rgb_experiment_b200.synth as S_
    monkeypatch.setattr(G, "CLUSTER_MIN_NODES", 1000)
    sg = S_.make_graph(50_000, 1_000_000, 8, 4, seed=3, features=False, device=DEV, power_law=False, homophily=0.0)
    g = p.Graph(sg.edge_index, sg.num_nodes, p.LOOP_ADD_REMAINING)
    assert G.cluster_stats["last"]["intra_group_edge_share"] < G.CLUSTER_MIN_INTRA
    assert g.groups is None and not g.fwd.clustered
    deg = (g.fwd.rowptr[1:] - g.fwd.rowptr[:-1])[g.fwd.row_order.long()]
    assert bool((deg[1:] <= deg[:-1]).all())
