"""End-to-end accuracy parity (north_star: final test accuracy within 0.5 pt at a fixed seed).

/root/reference does not exist on the GPU box, so the reference driver cannot run there; this test
restates the reference's training unit (itexperiments.py:417-473: Adam, NLLLoss, 1 train
forward/backward + eval forwards per epoch, full batch) over thin stacks shaped like
models/gcn.py / graphsage.py / gat.py / appnp_stack.py, once on the CPU oracle layers and once on
the CUDA layers, from identical initial weights."""
import copy

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def build(L, kind, fin, hid, out):
    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            if kind == "gcn":
                self.c1, self.c2 = L.GCNConv(fin, hid), L.GCNConv(hid, out)
            elif kind == "sage":
                self.c1, self.c2 = L.SAGEConv(fin, hid), L.SAGEConv(hid, out)
            elif kind == "gat":
                self.c1, self.c2 = L.GATConv(fin, hid // 8, 8), L.GATConv(hid, out, 1, concat=False)
            elif kind == "appnp":
                self.c1, self.c2 = nn.Linear(fin, hid), nn.Linear(hid, out)
                self.prop = L.APPNP(10, 0.1)
            self.bn = nn.BatchNorm1d(hid)

        def forward(self, x, ei):
            if kind == "appnp":
                return F.log_softmax(self.prop(self.c2(self.bn(self.c1(x))), ei), dim=1)
            return F.log_softmax(self.c2(self.bn(self.c1(x, ei)), ei), dim=1)

    return Net()


def train(model, x, y, ei, tr, te, epochs=30, lr=0.01):
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    for _ in range(epochs):
        model.train()
        opt.zero_grad()
        loss = F.nll_loss(model(x, ei)[tr], y[tr])
        loss.backward()
        opt.step()
    model.eval()
    with torch.no_grad():
        pred = model(x, ei).argmax(1)
    return (pred[te] == y[te]).float().mean().item(), loss.item()


@pytest.mark.parametrize("kind", ["gcn", "sage", "gat", "appnp"])
def test_accuracy_parity_oracle_vs_cuda(kind):
    import importlib
    import rgb_experiment_b200.synth as S
    from oracle import layers as OL
    PL = importlib.import_module("rgb_experiment_b200.shim.nn")
    sg = S.make_graph(3000, 24000, 32, 6, seed=7)
    n = sg.num_nodes
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(0))
    tr, te = perm[: n * 6 // 10], perm[n * 8 // 10:]
    torch.manual_seed(14530529)
    mo = build(OL, kind, 32, 64, 6)
    torch.manual_seed(14530529)
    mg = build(PL, kind, 32, 64, 6)
    mg.load_state_dict(copy.deepcopy(mo.state_dict()))
    acc_o, loss_o = train(mo, sg.x, sg.y, sg.edge_index, tr, te)
    mg = mg.to(DEV)
    acc_g, loss_g = train(mg, sg.x.to(DEV), sg.y.to(DEV), sg.edge_index.to(DEV), tr.to(DEV), te.to(DEV))
    assert acc_o > 0.5                                   # the planted structure is learnable
    assert abs(acc_o - acc_g) <= 0.005, (acc_o, acc_g)   # 0.5 pt
    assert abs(loss_o - loss_g) <= 1e-2 * max(1.0, abs(loss_o))
