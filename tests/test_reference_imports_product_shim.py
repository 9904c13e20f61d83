"""The UNMODIFIED reference package against the PRODUCT shim, in the build container (no GPU):
`models/__init__.py:1-13` imports all 13 model files eagerly, so a successful import proves that
the CUDA-backed stand-ins cover the whole import surface of SURVEY.md 8b with the real code, every
model constructs with the reference's own constructor calls, and a CPU run fails LOUDLY (the
product has no CPU path) instead of silently falling back."""
import sys

import pytest
import torch

pytestmark = pytest.mark.reference


@pytest.fixture()
def ref():
    from oracle import shim as oshim
    import rgb_experiment_b200 as R
    oshim.purge_reference()
    oshim.uninstall()
    R.install_shim()
    sys.path.insert(0, "/root/reference")
    import rgb_experiment
    yield rgb_experiment, R
    sys.path.remove("/root/reference")
    oshim.purge_reference()
    R.uninstall_shim()


def test_reference_imports_and_builds_every_model_on_the_product_shim(ref):
    rgb, R = ref
    import rgb_experiment.models as M
    import torch_geometric
    assert getattr(torch_geometric, "__rgbmp_shim__", False)
    built = {
        "GCN": M.GCN(input_dim=12, output_dim=3, num_layers=2, hidden_unit=8, dropout_rate=0.5),
        "GraphSAGE": M.GraphSAGE(input_dim=12, output_dim=3, num_layers=2, hidden_unit=8, dropout_rate=0.5),
        "GAT": M.GAT(input_dim=12, output_dim=3, num_layers=2, hidden_unit=2, dropout_rate=0.5, heads=4),
        "APPNPStack": M.APPNPStack(input_dim=12, output_dim=3, hidden_unit=8, dropout_rate=0.5, alpha=0.1, K=5),
        "SGC": M.SGC(input_dim=12, output_dim=3, K=2),
        "DAGNN": M.DAGNN(input_dim=12, output_dim=3, hidden_dim=8, K=5, dropout_rate=0.5),
        "FAGCN": M.FAGCN(input_dim=12, output_dim=3, num_layers=2, hidden_unit=8, dropout_rate=0.5, epsilon=0.3),
        "SuperGAT": M.SuperGAT(input_dim=12, output_dim=3, hidden_dim=2, heads=4, dropout_rate=0.6,
                               edge_sample_ratio=0.8, neg_sample_ratio=0.5),
        "GIN": M.GIN(input_dim=12, output_dim=3, num_layers=2, hidden_unit=8, dropout_rate=0.5),
        "GGNN": M.GGNN(input_dim=12, output_dim=3, num_layers=2, hidden_unit=16, dropout_rate=0.5),
        "GraphSAGE2": M.GraphSAGE2(input_dim=12, output_dim=3, num_layers=2, hidden_unit=8, dropout_rate=0.5),
    }
    import rgb_experiment_b200.shim.nn as PL
    for name, model in built.items():
        convs = [m for m in model.modules() if isinstance(m, (PL.MessagePassing, PL.LabelPropagation))]
        assert convs, name                                   # every GNN model contains at least one shim layer


def test_cpu_run_fails_loudly_not_silently(ref):
    rgb, R = ref
    from torch_geometric.data import Data
    import rgb_experiment_b200.synth as S
    g = S.make_graph(120, 600, 8, 3, seed=1)
    data = Data(x=g.x, y=g.y, edge_index=g.edge_index)
    with pytest.raises(RuntimeError, match="no CPU path"):
        rgb.experiment({"num_layers": 2, "hidden_unit": 8, "dropout_rate": 0.5}, model_name="gcn", specify_data=True,
                       data=data, use_cpu=True, need_to_reappear=True, epoch=2, print_print=False)


def test_patch_pta_binds_and_restores_the_reference_names(ref):
    rgb, R = ref
    import rgb_experiment.itexperiments as it
    import rgb_experiment.models.pta as pm
    from rgb_experiment_b200.shim.pta import PtaAdjacency
    orig = (it.edge_index2sparse_matrix, it.normalize_adj, it.label_propagation, pm.PTA.inference)
    restore = R.patch_pta()
    try:
        h = it.edge_index2sparse_matrix(torch.tensor([[0, 1], [1, 0]]), 2)
        assert isinstance(h, PtaAdjacency)
        import scipy.sparse as sp
        h = it.sparse_mx_to_torch_sparse_tensor(it.normalize_adj(h + sp.eye(2)))
        with pytest.raises(RuntimeError, match="CUDA"):       # product backend: no CPU path
            h.to("cpu")
    finally:
        restore()
    assert (it.edge_index2sparse_matrix, it.normalize_adj, it.label_propagation, pm.PTA.inference) == orig
