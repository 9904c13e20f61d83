"""Run the UNMODIFIED reference driver (``/root/reference``, build container only) on the CPU
oracle shim for every model name: proves the import surface / call signatures of SURVEY.md 8b
are complete, and is BASELINE.json configs[0] (GCN 2-layer hidden 64 on a Cora-shaped graph)."""
import sys

import pytest
import torch

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref():
    from oracle import shim
    shim.purge_reference()
    shim.install()
    sys.path.insert(0, "/root/reference")
    import rgb_experiment
    from torch_geometric.data import Data
    yield rgb_experiment, Data
    sys.path.remove("/root/reference")
    shim.purge_reference()
    shim.uninstall()


def small_data(Data, n=300, e=1800, f=24, c=4, seed=0):
    import rgb_experiment_b200.synth as S
    g = S.make_graph(n, e, f, c, seed=seed)
    return Data(x=g.x, y=g.y, edge_index=g.edge_index)


def test_gcn_cora_shaped_config0(ref):
    rgb, Data = ref
    import rgb_experiment_b200.synth as S
    g = S.make_named("cora")
    data = Data(x=g.x, y=g.y, edge_index=g.edge_index)
    r = rgb.experiment({"num_layers": 2, "hidden_unit": 64, "dropout_rate": 0.5}, model_name="gcn",
                       specify_data=True, data=data, use_cpu=True, need_to_reappear=True, epoch=15,
                       print_print=False)
    assert r["ACC"] > 0.5                       # planted classes are learnable; chance is 1/7


PARAMS = {
    "mlp": {"num_layers": 3, "hidden_unit": 16, "dropout_rate": 0.5},
    "gcn": {"num_layers": 2, "hidden_unit": 16, "dropout_rate": 0.5},
    "graphsage": {"num_layers": 2, "hidden_unit": 16, "dropout_rate": 0.5},
    "gat": {"num_layers": 2, "hidden_unit": 4, "dropout_rate": 0.5, "heads": 4},
    "ggnn": {"num_layers": 2, "hidden_unit": 32, "dropout_rate": 0.5},
    "appnpstack": {"hidden_unit": 16, "dropout_rate": 0.5, "alpha": 0.1, "K": 5},
    "graphsage2": {"num_layers": 2, "hidden_unit": 16, "dropout_rate": 0.5},
    "pta": {"nhid": 16, "dropout": 0, "epsilon": 100, "mode": 2, "K": 5, "alpha": 0.1},
    "dagnn": {"hidden_dim": 16, "K": 5, "dropout_rate": 0.5},
    "supergat": {"hidden_dim": 4, "heads": 4, "dropout_rate": 0.6, "edge_sample_ratio": 0.8, "neg_sample_ratio": 0.5},
    "sgc": {"K": 2},
    "gin": {"num_layers": 2, "hidden_unit": 16, "dropout_rate": 0.5},
    "fagcn": {"num_layers": 2, "hidden_unit": 16, "dropout_rate": 0.5, "epsilon": 0.3},
}


@pytest.mark.parametrize("name", sorted(PARAMS))
def test_every_model_name_runs_on_the_shim(ref, name):
    rgb, Data = ref
    data = small_data(Data)
    r = rgb.experiment(PARAMS[name], model_name=name, specify_data=True, data=data, use_cpu=True,
                       need_to_reappear=True, epoch=4, print_print=False)
    assert 0.0 <= r["ACC"] <= 1.0


def test_correct_and_smooth_post_process(ref):
    rgb, Data = ref
    data = small_data(Data)
    r = rgb.experiment(PARAMS["mlp"], model_name="mlp", specify_data=True, data=data, use_cpu=True,
                       need_to_reappear=True, epoch=4, print_print=False, post_cs=True,
                       cs_param=rgb.InitialParameters.default_cs_param)
    assert 0.0 <= r["ACC"] <= 1.0
