"""Install the ORACLE (pure CPU torch) stand-ins for torch_geometric / torch_scatter /
torch_sparse into ``sys.modules`` so that the reference package imports and runs
(TEST INFRASTRUCTURE; SURVEY.md section 8b lists the exact symbols).

Used by tests to (a) execute the reference's in-tree PyG-free functions for golden
vectors and (b) produce the oracle side of end-to-end ``experiment()`` accuracy parity.
"""
from __future__ import annotations

import sys
import types

from . import layers as L
from . import pyg_restated as R

_NAMES = ("torch_geometric", "torch_geometric.nn", "torch_geometric.nn.conv",
          "torch_geometric.utils", "torch_geometric.data", "torch_scatter", "torch_sparse")


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__oracle_shim__ = True
    return m


def install_matplotlib_stub():
    """itexperiments.py:1,27 and visualize_feature.py:6-8 import matplotlib at import time only."""
    try:
        import matplotlib  # noqa: F401
        return False
    except Exception:
        pass

    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Anything()

        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return _Anything()

    mpl = _mod("matplotlib", use=lambda *a, **k: None, rcParams={})
    fm = _mod("matplotlib.font_manager", FontProperties=_Anything)
    plt = types.ModuleType("matplotlib.pyplot")

    def _plt_getattr(name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    plt.__getattr__ = _plt_getattr                   # type: ignore[attr-defined]
    mpl.font_manager, mpl.pyplot = fm, plt
    sys.modules.update({"matplotlib": mpl, "matplotlib.font_manager": fm, "matplotlib.pyplot": plt})
    return True


def install():
    """(Re)install the oracle shim; returns the dict of modules placed in sys.modules."""
    install_matplotlib_stub()
    conv_names = ("MessagePassing", "GCNConv", "SAGEConv", "GATConv", "SuperGATConv", "APPNP",
                  "SGConv", "FAConv", "GINConv", "GatedGraphConv")
    conv = _mod("torch_geometric.nn.conv", **{n: getattr(L, n) for n in conv_names})
    nn_ = _mod("torch_geometric.nn", conv=conv, CorrectAndSmooth=L.CorrectAndSmooth,
               LabelPropagation=L.LabelPropagation, **{n: getattr(L, n) for n in conv_names})
    utils = _mod("torch_geometric.utils", remove_self_loops=R.remove_self_loops,
                 add_self_loops=R.add_self_loops, add_remaining_self_loops=R.add_remaining_self_loops,
                 to_undirected=R.to_undirected, to_networkx=L.to_networkx, softmax=R.softmax,
                 dropout_adj=L.dropout_adj, negative_sampling=L.negative_sampling, degree=None)
    data = _mod("torch_geometric.data", Data=L.Data)
    tg = _mod("torch_geometric", nn=nn_, utils=utils, data=data, __version__="oracle")
    ts = _mod("torch_scatter", scatter_add=R.scatter_add, scatter=R.scatter)
    tsp = _mod("torch_sparse", coalesce=R.coalesce)
    mods = {"torch_geometric": tg, "torch_geometric.nn": nn_, "torch_geometric.nn.conv": conv,
            "torch_geometric.utils": utils, "torch_geometric.data": data,
            "torch_scatter": ts, "torch_sparse": tsp}
    sys.modules.update(mods)
    return mods


def uninstall():
    for n in _NAMES:
        m = sys.modules.get(n)
        if m is not None and getattr(m, "__oracle_shim__", False):
            del sys.modules[n]


def purge_reference():
    """Forget an imported ``rgb_experiment`` so it can be re-imported against another shim."""
    for n in [k for k in sys.modules if k == "rgb_experiment" or k.startswith("rgb_experiment.")]:
        del sys.modules[n]
