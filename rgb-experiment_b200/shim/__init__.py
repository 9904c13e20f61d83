"""Install the CUDA-backed stand-ins for torch_geometric / torch_scatter / torch_sparse into
``sys.modules`` (exact symbol list: SURVEY.md 8b), so that

    from rgb_experiment import experiment
    experiment(model_init_param, model_name='gcn', specify_data=True, data=Data(...))

runs UNCHANGED on the B200 kernels.  Call ``install()`` before importing the reference package.
"""
from __future__ import annotations

import sys
import types

_NAMES = ("torch_geometric", "torch_geometric.nn", "torch_geometric.nn.conv", "torch_geometric.utils",
          "torch_geometric.data", "torch_scatter", "torch_sparse")


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__rgbmp_shim__ = True
    return m


def install_matplotlib_stub() -> bool:
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

    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    mpl.rcParams = {}
    fm = types.ModuleType("matplotlib.font_manager")
    fm.FontProperties = _Anything
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
    from .. import _lib
    _lib.lib()                                         # fail loudly if librgbmp.so is missing
    from . import nn as L
    from . import utils as U
    from .data import Data
    install_matplotlib_stub()
    conv_names = ("MessagePassing", "GCNConv", "SAGEConv", "GATConv", "SuperGATConv", "APPNP", "SGConv",
                  "FAConv", "GINConv", "GatedGraphConv")
    conv = _mod("torch_geometric.nn.conv", **{n: getattr(L, n) for n in conv_names})
    nn_ = _mod("torch_geometric.nn", conv=conv, CorrectAndSmooth=L.CorrectAndSmooth,
               LabelPropagation=L.LabelPropagation, **{n: getattr(L, n) for n in conv_names})
    utils = _mod("torch_geometric.utils", remove_self_loops=U.remove_self_loops, add_self_loops=U.add_self_loops,
                 add_remaining_self_loops=U.add_remaining_self_loops, to_undirected=U.to_undirected,
                 to_networkx=U.to_networkx, dropout_adj=U.dropout_adj, negative_sampling=U.negative_sampling)
    data = _mod("torch_geometric.data", Data=Data)
    tg = _mod("torch_geometric", nn=nn_, utils=utils, data=data, __version__="rgbmp-b200")
    ts = _mod("torch_scatter", scatter_add=U.scatter_add, scatter=U.scatter)
    tsp = _mod("torch_sparse", coalesce=U.coalesce)
    mods = {"torch_geometric": tg, "torch_geometric.nn": nn_, "torch_geometric.nn.conv": conv,
            "torch_geometric.utils": utils, "torch_geometric.data": data, "torch_scatter": ts, "torch_sparse": tsp}
    sys.modules.update(mods)
    return mods


def uninstall():
    for n in _NAMES:
        m = sys.modules.get(n)
        if m is not None and getattr(m, "__rgbmp_shim__", False):
            del sys.modules[n]
