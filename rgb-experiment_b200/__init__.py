"""rgb-experiment_b200: B200-native (sm_100a) message passing behind the PyG names that
PolarisRisingWar/rgb-experiment imports.

Layers (host side, Python like the reference):
    shim/            torch_geometric / torch_scatter / torch_sparse name-compatible subset
    shim/pta.py      opt-in rebinding of the reference's PTA prelude onto the K-hop kernels
    ops.py           torch.autograd.Function drop-ins
    memo.py          exact memo of repeated no-grad forwards
    graph.py         cached CSR / transpose-CSR graph objects built by the integer kernels
    _lib.py          ctypes binding of the C ABI declared in include/rgbmp.h
    csrc/            hand-written CUDA kernels + the extern "C" entry points (librgbmp.so)

There is NO CPU fallback: every op raises if librgbmp.so is missing or a tensor is not on a
CUDA device.  The directory name carries a hyphen, so it is imported as ``rgb_experiment_b200``
through the loader module of that name at the repository root.
"""
from . import _lib  # noqa: F401
from .graph import Graph, get_graph, LOOP_NONE, LOOP_ADD, LOOP_ADD_REMAINING, LOOP_REMOVE_THEN_ADD  # noqa: F401
from . import ops  # noqa: F401
from . import memo  # noqa: F401
from .shim import install as install_shim, uninstall as uninstall_shim  # noqa: F401
from .shim.pta import patch_reference as patch_pta  # noqa: F401

__version__ = "0.1.0"
