"""CPU oracle for the message-passing hot path -- TEST INFRASTRUCTURE ONLY.

This package is a pure-torch restatement (pyg_restated.py, layers.py) plus a plain-C one (csrc/rgb_oracle.c,
scalar loops in edge order, bound by c_oracle.py; bit-identical to the torch form) of the arithmetic that the
reference (PolarisRisingWar/rgb-experiment) delegates to torch_geometric /
torch_scatter / torch_sparse (none of which exist in this image or on the GPU
box).  It is the checker the CUDA path
is compared against.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product package (``rgb-experiment_b200/``) never does.

Parity pinning: the reference has NO tests, golden vectors or fixtures for this
path, and its arithmetic lives in un-vendored, un-pinned
third-party packages (PyG 1.7-2.0.x era).  The oracle is therefore pinned
against outputs of the reference's own in-tree, PyG-free functions executed in
the build container (``rgb_experiment/itexperiments.py:671-719`` normalize_adj /
label_propagation, ``rgb_experiment/models/pta.py:79-84`` PTA.inference,
``rgb_experiment/models/dagnn.py:12-31`` gcn_norm and ``:34-65`` Prop), committed
as ``tests/golden/*.npz`` by ``tests/golden/make_golden.py``.  Operators that
exist ONLY inside PyG (GATConv, SuperGATConv, FAConv, CorrectAndSmooth, softmax)
have no runnable reference here: for those the status is "parity unpinned" --
they follow the published PyG algorithm and are
checked by closed forms and fp64 autograd identities instead.
"""
from . import pyg_restated  # noqa: F401
