"""Oracle side of the end-to-end accuracy parity: run the UNMODIFIED reference driver
``rgb_experiment.experiment()`` (imported from /root/reference, never copied) on the CPU ORACLE shim for
every case of ``tests/experiment_cases.py`` and commit the accuracies.

    python tests/golden/make_experiment_golden.py [--threads T] [case ...]      # build container only

Output: tests/golden/experiment_acc.json  {case: {"ACC": ..., "f1_macro": ..., "ACC_by_threads": {T: acc}}}.
``tests/test_z_gpu_experiment.py`` makes the same calls on a B200 through the product shim and requires the
CUDA accuracy within 0.5 pt of the oracle's.  Training is a chaotic map of its rounding errors for some of the
reference's models (GIN: sums over hub neighbourhoods into ReLU stacks, early stopping on the validation
accuracy): the SAME oracle code run with 1, 3 or 8 host threads (a different fp32 summation order in the
GEMMs, nothing else) moves GIN's accuracy by a full point.  ``--threads T`` re-runs the cases with T threads and
records the accuracy under ACC_by_threads; the test takes the interval those runs span as "the oracle's accuracy".  The PTA case runs the reference's own scipy prelude and Python label
propagation (itexperiments.py:351-372, 671-719) -- no shim code at all on the oracle side.
"""
import json
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import experiment_cases as EC  # noqa: E402
from oracle import shim  # noqa: E402


def main(names, threads=None):
    if threads:
        torch.set_num_threads(threads)
    shim.install()
    sys.path.insert(0, "/root/reference")
    import rgb_experiment as rgb
    from torch_geometric.data import Data
    out = {}
    if os.path.exists(EC.GOLDEN_JSON):
        out = json.load(open(EC.GOLDEN_JSON))
    cache = {}
    for name in names:
        t0 = time.time()
        r = EC.run_case(rgb, Data, name, {"use_cpu": True}, cache)
        ent = out.get(name, {})
        if not threads or "ACC" not in ent:
            ent.update({"ACC": r["ACC"], "f1_macro": float(r["f1_macro"]), "seconds": round(time.time() - t0, 1),
                        "torch": torch.__version__, "threads": torch.get_num_threads()})
        ent.setdefault("ACC_by_threads", {})[str(torch.get_num_threads())] = r["ACC"]
        ent.setdefault("f1_by_threads", {})[str(torch.get_num_threads())] = float(r["f1_macro"])
        out[name] = ent
        print(name, out[name], flush=True)
        with open(EC.GOLDEN_JSON, "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    argv = sys.argv[1:]
    T = None
    if argv[:1] == ["--threads"]:
        T, argv = int(argv[1]), argv[2:]
    main(argv or list(EC.CASES), T)
