"""Oracle side of the end-to-end accuracy parity: run the UNMODIFIED reference driver
``rgb_experiment.experiment()`` (imported from /root/reference, never copied) on the CPU ORACLE shim for
every case of ``tests/experiment_cases.py`` and commit the accuracies.

    python tests/golden/make_experiment_golden.py [case ...]      # build container only

Output: tests/golden/experiment_acc.json  {case: {"ACC": ..., "f1_macro": ..., "seconds": ...}}.
``tests/test_z_gpu_experiment.py`` makes the same calls on a B200 through the product shim and requires
|ACC_cuda - ACC_oracle| <= 0.5 pt.  The PTA case runs the reference's own scipy prelude and Python label
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


def main(names):
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
        out[name] = {"ACC": r["ACC"], "f1_macro": float(r["f1_macro"]), "seconds": round(time.time() - t0, 1),
                     "torch": torch.__version__, "threads": torch.get_num_threads()}
        print(name, out[name], flush=True)
        with open(EC.GOLDEN_JSON, "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1:] or list(EC.CASES))
