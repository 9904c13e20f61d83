#!/usr/bin/env python
"""Build an A/B copy of librgbmp.so with extra -D flags on some translation units:
    python tools/build_variant.py TAG -DRGBMP_ATT_OCC=2 [--only att.cu,spmm_inst_f32v.cu]
writes rgb-experiment_b200/build/variants/librgbmp_TAG.so (git-ignored, travels to the GPU box); run a tool against
it with RGBMP_LIB=<that path>.  Units not named by --only are taken from the regular build's object cache."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rgb-experiment_b200"))
import build as B  # noqa: E402


def main():
    tag = sys.argv[1]
    defs = [a for a in sys.argv[2:] if a.startswith("-D")]
    only = None
    for i, a in enumerate(sys.argv):
        if a == "--only":
            only = set(sys.argv[i + 1].split(","))
    B.build()
    out_dir = os.path.join(B.BUILD, "variants")
    os.makedirs(out_dir, exist_ok=True)
    srcs = sorted(os.path.join(B.CSRC, f) for f in os.listdir(B.CSRC) if f.endswith(".cu"))

    def one(src):
        name = os.path.basename(src)
        if only is not None and name not in only:
            return B._compile(src, False)
        obj = os.path.join(out_dir, f"{name[:-3]}.{tag}.o")
        r = subprocess.run([B.NVCC, *B.FLAGS, *defs, "-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(one, srcs))
    lib = os.path.join(out_dir, f"librgbmp_{tag}.so")
    r = subprocess.run([B.NVCC, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                        "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr)
    print("built", lib)


if __name__ == "__main__":
    main()
