"""Build librgbmp.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python rgb-experiment_b200/build.py [--force]

Objects are cached under rgb-experiment_b200/build/ keyed on a hash of the source, the headers
and the flags; translation units are compiled in parallel.  The resulting
rgb-experiment_b200/librgbmp.so is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "librgbmp.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-I", INCLUDE]


def _hash_inputs(src: str) -> str:
    h = hashlib.sha256()
    h.update(" ".join(FLAGS).encode())
    deps = [src] + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    deps.append(os.path.join(INCLUDE, "rgbmp.h"))
    for d in deps:
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def _compile(src: str, force: bool) -> str:
    name = os.path.splitext(os.path.basename(src))[0]
    tag = _hash_inputs(src)
    obj = os.path.join(BUILD, f"{name}.{tag}.o")
    if os.path.exists(obj) and not force:
        return obj
    for old in os.listdir(BUILD):
        if old.startswith(name + ".") and old.endswith(".o"):
            os.remove(os.path.join(BUILD, old))
    cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force), srcs))
    stamp = os.path.join(BUILD, "link.stamp")
    want = " ".join(sorted(os.path.basename(o) for o in objs))
    have = open(stamp).read() if os.path.exists(stamp) else ""
    if force or want != have or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        with open(stamp, "w") as f:
            f.write(want)
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
