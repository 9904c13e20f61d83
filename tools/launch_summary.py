#!/usr/bin/env python
"""Aggregate an ncu --csv launch list (gpu__time_duration.sum [+ dram__bytes_read/write.sum]) per kernel name.
    python tools/launch_summary.py gpurun_out/x.csv [...]"""
import collections
import csv
import sys

SCALE = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3,
         "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def summarise(path):
    hdr, agg = None, collections.OrderedDict()
    for r in csv.reader(open(path)):
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        if len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        try:
            val = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        name = d["Kernel Name"].split("(")[0]
        a = agg.setdefault(name, {"n": 0, "ms": 0.0, "bytes": 0.0})
        m = d["Metric Name"]
        if m == "gpu__time_duration.sum":
            a["n"] += 1
            a["ms"] += val * SCALE.get(d["Metric Unit"], 1.0)
        elif m.startswith("dram__bytes"):
            a["bytes"] += val * SCALE.get(d["Metric Unit"], 1.0)
    tot = sum(a["ms"] for a in agg.values())
    print(f"# {path}: {sum(a['n'] for a in agg.values())} launches, {tot:.3f} ms in kernels")
    print(f"# {'n':>4} {'ms':>9} {'share':>6} {'DRAM GB':>8} {'GB/s':>7}  kernel")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        gbs = a["bytes"] / 1e6 / a["ms"] if a["ms"] > 0 else 0.0
        print(f"  {a['n']:4d} {a['ms']:9.3f} {100 * a['ms'] / tot:5.1f}% {a['bytes'] / 1e9:8.2f} {gbs:7.0f}  {name}")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        summarise(p)
