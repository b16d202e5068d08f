#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of
bench.py into a per-kernel table of ONE steady step (the launches between two L2-flush memsets).
    python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_step.md"""
import collections
import csv
import re
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    byid = collections.OrderedDict()
    for r in rows:
        name = re.sub(r'\(.*', '', re.sub(r'nsp::<unnamed>::', '', r[4]))
        d = byid.setdefault(r[0], {"name": name, "grid": r[8]})
        d[r[12]] = float(r[14])
    L = list(byid.values())
    cuts = [i for i, d in enumerate(L) if 'FillFunctor<unsigned cha' in d["name"]]
    step = L[cuts[0] + 1:cuts[1]] if len(cuts) >= 2 else L      # `bench.py --ncu-step` captures exactly one step: take it all
    agg = collections.OrderedDict()
    for d in step:
        a = agg.setdefault(d["name"][:90], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0) / 1e3
        a[2] += (d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)) / 1e6
    tot = sum(v[1] for v in agg.values())
    print("One steady step: %d launches, %.1f us summed kernel time under ncu (cold caches, serialised)\n" % (len(step), tot))
    print("| kernel | launches | total us | share | avg us | DRAM MB (read+write) |")
    print("|---|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f%% | %.1f | %.1f |" % (k, v[0], v[1], 100 * v[1] / tot, v[1] / v[0], v[2]))
    ours = [d for d in step if "at::" not in d["name"]]
    gem = [d for d in ours if d["name"].startswith("void gemm_kernel") or "wgrad_kernel" in d["name"]]
    if gem:
        print("\nGEMM kernels: %d launches, %.1f us (%.1f%% of the step), DRAM %.1f MB per launch on average" % (
            len(gem), sum(d["gpu__time_duration.sum"] for d in gem) / 1e3,
            100 * sum(d["gpu__time_duration.sum"] for d in gem) / 1e3 / tot,
            sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in gem) / 1e6 / len(gem)))


if __name__ == "__main__":
    main(sys.argv[1])
