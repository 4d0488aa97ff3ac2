"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys

for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    start = rows.index(hdr) + 1
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start:]:
        name = r[ki].split("(")[0][:70]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"== {path}: {sum(v[0] for v in agg.values())} launches, {tot/1e3:.2f} ms of kernel time")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"  {k:72s} n={v[0]:4d} total={v[1]:10.1f} us avg={v[1]/v[0]:9.1f} share={v[1]/tot:.3f}")
