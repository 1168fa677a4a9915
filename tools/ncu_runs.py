"""Split an ncu `--metrics gpu__time_duration.sum --csv` launch list into runs (a run ends with `end_kernel`) and print the
per-kernel totals of the last run of each distinct shape.  Usage: python tools/ncu_runs.py file.csv [end_kernel_substring]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
end = sys.argv[2] if len(sys.argv) > 2 else "horner"
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i
        break
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
gi = h.index("Grid Size") if "Grid Size" in h else None
seq = [(r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", ""), float(r[vi].replace(",", "")), r[gi] if gi is not None else "") for r in rows[start + 2:] if len(r) > vi]
runs, run = [], []
last_flag = int(sys.argv[3]) if len(sys.argv) > 3 else 0
for k, v, g in seq:
    run.append((k, v, g))
    if end in k and (not last_flag or "1, 1, 1" in g or True):
        runs.append(run); run = []
seen = {}
for r in runs:
    sig = tuple(k for k, _, _ in r)
    seen[sig] = r
for sig, r in seen.items():
    agg = collections.OrderedDict()
    for k, v, g in r:
        a = agg.setdefault(k, [0.0, 0]); a[0] += v; a[1] += 1
    tot = sum(a[0] for a in agg.values())
    print(f"---- run of {len(r)} launches, total {tot / 1000:.1f} us")
    for k, (v, c) in agg.items():
        print(f"  {k[:44]:44s} {v / 1000:9.1f} us  x{c}")
