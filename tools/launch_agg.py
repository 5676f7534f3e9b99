"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name (mean / max / count)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
i0 = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[i0]
k, v = h.index("Kernel Name"), h.index("Metric Value")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
data = [(r[k].split("(")[0][:60], float(r[v].replace(",", ""))) for r in rows[i0 + 1:] if len(r) > v]
data = [d for d in data if not d[0].startswith(("void at::", "at::", "void thrust", "void cub"))][skip:]
if len(sys.argv) > 3:
    for n, t in data[-int(sys.argv[3]):]:
        print(f"{t:10.1f} {n}")
agg = collections.defaultdict(list)
for n, t in data:
    agg[n].append(t)
tot = sum(t for _, t in data)
print("launches", len(data), "total", tot)
for n, l in sorted(agg.items(), key=lambda a: -sum(a[1])):
    print(f"{n:60s} n={len(l):5d} mean={sum(l) / len(l):10.1f} max={max(l):10.1f} share={100 * sum(l) / tot:5.1f}%")
