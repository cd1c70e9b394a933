"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: usage launch_shares.py file.csv [first_kernel_substring]
(with a substring: only the launches from its LAST occurrence on, i.e. the last timed call)."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
start = 0
if len(sys.argv) > 2:
    idx = [i for i, x in enumerate(rows) if sys.argv[2] in x["Kernel Name"]]
    start = idx[-1] if idx else 0
agg, tot = collections.OrderedDict(), 0.0
for x in rows[start:]:
    n = x["Kernel Name"].split("(")[0].replace("void ", "").replace("slideo::<unnamed>::", "")[:70]
    v = float(x["Metric Value"].replace(",", ""))
    v = v / 1e3 if x["Metric Unit"] == "ns" else v * 1e3 if x["Metric Unit"] == "ms" else v
    a = agg.setdefault(n, [0.0, 0])
    a[0] += v
    a[1] += 1
    tot += v
print(f"{'us':>12s} {'n':>5s} {'share':>6s}  kernel")
for n, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{v:12.1f} {c:5d} {100 * v / tot:5.1f}%  {n}")
print(f"{tot:12.1f} total")
