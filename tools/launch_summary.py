"""Per-kernel launch count and median duration from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i
        break
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
d = defaultdict(list)
order = []
for r in rows[start + 1:]:
    if len(r) > vi:
        k = r[ki][:70]
        if k not in d:
            order.append(k)
        d[k].append(float(r[vi].replace(",", "")))
for k in order:
    v = sorted(d[k])
    print("%-72s n=%-4d median %8.1f us   total %9.1f us" % (k, len(v), v[len(v) // 2] / 1000, sum(v) / 1000))
