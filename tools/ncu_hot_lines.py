"""Top source lines by warp-stall samples, per kernel, from
`ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > X.csv`;  usage: ncu_hot_lines.py X.csv [N] [kernel-substring]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
only = sys.argv[3] if len(sys.argv) > 3 else ""
hdr, fname, func = None, "?", "?"
agg = defaultdict(lambda: defaultdict(float))
src = {}
line = None
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Function Name":
        func = r[1].split("(")[0].replace("void ", "").replace("b200r::", "")
        continue
    if only and only not in func:
        continue
    if len(r) > 10 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[0]:
        line = (func.split('<')[0] + ' ' + fname, int(r[0]))
        src[line] = r[1].strip()
    if not r[2]:
        continue
    for k, v in zip(hdr[4:], r[4:]):
        try:
            agg[line][k] += float(v)
        except ValueError:
            pass
tot = sum(a["# Samples"] for a in agg.values())
tot_i = sum(a["Instructions Executed"] for a in agg.values())
print("total samples %d, warp instructions %d" % (tot, tot_i))
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    s = sorted(((a[k], k[6:]) for k in stalls), reverse=True)[:3]
    print("%5.1f%% smp %5.1f%% ins  %s:%d  %-70s %s" % (
        100 * a["# Samples"] / tot, 100 * a["Instructions Executed"] / tot_i, ln[0][:44], ln[1], src[ln][:70],
        " ".join("%s=%d" % (n, v) for v, n in s if v)))
