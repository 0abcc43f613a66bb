"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown on stdout).
usage: python scripts/summarize_launches.py gpurun_out/launches.csv [first_id [count]]"""
import csv, re, sys
from collections import OrderedDict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
count = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 30
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("sb200::", "")
    rows.append((int(r["ID"]), name, r["Grid Size"], float(r["Metric Value"]) / 1e3))
rows = [r for r in rows if skip <= r[0] < skip + count]
agg = OrderedDict()
for _, name, grid, us in rows:
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += us
total = sum(a[1] for a in agg.values())
print(f"launches: {len(rows)}   total device time (serialised, cold cache): {total:.1f} us\n")
print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {n} | {us:.1f} | {100 * us / total:.1f} % |")
