"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name, launches and total / mean
duration, for the LAST `--launches` launches in the file (one training step)."""
import argparse
import csv
import re
from collections import defaultdict

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--launches", type=int, default=0, help="0 = all")
args = ap.parse_args()
rows = []
with open(args.csv) as fh:
    lines = [l for l in fh if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
for r in rd:
    if r[im] != "gpu__time_duration.sum":
        continue
    rows.append((r[ik], float(r[iv].replace(",", "")) / 1e3))      # ns -> us
if args.launches:
    rows = rows[-args.launches:]
agg = defaultdict(lambda: [0, 0.0, 0.0])
for k, us in rows:
    k = re.sub(r"\(.*", "", k).replace("void ", "").replace("alignn::", "")[:70]
    a = agg[k]
    a[0] += 1
    a[1] += us
    a[2] = max(a[2], us)
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {tot / 1e3:.3f} ms of kernel time (serialised, cold-cache: compare shares)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[1] / tot * 100:6.2f}%  {a[1]:9.1f} us  n={a[0]:4d}  mean {a[1] / a[0]:7.1f}  max {a[2]:7.1f}  {k}")
