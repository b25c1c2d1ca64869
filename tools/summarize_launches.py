"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel (+grid)."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
by_grid = len(sys.argv) > 2 and sys.argv[2] == "--grid"
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
    name = r["Kernel Name"].split("(")[0]
    key = (name, r.get("Grid Size", "")) if by_grid else (name,)
    rows.append((key, ns))
agg = defaultdict(lambda: [0, 0.0])
for k, ns in rows:
    agg[k][0] += 1
    agg[k][1] += ns
total = sum(v[1] for v in agg.values())
print(f"total kernel time {total / 1e6:.3f} ms over {len(rows)} launches")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{ns / 1e6:9.3f} ms {100 * ns / total:5.1f}%  n={n:5d}  avg {ns / n / 1e3:8.2f} us  {' '.join(k)}")
