"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, re, sys
from collections import defaultdict
path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*$", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    agg[name][0] += 1
    agg[name][1] += us
    tot += us
print(f"total {tot/1000:.3f} ms over {sum(a[0] for a in agg.values())} launches")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{us/1000:9.3f} ms {100*us/tot:5.1f}%  x{n:4d}  avg {us/n:9.1f} us  {name[:110]}")
