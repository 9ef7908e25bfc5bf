"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
total = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v_us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    agg[name][0] += 1
    agg[name][1] += v_us
    total += v_us
print(f"{'kernel':60s} {'launches':>8s} {'total_us':>10s} {'share':>7s} {'avg_us':>8s}")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:60]:60s} {n:8d} {t:10.1f} {100 * t / total:6.1f}% {t / n:8.1f}")
print(f"{'TOTAL':60s} {sum(n for n, _ in agg.values()):8d} {total:10.1f}")
