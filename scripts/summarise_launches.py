"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel."""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(float); cnt = defaultdict(int)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    name = r["Kernel Name"].split("(")[0]
    tot[name] += v * scale; cnt[name] += 1
all_us = sum(tot.values())
print(f"{'kernel':28s} {'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share':>7s}")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"{k:28s} {cnt[k]:8d} {tot[k]:12.1f} {tot[k]/cnt[k]:10.1f} {100*tot[k]/all_us:6.1f}%")
print(f"{'TOTAL':28s} {sum(cnt.values()):8d} {all_us:12.1f}")
