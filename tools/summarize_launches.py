"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of total device time)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    try:
        t = float(row["Metric Value"].replace(",", ""))
    except (KeyError, ValueError):
        continue
    unit = row["Metric Unit"]
    t = t / 1e3 if unit == "ns" else t * 1e3 if unit == "ms" else t          # -> microseconds
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    agg[name][0] += 1
    agg[name][1] += t
tot = sum(v[1] for v in agg.values())
print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot/1e3:.2f} ms total device time (ncu: cold-cache, serialised)")
print(f"{'kernel':72s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:72]:72s} {v[0]:8d} {v[1]/1e3:10.2f} {v[1]/tot*100:6.1f}% {v[1]/v[0]:9.1f}")
