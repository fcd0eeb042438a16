"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name.
Usage: python tools/launch_agg.py file.csv [skip_first_fraction]  — with two identical iterations in the capture,
pass 0.5 to keep only the second (warm) one."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = [(x["Kernel Name"], float(x["Metric Value"].replace(",", ""))) for x in csv.DictReader(lines)
        if x["Metric Name"] == "gpu__time_duration.sum"]
skip = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
rows = rows[int(len(rows) * skip):]
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in rows:
    k = n.split("(")[0][-70:]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v for _, v in rows)
print(f"{len(rows)} launches, {tot / 1000:.1f} us")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 22]:
    print(f"{v / 1000:9.1f} us {100 * v / tot:5.1f}% {c:5d}  {k}")
