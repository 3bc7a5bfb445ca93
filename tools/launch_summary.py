"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
    python tools/launch_summary.py gpurun_out/launches.csv [skip_first_n]"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if l.startswith('"')]
rd = csv.DictReader(lines)
for r in rd:
    if r["Metric Name"] == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in rows:
    k = re.sub(r"\(.*", "", k)
    k = re.sub(r"^void ", "", k)[:90]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot / 1e6:.3f} ms total")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{t / 1e6:9.3f} ms {100 * t / tot:5.1f}%  x{n:5d}  {k}")
