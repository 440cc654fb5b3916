"""Per-kernel share of the step from an ncu launch list (gpu__time_duration.sum per launch, serialised, cold cache).
python tools/launch_shares.py profiles/r01_launches.csv"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
t, n = collections.defaultdict(float), collections.Counter()
for r in rows:
    if r["Metric Name"] == "gpu__time_duration.sum":
        k = r["Kernel Name"].split("(")[0].replace("void ", "")
        t[k] += float(r["Metric Value"].replace(",", ""))
        n[k] += 1
tot = sum(t.values())
print(f"{len(rows)} launches captured, {tot / 1e6:.3f} ms of kernel time")
for k, v in sorted(t.items(), key=lambda kv: -kv[1]):
    print(f"{100 * v / tot:5.1f} %  {v / n[k] / 1e3:9.1f} us x {n[k]:3d}  {k[:100]}")
