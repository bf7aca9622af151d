"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    agg[row["Kernel Name"].split("(")[0]].append(float(row["Metric Value"].replace(",", "")) / 1e3)
tot = sum(sum(v) for v in agg.values())
print("kernel, launches, total_us, share_pct, first_launches_us")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%s, %d, %.1f, %.1f, %s" % (k, len(v), sum(v), 100 * sum(v) / tot, [round(x, 1) for x in v[:6]]))
