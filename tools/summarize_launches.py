"""Summarise an ncu launch-list CSV (gpu__time_duration.sum per launch): totals per kernel + per-launch list."""
import collections
import csv
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
tot, cnt = collections.defaultdict(float), collections.Counter()
for row in rows:
    k = row["Kernel Name"][:44]
    tot[k] += float(row["Metric Value"].replace(",", ""))
    cnt[k] += 1
total = sum(tot.values())
print("total %.1f us over %d launches" % (total / 1e3, len(rows)))
for k in sorted(tot, key=lambda k: -tot[k]):
    print("%-46s %4d %10.1f us %5.1f%%" % (k, cnt[k], tot[k] / 1e3, 100 * tot[k] / total))
if len(sys.argv) > 2:
    i = 0
    for row in rows:
        if "igemm" in row["Kernel Name"] or "conv_halo" in row["Kernel Name"]:
            print("%3d %-9s %-14s %8.1f" % (i, row["Kernel Name"][:9], row["Grid Size"], float(row["Metric Value"].replace(",", "")) / 1e3), end=" | ")
            i += 1
            if i % 4 == 0:
                print()
    print()
