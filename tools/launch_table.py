"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv): usage launch_table.py launches.csv"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
first = {}
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except (ValueError, KeyError):
        continue
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1e-6)
    k = row["Kernel Name"][:64]
    a = agg.setdefault(k, [0, 0.0, 0.0]); a[0] += 1; a[1] += v; a[2] = max(a[2], v)
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-66s n=%4d total %9.3f ms  max %8.3f ms  %5.1f%%" % (k, a[0], a[1], a[2], 100 * a[1] / tot))
