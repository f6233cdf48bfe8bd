"""Hot-code footprint of the align kernel from an ncu source page (ncu -i X.ncu-rep --page source --csv):
instructions executed >= THRESH times per launch, grouped into contiguous regions.  usage: hot_code.py src.csv [thresh]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 1e5
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]


def num(x):
    try:
        return float(x.replace(',', ''))
    except ValueError:
        return 0.0


tot = sum(num(r[ix["# Samples"]]) for r in data)
hot = [r for r in data if num(r[ix["Instructions Executed"]]) >= thresh]
print("instructions", len(data), "samples", tot, "| hot (>= %g exec): %d = %d bytes" % (thresh, len(hot), len(hot) * 16))
addrs = [int(r[ix["Address"]], 16) for r in data]
base = addrs[0]
regions, cur = [], None
for a, r in zip(addrs, data):
    e = num(r[ix["Instructions Executed"]])
    if e >= thresh:
        if cur and a - cur[1] <= 16 * 8:
            cur[1] = a; cur[2] += 1; cur[3] += num(r[ix["# Samples"]]); cur[4] = max(cur[4], e); cur[5] += num(r[ix["Instructions Executed"]])
        else:
            cur = [a, a, 1, num(r[ix["# Samples"]]), e, num(r[ix["Instructions Executed"]]), r[ix["Source"]][:60]]
            regions.append(cur)
print("hot regions:", len(regions))
for g in regions:
    if g[2] >= 24:
        print("  off %7d..%7d  n=%5d  samples %5.1f%%  exec %.2e  max %.2e  %s" % (g[0] - base, g[1] - base, g[2], 100 * g[3] / tot, g[5], g[4], g[6]))
