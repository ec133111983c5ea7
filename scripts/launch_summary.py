"""Summarises an ncu --csv launch list (gpu__time_duration.sum) per kernel name."""
import collections
import csv
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.reader(lines); hdr = next(r)
ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict(); tot = 0.0
for row in r:
    if len(row) <= vi:
        continue
    v = float(row[vi].replace(",", "")); u = row[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    a = agg.setdefault(row[ki][:100], [0.0, 0]); a[0] += ms; a[1] += 1; tot += ms
for n, (ms, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print("%8.3f ms %5.1f%% x%-3d %s" % (ms, 100 * ms / tot, c, n))
print("total %.3f ms over %d launches" % (tot, sum(c for _, c in agg.values())))
