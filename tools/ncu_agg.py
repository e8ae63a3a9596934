"""Aggregate an ncu --csv launch list (gpu__time_duration.sum) per kernel name: total us, launches."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    a = agg.setdefault(r[ki][:90], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(t for _, t in agg.values())
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{t / 1e3:10.1f} us {100 * t / tot:5.1f}% {c:5d}  {k}")
