"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
total = 0.0
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    u = r[ui]
    v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)  # -> us
    name = re.sub(r"<.*", "", r[ki])[:90]
    agg[name][0] += 1
    agg[name][1] += v
    total += v
print(f"total {total/1e3:.3f} ms over {sum(a[0] for a in agg.values())} launches")
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{t/1e3:9.3f} ms {100*t/total:5.1f}%  x{n:<5d} {name}")
