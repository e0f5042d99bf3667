"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / mean / share."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    if "spin_kernel" in r[ki]:
        continue  # bench.py parks the GPU on torch.cuda._sleep in its instrumented pass; not part of a training step
    k = r[ki][:90]
    v = float(r[vi].replace(",", ""))
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("%-92s %6s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%-92s %6d %12.1f %10.1f %7.3f" % (k, a[0], a[1] / 1e3, a[1] / a[0] / 1e3, a[1] / tot))
