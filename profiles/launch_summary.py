"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: python profiles/launch_summary.py file.csv"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    a = agg.setdefault(r[ki][:80], [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, a in agg.items():
    print("%5d launches  %9.2f us avg  %9.1f us total  %5.1f %%  %s" % (a[0], a[1] / a[0] / 1e3, a[1] / 1e3, 100 * a[1] / tot, k))
print("total %.1f us" % (tot / 1e3))
