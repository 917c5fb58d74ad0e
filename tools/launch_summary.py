"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv).  Usage: launch_summary.py file.csv"""
import csv, sys, re, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
ui = h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^void ", "", name)
    v = float(r[vi].replace(",", ""))
    if r[ui] == "ns": v /= 1000.0
    elif r[ui] == "ms": v *= 1000.0
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{sum(a[0] for a in agg.values())} launches, {tot / 1000:.2f} ms")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[1] / tot * 100:6.2f} %  {a[0]:6d} x {a[1] / a[0]:8.2f} us   {k[:110]}")
