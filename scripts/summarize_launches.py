"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total ms, share."""
import csv, sys, collections
path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = []
with open(path, newline="") as fp:
    lines = [l for l in fp if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = collections.OrderedDict()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[u]
    a = agg.setdefault(r["Kernel Name"][:100], [0, 0.0])
    a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
print("%-100s %9s %10s %7s" % ("kernel (all captured launches / %g steps)" % steps, "launches", "ms/step", "share"))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-100s %9.1f %10.3f %6.1f%%" % (k, n / steps, t / steps, 100 * t / tot))
print("%-100s %9.1f %10.3f" % ("TOTAL", sum(a[0] for a in agg.values()) / steps, tot / steps))
