"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total ms, share."""
import csv, sys, collections
path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = []
with open(path, newline="") as fp:
    lines = [l for l in fp if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = collections.OrderedDict()
dram = collections.OrderedDict()       # kernel -> [launch ids, bytes] when dram__bytes_* were captured too
for r in rd:
    if r.get("Metric Name") in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1.0)
        d = dram.setdefault(r["Kernel Name"][:100], [set(), 0.0])
        d[0].add(r["ID"]); d[1] += v
        continue
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

if dram:
    import json, os
    tc = [(k, v) for k, v in dram.items() if "igemm_tc_pixel_kernel" in k]
    n = sum(len(v[0]) for _, v in tc); b = sum(v[1] for _, v in tc)
    if n:
        out = {"dram_bytes_per_launch": b / n, "launches": n, "kernel": "igemm_tc_pixel_kernel<*>",
               "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum over the tcgen05 conv launches of one bench "
                         "invocation (%s), averaged per launch" % os.path.basename(path)}
        print("tcgen05 conv kernels: %.2f MB DRAM traffic per launch over %d launches" % (b / n / 1e6, n))
        if len(sys.argv) > 3:
            json.dump(out, open(sys.argv[3], "w"))
