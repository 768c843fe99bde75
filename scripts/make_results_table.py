"""Markdown results table from the bench lines of one measurement visit.

    python scripts/make_results_table.py profiles/r02_final      # reads <prefix>_bench.json, _bench_cfg{1,3,4,5}.json,
                                                                 # _bench_reference.json, _bench_n2.json when present
"""
import json, os, sys

prefix = sys.argv[1]


def load(path):
    if not os.path.exists(path):
        return None
    with open(path) as fp:
        for line in fp:
            line = line.strip()
            if line.startswith("{"):
                return json.loads(line)
    return None


rows = [("2 (headline)", load(prefix + "_bench.json"))] + [(str(c), load("%s_bench_cfg%d.json" % (prefix, c))) for c in (1, 3, 4, 5)]
n2 = load(prefix + "_bench_n2.json")
if n2:
    rows.append(("2 @ 2 GPUs", n2))
print("| config | workload | GPUs | value | ms/step | e2e | tcgen05 conv TFLOP/s (algorithmic / executed) | frac of bf16 peak | tc ms/step | CUDA-core conv ms/step | CPU oracle (cores) |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for name, d in rows:
    if d is None:
        continue
    r, e, c = d.get("roofline") or {}, d.get("e2e") or {}, d.get("cpu_baseline") or {}
    wl = d["config"].get("workload", "")[:70]
    print("| %s | %s | %d | %.1f %s | %.2f | %s | %s | %s | %s | %s | %s |" % (
        name, wl, d["n_gpus"], d["value"], d["unit"], d["ms_per_step"],
        ("%.1f" % e["value"]) if e.get("value") else "-",
        ("%.1f / %.1f" % (r["achieved"], r.get("executed_tflops", 0.0))) if r.get("achieved") else "-",
        ("%.3f" % r["frac"]) if r.get("frac") else "-",
        ("%.2f" % r["tc_ms_per_step"]) if r.get("tc_ms_per_step") else "-",
        ("%.2f" % r["cuda_core_conv_ms_per_step"]) if r.get("cuda_core_conv_ms_per_step") else "-",
        ("%.2f %s (%d)" % (c["value"], c.get("unit", ""), c.get("cores", 0))) if c.get("value") else "-"))
ref = load(prefix + "_bench_reference.json")
if ref:
    print("\nReference arm (`bench.py --impl reference`, %s): %.2f %s, %.0f ms per step, %s cores - %s" % (
        ref.get("cpu_baseline", {}).get("kind", "port"), ref["value"], ref["unit"], ref["ms_per_step"],
        ref.get("cpu_baseline", {}).get("cores", "?"), ref.get("cpu_baseline", {}).get("sample", "")))
head = rows[0][1]
if head and head.get("clocks"):
    print("\nClocks during the timed region: %s" % json.dumps(head["clocks"]))
