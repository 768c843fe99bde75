"""Writes confignet_b200/metrics/controllability_tables.json from the reference's two DATA modules
(metrics/controllability_metric_configs.py: the eight attribute configurations; metrics/blendshape_names.py: the
face model's blendshape order), loaded by file path (both are pure Python without third-party imports).
The values are parameters of the metric, not code; run in the build container only.

    python scripts/make_metric_tables_from_reference.py
"""
import importlib.util
import json
import os

REF = "/root/reference/confignet/metrics"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "confignet_b200", "metrics",
                   "controllability_tables.json")


def load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


cfgs = load("controllability_metric_configs").ControllabilityMetricConfigs.all_configs()
names = load("blendshape_names").blendshape_names
out = {"blendshape_names": list(names),
       "configs": [[name, dict(c._asdict())] for name, c in cfgs]}        # all_configs(): sorted by attribute name
with open(OUT, "w") as fp:
    json.dump(out, fp, indent=1)
print("wrote", OUT, len(cfgs), "configs,", len(names), "blendshape names")
