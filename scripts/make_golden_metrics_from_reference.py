"""Generates tests/golden/reference_metrics.npz + reference_metrics.json by EXECUTING the reference's own metric code
(confignet/metrics/inception_distance.py: compute_FID, compute_KID; confignet/metrics/metrics.py: ControllabilityMetrics,
InceptionMetrics) from /root/reference.  scikit-learn, SciPy and OpenCV are installed here; TensorFlow and matplotlib are
stubbed, and the networks are replaced by the deterministic fakes of tests/metrics_fakes.py (model, attribute classifier,
feature extractor) - the code paths recorded here never touch a stub.  Run in the build container only; outputs committed.

    python scripts/make_golden_metrics_from_reference.py
"""
import importlib
import json
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"


class _Anything:
    def __init__(self, *a, **k): pass
    def __call__(self, *a, **k): return _Anything()
    def __getattr__(self, name): return _Anything
    def __enter__(self): return self
    def __exit__(self, *a): return False


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = self.__name__ + "." + name
        return sys.modules[sub] if sub in sys.modules else _Anything


for m in ["tensorflow", "tensorflow.keras", "tensorflow.keras.applications", "tensorflow.keras.applications.inception_v3",
          "matplotlib", "matplotlib.pyplot"]:
    parts = m.split(".")
    for i in range(1, len(parts) + 1):
        n = ".".join(parts[:i])
        if n not in sys.modules:
            sys.modules[n] = _Stub(n)
pkg = types.ModuleType("confignet"); pkg.__path__ = [os.path.join(REF, "confignet")]; sys.modules["confignet"] = pkg
mpkg = types.ModuleType("confignet.metrics"); mpkg.__path__ = [os.path.join(REF, "confignet", "metrics")]; sys.modules["confignet.metrics"] = mpkg
idist = importlib.import_module("confignet.metrics.inception_distance")
rmetrics = importlib.import_module("confignet.metrics.metrics")
rceleba = importlib.import_module("confignet.metrics.celeba_attribute_prediction")
import metrics_fakes as FK          # noqa: E402

arrays, meta = {}, {}
# ---- compute_FID / compute_KID on float32 features (what get_features returns) and float64 ones
r = np.random.RandomState(0)
for tag, (m, n, d, dt) in {"small": (40, 50, 16, np.float32), "wide": (30, 25, 96, np.float32), "f64": (20, 22, 8, np.float64)}.items():
    g = (r.standard_normal((m, d)) * 0.7 + 0.2).astype(dt)
    q = r.standard_normal((n, d)).astype(dt)
    arrays["feat_g_" + tag], arrays["feat_r_" + tag] = g, q
    meta["fid_" + tag] = float(idist.compute_FID(g, q))
    meta["kid_" + tag] = float(idist.compute_KID(g, q))


# ---- ControllabilityMetrics host logic around the fakes
class FakeClassifier(rceleba.CelebaAttributeClassifier):
    def __init__(self):
        self.config = {"predicted_attributes": list(FK.ATTRIBUTES), "input_shape": (8, 8, 3)}

    def predict_attributes(self, images):
        return FK.fake_predict_attributes(images)


imgs = np.random.RandomState(1).randint(0, 256, (5, 8, 8, 3)).astype(np.uint8)
arrays["contr_input_images"] = imgs
for iters in (0, 2):
    model = FK.FakeModel()
    cm = rmetrics.ControllabilityMetrics(model, FakeClassifier(), per_image_tuning_iters=iters)
    out_dir = tempfile.mkdtemp()
    md = {"training_step_number": [0]}
    cm.update_and_log_metrics(imgs, md, out_dir)
    with open(os.path.join(out_dir, "controllability_metrics.json")) as fp:
        meta["contr_json_iters%d" % iters] = json.load(fp)
    meta["contr_log_iters%d" % iters] = model.log
    meta["contr_keys_iters%d" % iters] = list(md.keys())
meta["config_names"] = [n for n, _ in rmetrics.ControllabilityMetricConfigs.all_configs()]

# ---- InceptionMetrics: the sample draw, two consecutive updates, the text table
ds = FK.FakeDataset()
np.random.seed(11)
im = rmetrics.InceptionMetrics.__new__(rmetrics.InceptionMetrics)
im.n_samples_for_metrics = 20
im.inception_feature_extractor = types.SimpleNamespace(get_features=FK.fake_inception_features)
idx = np.random.randint(0, ds.imgs.shape[0], 20)           # metrics.py:206, the line __init__ runs after building the network
im.gt_inception_features = ds.inception_features[idx]
meta["inception_next_draw"] = int(np.random.randint(0, 2 ** 31 - 1))
out_dir = tempfile.mkdtemp()
md = {"training_step_number": [0]}
gen = np.random.RandomState(2).randint(0, 256, (12, 8, 8, 3)).astype(np.uint8)
arrays["inception_generated"] = gen
im.update_and_log_metrics(gen, md, out_dir)
md["training_step_number"].append(1000)
im.update_and_log_metrics(gen[::-1] // 2, md, out_dir)
meta["inception_metrics_dict"] = {k: [float(v) for v in vals] for k, vals in md.items()}
with open(os.path.join(out_dir, "inception_metrics.txt")) as fp:
    meta["inception_metrics_txt"] = fp.read()

np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reference_metrics.npz"), **arrays)
with open(os.path.join(ROOT, "tests", "golden", "reference_metrics.json"), "w") as fp:
    json.dump(meta, fp)
print("wrote reference_metrics.{npz,json}:", {k: meta[k] for k in meta if k.startswith(("fid", "kid"))})
