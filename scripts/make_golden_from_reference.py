"""Generates tests/golden/reference_host_logic.json by EXECUTING the reference's own pure-Python host
logic (latent layout, config merging, samplers, image flips) from /root/reference with TensorFlow and
the other missing third-party imports stubbed out.  Only code paths that never touch a stubbed module
are called, so the outputs are the reference's real outputs.  Run in the build container only
(/root/reference does not exist on the GPU box); the JSON is committed.

    python scripts/make_golden_from_reference.py
"""
import importlib
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "reference_host_logic.json")


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = self.__name__ + "." + name
        if sub in sys.modules:
            return sys.modules[sub]
        return _Anything


def stub(name):
    parts = name.split(".")
    for i in range(1, len(parts) + 1):
        n = ".".join(parts[:i])
        if n not in sys.modules:
            sys.modules[n] = _Stub(n)


for m in ["tensorflow", "tensorflow.keras", "tensorflow.keras.layers", "tensorflow.keras.models",
          "tensorflow.keras.backend", "tensorflow.keras.utils", "tensorflow.keras.applications",
          "tensorflow.keras.initializers", "tensorflow.keras.regularizers", "tensorflow.keras.constraints",
          "matplotlib", "matplotlib.pyplot", "transformations", "azureml", "azureml.core",
          "azureml.core.run", "sklearn.mixture", "scipy.linalg"]:
    if m.split(".")[0] in ("sklearn", "scipy"):
        continue
    stub(m)

sys.path.insert(0, REF)
# the package __init__ imports everything; import the submodules we need directly
pkg = types.ModuleType("confignet")
pkg.__path__ = [os.path.join(REF, "confignet")]
sys.modules["confignet"] = pkg
for heavy in ["confignet.perceptual_loss", "confignet.metrics", "confignet.metrics.metrics"]:
    sys.modules[heavy] = _Stub(heavy)
utils = importlib.import_module("confignet.confignet_utils")
fs = importlib.import_module("confignet.confignet_first_stage")

TEST_DIMS = {"beard_style_embedding": 9, "blendshape_values": 62, "bone_rotations:left_eye": 3, "eye_color": 4,
             "eyebrow_style_embedding": 44, "geometry_identity_params": 53, "hdri_embedding": 50,
             "head_hair_color": 3, "head_hair_style_embedding": 18, "lower_eyelash_style": 3,
             "texture_embedding": 50, "upper_eyelash_style": 3}

out = {}
# ---- config merging + latent layout, default config with the test-dataset input dims
cfg = {"output_shape": (256, 256, 3), "batch_size": 4,
       "facemodel_inputs": {k: (TEST_DIMS[k], v[1]) for k, v in fs.DEFAULT_CONFIG["facemodel_inputs"].items()}}
model = fs.ConfigNetFirstStage(cfg, initialize=False)
out["latent_dim"] = model.config["latent_dim"]
out["facemodel_inputs_sorted"] = [[k, list(v)] for k, v in model.config["facemodel_inputs"].items()]
out["facemodel_input_dim"] = model.facemodel_input_dim
out["latent_idxs"] = {k: [model.get_facemodel_param_idxs_in_latent(k).start, model.get_facemodel_param_idxs_in_latent(k).stop]
                      for k in model.config["facemodel_inputs"]}
out["merged_scalar_keys"] = {k: model.config[k] for k in ["batch_size", "latent_regression_weight", "n_discr_layers",
                                                          "image_loss_weight", "eye_loss_weight", "model_type"]}
out["merged_optimizer"] = model.config["optimizer"]
# a second layout: released 256 model has latent 144 (one parameter without an input dim is dropped)
cfg2 = {"facemodel_inputs": {k: ((TEST_DIMS[k] if k != "head_hair_color" else None), v[1])
                             for k, v in fs.DEFAULT_CONFIG["facemodel_inputs"].items()}}
m2 = fs.ConfigNetFirstStage(cfg2, initialize=False)
out["layout2_latent_dim"] = m2.config["latent_dim"]
out["layout2_idxs"] = {k: [m2.get_facemodel_param_idxs_in_latent(k).start, m2.get_facemodel_param_idxs_in_latent(k).stop]
                       for k in m2.config["facemodel_inputs"]}
# ---- merge_configs corner cases
out["merge_cases"] = []
for d, i in [({"a": 1, "b": {"c": 2, "d": 3}}, {"b": {"c": 5}, "e": 7}), ({"a": {"x": 1}}, {"a": {"y": 2}}), ({"a": 1}, {})]:
    out["merge_cases"].append({"default": d, "input": i, "result": utils.merge_configs(d, i)})
# ---- samplers (NumPy global RNG, seed 0 as training_utils.py:8-11)
np.random.seed(0)
out["sample_rotations_seed0_n5"] = model.sample_rotations(5).tolist()
out["sample_latent_seed0_after_rot_n2"] = model.sample_latent_vector(2).tolist()
np.random.seed(0)
imgs = np.arange(4 * 2 * 3 * 3, dtype=np.float32).reshape(4, 2, 3, 3)
flipped = utils.flip_random_subset_of_images(imgs.copy())
out["flip_seed0_input_shape"] = list(imgs.shape)
out["flip_seed0_result"] = flipped.tolist()
# ---- set_facemodel_param_in_latents column arithmetic with a fake per-parameter encoder
class _FakeMLP:
    def predict(self, v):
        return np.full((v.shape[0], 30), 9.0, np.float32)
class _FakeEnc:
    per_facemodel_input_mlps = {"blendshape_values": _FakeMLP()}
model.synthetic_encoder = _FakeEnc()
lat = np.zeros((2, model.config["latent_dim"]), np.float32)
new = model.set_facemodel_param_in_latents(lat, "blendshape_values", np.zeros(62, np.float32))
out["set_param_changed_columns"] = np.nonzero(new[0])[0].tolist()
out["set_param_input_untouched"] = bool((lat == 0).all())
# ---- golden .npz shapes shipped with the reference's tests
shapes = {}
for f in ["confignet_basic_ref_256", "confignet_basic_ref_512", "confignet_finetune_ref_256", "latentgan_ref_256"]:
    z = np.load(os.path.join(REF, "tests", "test_assets", f + ".npz"))
    shapes[f] = {k: [list(z[k].shape), str(z[k].dtype)] for k in z.files}
out["reference_golden_npz_shapes"] = shapes

with open(OUT, "w") as fp:
    json.dump(out, fp, indent=1, sort_keys=True)
print("wrote", OUT, "latent_dim", out["latent_dim"], out["layout2_latent_dim"])
