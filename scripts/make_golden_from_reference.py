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

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


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
stub("cv2")
sys.modules["confignet.perceptual_loss"] = _Stub("confignet.perceptual_loss")
mpkg = types.ModuleType("confignet.metrics")
mpkg.__path__ = [os.path.join(REF, "confignet", "metrics")]
sys.modules["confignet.metrics"] = mpkg           # the REAL metrics.py: InceptionMetrics.__init__ draws from the NumPy stream
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
# ---- setup_training (confignet_first_stage.py:562-595, metrics.py:202-207): what it draws from the global NumPy
#      stream before the first step, and the checkpoint / metric inputs it builds; then the train() loop
#      (:597-626) with the step methods replaced by recorders: call order, optimizer sharing, loss history, checkpoint cadence
import tempfile
fs.InceptionMetrics.__init__.__globals__["InceptionFeatureExtractor"] = _Anything      # InceptionV3 + ImageNet weights
sys.modules["tensorflow"].summary = _Anything()                                        # TensorBoard writer: not on this path


def make_set(n, seed):
    r = np.random.RandomState(seed)
    ds = types.SimpleNamespace()
    ds.imgs = r.randint(0, 256, (n, 4, 4, 3)).astype(np.uint8)
    ds.eye_masks = np.zeros((n, 4, 4), np.uint8)
    ds.inception_features = r.rand(n, 5).astype(np.float32)
    ds.metadata_inputs = {k: r.rand(n, d[0]).astype(np.float32) for k, d in model.config["facemodel_inputs"].items()}
    ds.metadata_inputs["rotations"] = r.rand(n, 3).astype(np.float32)
    ds.metadata_input_distributions = {"tag": seed}
    return ds


real_set, synth_set = make_set(13, 41), make_set(11, 42)
tmp = tempfile.mkdtemp(prefix="cn_setup_")
np.random.seed(3)
model.setup_training(os.path.join(tmp, "log"), synth_set, 7, real_training_set=real_set)
ck, gm = model._checkpoint_visualization_input, model._generator_input_for_metrics
out["setup_training"] = {
    "datasets": {"real": [13, 41], "synth": [11, 42]}, "seed": 3, "n_samples_for_metrics": 7,
    "metric_latent": np.asarray(gm["latent"]).tolist(), "metric_rotation": np.asarray(gm["rotation"]).tolist(),
    "checkpoint_latent_shape": list(ck["latent"].shape), "checkpoint_latent_rows_0_10_59": ck["latent"][[0, 10, 59]].tolist(),
    "checkpoint_rotation": ck["rotation"].tolist(),
    "checkpoint_facemodel_param_shapes": [list(p_.shape) for p_ in ck["facemodel_params"]],
    "checkpoint_facemodel_param0": np.asarray(ck["facemodel_params"][0]).tolist(),
    "checkpoint_gt_imgs_sum_per_image": np.asarray(ck["gt_imgs"]).reshape(10, -1).sum(axis=1).tolist(),
    "distributions": model.facemodel_param_distributions,
    "next_draw_after_setup": int(np.random.randint(0, 2 ** 31 - 1)),
}

calls = []


def recorder(name, n_sets):
    def f(*args):
        sets, opt = args[:n_sets], args[n_sets]
        calls.append([name] + ["real" if s_ is real_set else "synth" for s_ in sets] + [opt.tag])
        return {"loss_sum": float(len(calls)), "extra": 0.5}
    return f


class _Opt:
    made = []

    def __init__(self, **kw):
        self.tag = "optimizer%d" % len(_Opt.made)
        self.kw = kw
        _Opt.made.append(self)


fs.keras.optimizers = types.SimpleNamespace(Adam=_Opt)
fs.time.clock = fs.time.perf_counter                        # removed from the time module in Python 3.8
model.discriminator_training_step = recorder("discriminator_training_step", 1)
model.synth_discriminator_training_step = recorder("synth_discriminator_training_step", 1)
model.latent_discriminator_training_step = recorder("latent_discriminator_training_step", 1)
model.generator_training_step = recorder("generator_training_step", 2)
model.update_smoothed_weights = lambda: calls.append(["update_smoothed_weights"])
model.run_checkpoints = lambda output_dir, iteration_time, aml_run=None: calls.append(
    ["run_checkpoints", model.get_training_step_number()])
model.config["n_discriminator_updates"], model.config["n_generator_updates"] = 2, 1
model.g_losses, model.d_losses, model.synth_d_losses, model.latent_d_losses = {}, {}, {}, {}
np.random.seed(3)
model.train(real_set, synth_set, tmp, os.path.join(tmp, "log"), n_steps=2, n_samples_for_metrics=7)
out["train_loop"] = {"calls": calls, "optimizer_kwargs": [o.kw for o in _Opt.made],
                     "g_losses": model.g_losses, "d_losses": model.d_losses, "synth_d_losses": model.synth_d_losses,
                     "latent_d_losses": model.latent_d_losses,
                     "next_draw_after_train": int(np.random.randint(0, 2 ** 31 - 1))}
# resuming: a model whose history already holds 2 steps asked for n_steps=3 runs one more iteration
n_before = len(calls)
model.train(real_set, synth_set, tmp, os.path.join(tmp, "log"), n_steps=3, n_samples_for_metrics=7)
out["train_loop"]["resumed_iterations"] = sum(1 for c in calls[n_before:] if c[0] == "update_smoothed_weights")
model.config["n_discriminator_updates"] = 1

# ---- ConfigNet (second stage) setup_training / train (confignet_second_stage.py:255-299): two more draws for the
#      validation rows, the latent discriminator step sees the real set too
pkg.ConfigNetFirstStage = fs.ConfigNetFirstStage             # what confignet/__init__.py exports
pkg.confignet_utils = utils
second = importlib.import_module("confignet.confignet_second_stage")
celeba = importlib.import_module("confignet.metrics.celeba_attribute_prediction")
m_s2 = second.ConfigNet(cfg, initialize=False)
val_set = make_set(9, 43)
np.random.seed(4)
m_s2.setup_training(os.path.join(tmp, "log2"), synth_set, 7, object.__new__(celeba.CelebaAttributeClassifier),
                    real_training_set=real_set, validation_set=val_set)
out["stage2_setup_training"] = {
    "seed": 4, "validation": [9, 43],
    "checkpoint_input_images_sum_per_image": ((m_s2._checkpoint_visualization_input["input_images"] + 1.0) * 127.5).reshape(10, -1).sum(axis=1).round().tolist(),
    "metric_input_images_sum_per_image": ((m_s2._generator_input_for_metrics["input_images"] + 1.0) * 127.5).reshape(7, -1).sum(axis=1).round().tolist(),
    "next_draw_after_setup": int(np.random.randint(0, 2 ** 31 - 1)),
}
calls2 = []


def recorder2(name, n_sets):
    def f(*args):
        sets, opt = args[:n_sets], args[n_sets]
        calls2.append([name] + ["real" if s_ is real_set else "synth" for s_ in sets] + [opt.tag])
        return {"loss_sum": float(len(calls2))}
    return f


second.keras.optimizers = types.SimpleNamespace(Adam=_Opt)
_Opt.made.clear()
m_s2.discriminator_training_step = recorder2("discriminator_training_step", 1)
m_s2.synth_discriminator_training_step = recorder2("synth_discriminator_training_step", 1)
m_s2.latent_discriminator_training_step = recorder2("latent_discriminator_training_step", 2)
m_s2.generator_training_step = recorder2("generator_training_step", 2)
m_s2.update_smoothed_weights = lambda: calls2.append(["update_smoothed_weights"])
m_s2.run_checkpoints = lambda output_dir, iteration_time, aml_run=None: calls2.append(["run_checkpoints", m_s2.get_training_step_number()])
m_s2.train(real_set, synth_set, val_set, object.__new__(celeba.CelebaAttributeClassifier), tmp, os.path.join(tmp, "log2"),
           n_steps=2, n_samples_for_metrics=7)
out["stage2_train_loop"] = {"calls": calls2}

# ---- LatentGAN.setup_logs / train (latent_gan.py:200-247): draws before the first step, loop order, checkpoint cadence
lg = importlib.import_module("confignet.latent_gan")
lg.InceptionMetrics.__init__.__globals__["InceptionFeatureExtractor"] = _Anything
gan = lg.LatentGAN({"latent_dim": 145, "n_samples_for_metrics": 9, "verbose_log_period": 2})
np.random.seed(5)
gan.setup_logs(os.path.join(tmp, "gan_log"), real_set, model)
out["latent_gan_setup"] = {
    "seed": 5, "log_latents_shape": list(gan.inputs_for_logs["latents"].shape),
    "log_latents_rows_0_35": gan.inputs_for_logs["latents"][[0, 35]].tolist(),
    "log_rotations_all_zero": bool((gan.inputs_for_logs["rotations"] == 0).all()),
    "metric_latents": gan.inputs_for_metrics["latents"].tolist(), "metric_rotations": gan.inputs_for_metrics["rotations"].tolist(),
    "next_draw_after_setup": int(np.random.randint(0, 2 ** 31 - 1)),
}
gcalls = []
gan.extract_embeddings = lambda confignet_model, training_set: gcalls.append(["extract_embeddings"]) or "embeddings"
gan.discriminator_training_step = lambda emb, opt: gcalls.append(["discriminator_training_step", emb, opt.tag]) or {"loss_sum": 1.0}
gan.generator_training_step = lambda opt: gcalls.append(["generator_training_step", opt.tag]) or {"loss_sum": 2.0}
gan.update_smoothed_weights = lambda: gcalls.append(["update_smoothed_weights"])
gan.save = lambda d, name: gcalls.append(["save", os.path.relpath(d, tmp), name])
gan.generator_smoothed = types.SimpleNamespace(predict=lambda x: np.zeros((np.asarray(x).shape[0], 145), np.float32))
gan._inception_metric_object_cls = None
lg.keras.optimizers = types.SimpleNamespace(Adam=_Opt)
_Opt.made.clear()
fake_model = types.SimpleNamespace(config=model.config, sample_rotations=model.sample_rotations,
                                   generate_images=lambda e, r: np.zeros((np.asarray(e).shape[0], 4, 4, 3), np.uint8))
lg.InceptionMetrics.get_metrics = lambda self, imgs: (0.0, 0.0)
lg.confignet_utils.build_image_matrix = lambda imgs, r, c: np.zeros((4, 4, 3), np.uint8)
gan.train(real_set, fake_model, tmp, os.path.join(tmp, "gan_log"), 3)
out["latent_gan_train"] = {"calls": gcalls, "optimizer_kwargs": [o.kw for o in _Opt.made]}

# ---- golden .npz shapes shipped with the reference's tests
shapes = {}
for f in ["confignet_basic_ref_256", "confignet_basic_ref_512", "confignet_finetune_ref_256", "latentgan_ref_256"]:
    z = np.load(os.path.join(REF, "tests", "test_assets", f + ".npz"))
    shapes[f] = {k: [list(z[k].shape), str(z[k].dtype)] for k in z.files}
out["reference_golden_npz_shapes"] = shapes

with open(OUT, "w") as fp:
    json.dump(out, fp, indent=1, sort_keys=True)
print("wrote", OUT, "latent_dim", out["latent_dim"], out["layout2_latent_dim"])
