"""Generates tests/golden/reference_float_logic.npz by EXECUTING the reference's own floating-point helper code
from /root/reference - the 3-D resampler and the Euler-angle matrix (confignet_utils.py:63-145), the layer-style
statistics (:147-159), the GAN / eye / R1 / latent-regression loss formulas (losses.py:7-18,75-90), the
batch-normalised regression loss (confignet_second_stage.py:93-107) and InstanceNormalization.call
(dnn_models/instance_normalization.py:108-131) - with `tensorflow` replaced by a small NumPy-backed shim (float64)
that implements exactly the tf / keras.backend functions those bodies call.  TensorFlow 2.1 is not installable
here, so this pins the STRUCTURE of the reference code (index arithmetic, interpolation order, where each epsilon
sits, reduction axes, loss weights) - the elementary tf ops themselves are restated by the shim, each in one line.
The oracle (oracle/*.py, an independent torch restatement) is checked against these vectors in
tests/test_host_cpu.py::test_oracle_matches_reference_float_code.

Run in the build container only (/root/reference does not exist on the GPU box); the .npz is committed.

    python scripts/make_golden_float_from_reference.py
"""
import importlib
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "reference_float_logic.npz")
F = np.float64


# ------------------------------------------------------------------------------------------------ the shim
def _a(x):
    return np.asarray(x)


tf = types.ModuleType("tensorflow")
tf.float32, tf.int32 = "float32", "int32"
tf.cast = lambda x, dtype: _a(x).astype(np.int64 if dtype == "int32" else F)      # float32 -> float64: gold precision
tf.convert_to_tensor = lambda x, dtype=None: tf.cast(x, dtype) if dtype else _a(x)
tf.shape = lambda x: _a(x).shape
tf.tile = lambda x, reps: np.tile(_a(x), tuple(int(r) for r in reps))
tf.expand_dims = lambda x, axis: np.expand_dims(_a(x), axis)
tf.matmul = lambda a, b: np.matmul(_a(a), _a(b))
tf.transpose = lambda x, perm: np.transpose(_a(x), perm)
tf.clip_by_value = lambda x, lo, hi: np.clip(_a(x), lo, hi)
tf.reshape = lambda x, shape: np.reshape(_a(x), tuple(int(s) for s in shape))
tf.floor = lambda x: np.floor(_a(x))
tf.range = lambda n: np.arange(int(n))
tf.stack = lambda xs, axis=0: np.stack([_a(x) for x in xs], axis=axis)
tf.gather_nd = lambda params, idx: _a(params)[tuple(_a(idx).astype(np.int64).T)]
tf.sin, tf.cos, tf.sqrt, tf.square = np.sin, np.cos, np.sqrt, np.square
tf.squeeze = lambda x, axis=None: np.squeeze(_a(x), axis=axis)
tf.reduce_mean = lambda x, axis=None, keepdims=False: np.mean(_a(x), axis=tuple(axis) if isinstance(axis, (list, range)) else axis, keepdims=keepdims)
tf.reduce_sum = lambda x, axis=None, keepdims=False: np.sum(_a(x), axis=tuple(axis) if isinstance(axis, (list, range)) else axis, keepdims=keepdims)
tf.concat = lambda xs, axis: np.concatenate([_a(x) for x in xs], axis=axis)
tf.ones = lambda shape, dtype=None: np.ones(shape, F)
tf.zeros = lambda shape, dtype=None: np.zeros(shape, F)
tf.math = types.SimpleNamespace(softplus=lambda x: np.logaddexp(0.0, _a(x)),
                                reduce_variance=lambda x, axis=None, keepdims=False: np.var(_a(x), axis=axis, keepdims=keepdims))
tf.losses = types.SimpleNamespace(mean_squared_error=lambda a, b: np.mean(np.square(_a(a) - _a(b)), axis=-1))

K = types.ModuleType("tensorflow.keras.backend")
K.mean = lambda x, axis=None, keepdims=False: np.mean(_a(x), axis=tuple(axis) if isinstance(axis, list) else axis, keepdims=keepdims)
K.std = lambda x, axis=None, keepdims=False: np.std(_a(x), axis=tuple(axis) if isinstance(axis, list) else axis, keepdims=keepdims)
K.int_shape = lambda x: tuple(_a(x).shape)
K.reshape = lambda x, shape: np.reshape(_a(x), shape)


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


keras = _Stub("tensorflow.keras")
keras.backend = K
tf.keras = keras
sys.modules["tensorflow"] = tf
sys.modules["tensorflow.keras"] = keras
sys.modules["tensorflow.keras.backend"] = K
for m in ["tensorflow.keras.layers", "tensorflow.keras.models", "tensorflow.keras.utils", "tensorflow.keras.applications",
          "tensorflow.keras.initializers", "tensorflow.keras.regularizers", "tensorflow.keras.constraints",
          "matplotlib", "matplotlib.pyplot", "transformations", "azureml", "azureml.core", "azureml.core.run"]:
    parts = m.split(".")
    for i in range(1, len(parts) + 1):
        n = ".".join(parts[:i])
        if n not in sys.modules:
            sys.modules[n] = _Stub(n)

sys.path.insert(0, REF)
pkg = types.ModuleType("confignet")
pkg.__path__ = [os.path.join(REF, "confignet")]
sys.modules["confignet"] = pkg
for heavy in ["confignet.perceptual_loss", "confignet.metrics", "confignet.metrics.metrics"]:
    sys.modules[heavy] = _Stub(heavy)
utils = importlib.import_module("confignet.confignet_utils")
losses = importlib.import_module("confignet.losses")
first = importlib.import_module("confignet.confignet_first_stage")
pkg.ConfigNetFirstStage = first.ConfigNetFirstStage          # what confignet/__init__.py would have exported
pkg.confignet_utils = utils
second = importlib.import_module("confignet.confignet_second_stage")
inorm = importlib.import_module("confignet.dnn_models.instance_normalization")

rng = np.random.RandomState(7)
out = {}

# ---- euler_angles_to_matrix + transform_3d_grid_tf (confignet_utils.py:63-145)
angles = np.array([[0.3, -0.12, 0.0], [-0.5, 0.17, 0.05], [0.0, 0.0, 0.0], [0.9, -0.4, 0.3]], F)
out["euler_angles"] = angles
out["euler_matrix"] = utils.euler_angles_to_matrix(angles)
grid = rng.randn(4, 6, 6, 6, 3)
out["rot_grid"] = grid
out["rot_out"] = utils.transform_3d_grid_tf(grid, out["euler_matrix"])

# ---- get_layer_style (confignet_utils.py:147-159)
f4 = rng.randn(2, 5, 4, 6)
f5 = rng.randn(2, 3, 4, 5, 2)
for name, f in (("style4", f4), ("style5", f5)):
    mean, std = utils.get_layer_style(f)
    out[name + "_in"], out[name + "_mean"], out[name + "_std"] = f, mean, std

# ---- losses.py:7-18
scores = rng.randn(6, 1) * 2
out["scores"] = scores
out["gan_g_loss"] = losses.GAN_G_loss(scores)
out["gan_d_loss_ones"] = losses.GAN_D_loss(np.ones((6, 1)), scores)
out["gan_d_loss_zeros"] = losses.GAN_D_loss(np.zeros((6, 1)), scores)
gt, gen = rng.rand(3, 8, 8, 3) * 2 - 1, rng.rand(3, 8, 8, 3) * 2 - 1
masks = (rng.rand(3, 8, 8) < 0.3).astype(np.uint8)
out["eye_gt"], out["eye_gen"], out["eye_masks"] = gt, gen, masks
out["eye_loss"] = losses.eye_loss(gt, gen, masks)


# ---- gradient_regularization (losses.py:75-82): the tape is replaced by one that returns a given gradient
class _Tape:
    def __init__(self, g):
        self.g = g

    def gradient(self, out_, in_):
        return self.g


g_in = rng.randn(3, 8, 8, 3)
out["r1_grad"] = g_in
out["r1_penalty"] = losses.gradient_regularization(_Tape(g_in), None, None)

# ---- compute_latent_regression_loss (losses.py:85-90) with a regressor that returns a given tensor
labels, reg_out = rng.randn(5, 148), rng.randn(5, 148)
out["lr_labels"], out["lr_out"] = labels, reg_out
out["latent_regression_loss"] = losses.compute_latent_regression_loss(None, labels, lambda _: reg_out)

# ---- compute_normalized_latent_regression_loss (confignet_second_stage.py:93-107), unbound on a stand-in self
fake_self = types.SimpleNamespace(latent_regressor=lambda _: reg_out, config={"latent_regression_weight": 10.0})
out["normalized_latent_regression_loss"] = second.ConfigNet.compute_normalized_latent_regression_loss(fake_self, None, labels)

# ---- InstanceNormalization.call (instance_normalization.py:108-131), unbound on a stand-in self (axis=-1 as in
#      building_blocks.py:93)
x = rng.randn(2, 6, 5, 4) * 3 + 1
gamma, beta = rng.rand(4) + 0.5, rng.randn(4)
layer = types.SimpleNamespace(axis=-1, epsilon=1e-3, scale=True, center=True, gamma=gamma, beta=beta)
out["in_x"], out["in_gamma"], out["in_beta"] = x, gamma, beta
out["in_out"] = inorm.InstanceNormalization.call(layer, x)

np.savez_compressed(OUT, **{k: np.asarray(v, F if np.asarray(v).dtype.kind == "f" else None) for k, v in out.items()})
print("wrote", OUT, {k: np.asarray(v).shape for k, v in out.items()})
