"""KID and FID (reference metrics/inception_distance.py).  The 2048 InceptionV3 features per image come from the B200
kernels (nets.inception_v3_features: every conv of the network is one implicit-GEMM launch with folded BatchNorm + ReLU);
the two distances are the reference's NumPy / SciPy formulas on the host - 1000 x 2048 features every 1000 training steps."""
import os
import warnings
import numpy as np
import scipy.linalg
import torch

from .. import ops
from ..runtime import ParamGroup
from . import nets

DEVICE_CHUNK = 128          # images per device pass (activations of the stem: 127 x 127 x 32 floats per image)
ENV_WEIGHTS = "CONFIGNET_INCEPTION_WEIGHTS"


def _sorted_by_layer_number(names, prefix):
    """keras numbers layers with a process-wide counter: conv2d_95 ... may be the first conv of this model"""
    own = [n for n in names if n.startswith(prefix)]
    num = lambda n: int(n[len(prefix):].split("/")[0].lstrip("_") or 0)
    return sorted(set(n.split("/")[0] for n in own), key=lambda n: num(n))


def load_inception_arrays(path_or_dict):
    """-> raw Keras variables under nets.inception_v3_spec() names.  Accepts an .npz / dict keyed by Keras variable names
    ('conv2d_7/kernel', 'batch_normalization_7/beta', ... with whatever layer numbers the exporting process had; they are
    renumbered from 1 in layer-number order = creation order), e.g. written on a machine with TensorFlow by
        m = tf.keras.applications.InceptionV3(include_top=False, weights="imagenet", pooling="avg")
        np.savez("inception_v3.npz", **{w.name.split(":")[0]: w.numpy() for w in m.weights})"""
    z = path_or_dict if isinstance(path_or_dict, dict) else dict(np.load(path_or_dict, allow_pickle=True))
    spec = nets.inception_v3_spec()
    if all(k in z for k in spec):
        return {k: np.asarray(z[k], np.float32) for k in spec}
    convs = _sorted_by_layer_number(z.keys(), "conv2d")
    bns = _sorted_by_layer_number(z.keys(), "batch_normalization")
    if len(convs) != 94 or len(bns) != 94:
        raise ValueError("InceptionV3 weights: expected 94 conv2d and 94 batch_normalization layers, got %d / %d" % (len(convs), len(bns)))
    out = {}
    for i, (c, b) in enumerate(zip(convs, bns), start=1):
        out["conv2d_%d/kernel" % i] = z[c + "/kernel"]
        for v in ("beta", "moving_mean", "moving_variance"):
            out["batch_normalization_%d/%s" % (i, v)] = z[b + "/" + v]
    for k, (shape, _) in spec.items():
        if tuple(out[k].shape) != tuple(shape):
            raise ValueError("InceptionV3 weights: %s has shape %s, expected %s" % (k, tuple(out[k].shape), tuple(shape)))
    return {k: np.asarray(out[k], np.float32) for k in spec}


class InceptionFeatureExtractor:
    """metrics/inception_distance.py:8-27.  ``weights``: an .npz / dict of the ImageNet InceptionV3 (load_inception_arrays),
    default from $CONFIGNET_INCEPTION_WEIGHTS; without one the network holds SEEDED STAND-IN weights (the pretrained file
    cannot be downloaded offline) and warns: KID / FID are then distances between random features."""

    def __init__(self, input_shape, weights=None, device=None, seed=4321):
        self.input_shape = tuple(input_shape)
        self.device = torch.device(device if device is not None else "cuda:0")
        self.output_shape = (None, 2048)
        weights = weights if weights is not None else os.environ.get(ENV_WEIGHTS)
        if weights is None:
            warnings.warn("InceptionFeatureExtractor: no InceptionV3 weights given (%s unset) - seeded stand-in weights, "
                          "KID / FID are not comparable with published numbers" % ENV_WEIGHTS, stacklevel=2)
            raw = nets.init_stand_in(nets.inception_v3_spec(), seed)
        else:
            raw = load_inception_arrays(weights)
        self.pretrained = weights is not None
        self.raw_weights = raw
        self.group = None               # folded kernels in HBM, created at the first call (the host half runs without a GPU)

    def _params(self):
        if self.group is None:
            self.group = ParamGroup(nets.fold_inception_params(self.raw_weights), self.device)
            self.group.set_frozen()
        return self.group.params

    def features_device(self, images):
        """images: uint8 or float [0,255] (n,H,W,3), NumPy or device tensor -> (n,2048) device tensor"""
        x = images if isinstance(images, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(images))
        x = x.to(self.device, non_blocking=True)
        # inception_v3.preprocess_input (mode 'tf'): astype(float32) / 127.5 - 1
        x = ops.from_uint8(x) if x.dtype == torch.uint8 else nets.pixel_map(x.float(), 1)
        return nets.inception_v3_features(self._params(), x)

    def get_features(self, images, max_chunk_size=1000):
        n_imgs = images.shape[0]
        features = np.zeros((n_imgs, 2048), np.float32)
        n_chunks = 1 + n_imgs // max_chunk_size
        for i in range(n_chunks):
            chunk_begin = i * max_chunk_size
            chunk_end = min((i + 1) * max_chunk_size, n_imgs)
            if chunk_end - chunk_begin <= 0:
                break
            parts = [self.features_device(images[b:min(b + DEVICE_CHUNK, chunk_end)])
                     for b in range(chunk_begin, chunk_end, DEVICE_CHUNK)]
            features[chunk_begin:chunk_end] = torch.cat(parts, dim=0).cpu().numpy()
        return features


def polynomial_kernel(x, y=None, degree=3, coef0=1.0):
    """sklearn.metrics.pairwise.polynomial_kernel(gamma=None): (x.y / n_features + coef0) ** degree in the input dtype"""
    y = x if y is None else y
    k = np.dot(x, y.T)
    k *= 1.0 / x.shape[1]
    k += coef0
    k **= degree
    return k


def compute_FID(features_g, features_r):
    """metrics/inception_distance.py:29-43"""
    mean_g = np.mean(features_g, axis=0)
    mean_r = np.mean(features_r, axis=0)
    cov_g = np.cov(features_g, rowvar=False)
    cov_r = np.cov(features_r, rowvar=False)
    centroid_distance = np.linalg.norm(mean_g - mean_r) ** 2
    covariance_distance = np.real(np.trace(cov_g + cov_r - 2 * scipy.linalg.sqrtm(np.dot(cov_g, cov_r))))
    return centroid_distance + covariance_distance


def compute_KID(features_g, features_r):
    """metrics/inception_distance.py:45-59 (eq. 4 of arXiv:1801.01401: the unbiased MMD^2 estimate, cubic kernel)"""
    kernel_gen_gen = polynomial_kernel(features_g)
    kernel_real_real = polynomial_kernel(features_r)
    kernel_gen_real = polynomial_kernel(features_g, features_r)
    m, n = features_g.shape[0], features_r.shape[0]
    term1 = (1 / (m * (m - 1))) * (np.sum(kernel_gen_gen) - np.sum(np.diagonal(kernel_gen_gen)))
    term2 = (1 / (n * (n - 1))) * (np.sum(kernel_real_real) - np.sum(np.diagonal(kernel_real_real)))
    term3 = (1 / (m * n)) * np.sum(kernel_gen_real)
    return term1 + term2 - 2 * term3
