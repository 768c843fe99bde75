"""Training-time metrics, GPU half: the kernels of csrc/metrics.cu, InceptionV3 and the
MobileNetV2 attribute classifier on the B200 kernels against the fp64 oracle, and the metric classes end to end."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import metrics_fakes as FK                                          # noqa: E402
from oracle import metrics_oracle as MO                             # noqa: E402
from oracle import confignet_oracle as O                            # noqa: E402

pytestmark = pytest.mark.gpu


def _dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a)).to("cuda:0", dtype)


def _rel(got, want):
    want = np.asarray(want, np.float64)
    return float(np.abs(np.asarray(got, np.float64) - want).max() / max(np.abs(want).max(), 1e-30))


@pytest.mark.parametrize("shape", [(2, 17, 17, 64), (1, 35, 29, 6), (3, 8, 8, 192), (1, 9, 11, 5)])
def test_window_pooling(shape):
    from confignet_b200 import _lib as L
    from confignet_b200.metrics import nets
    x = np.random.RandomState(0).standard_normal(shape).astype(np.float32)
    xt = torch.tensor(x, dtype=torch.float64)
    got = nets.pool2d(_dev(x), 3, 2, False, L.POOL_MAX).cpu().numpy()
    assert np.array_equal(got, MO.maxpool_valid(xt).numpy().astype(np.float32))                # a selection: exact
    got = nets.pool2d(_dev(x), 3, 1, True, L.POOL_AVG_VALID).cpu().numpy()
    assert _rel(got, MO.avgpool_same_3x3(xt).numpy()) < 1e-6                                  # borders divide by 4 / 6, not 9


@pytest.mark.parametrize("cfg", [(2, 16, 16, 32, 1), (2, 16, 16, 32, 2), (1, 15, 17, 96, 2), (1, 7, 9, 6, 1), (1, 8, 8, 5, 2)])
def test_depthwise_conv(cfg):
    from confignet_b200 import _lib as L
    from confignet_b200.metrics import nets
    n, h, w, c, stride = cfg
    rng = np.random.RandomState(1)
    x = rng.standard_normal((n, h, w, c)).astype(np.float32)
    k = rng.standard_normal((3, 3, c, 1)).astype(np.float32)
    b = rng.standard_normal(c).astype(np.float32)
    xt = torch.tensor(x, dtype=torch.float64)
    if stride == 2:          # keras-applications: ZeroPadding2D(correct_pad) + VALID
        xt = MO._pad_hw(xt, MO._correct_pad(h), MO._correct_pad(w))
    want = MO.relu6(MO.depthwise3x3(xt, torch.tensor(k, dtype=torch.float64), stride) + torch.tensor(b, dtype=torch.float64))
    got = nets.dwconv3x3(_dev(x), _dev(k.reshape(3, 3, c)), _dev(b), stride, L.ACT_RELU6).cpu().numpy()
    assert got.shape == tuple(want.shape) and _rel(got, want.numpy()) < 1e-6


def test_resize_and_pixel_maps():
    from confignet_b200.metrics import nets
    rng = np.random.RandomState(2)
    for (h, w, oh, ow) in [(256, 256, 128, 128), (256, 256, 224, 224), (100, 130, 64, 64), (64, 64, 128, 128), (31, 57, 16, 24)]:
        img = rng.randint(0, 256, (3, h, w, 3)).astype(np.uint8)
        got = nets.resize_images(_dev(img, torch.uint8), oh, ow).cpu().numpy()
        want = np.stack([MO.cv2_resize_linear(i, oh, ow) for i in img])
        assert np.array_equal(got, want), (h, w, oh, ow)                                       # fixed-point form: bit-exact
        imf = rng.uniform(0, 255, (2, h, w, 3)).astype(np.float32)
        got = nets.resize_images(_dev(imf), oh, ow).cpu().numpy()
        want = np.stack([MO.cv2_resize_linear(i, oh, ow) for i in imf])
        assert np.abs(got - want).max() < 1e-3, (h, w, oh, ow)
    x = rng.uniform(-1, 1, (5, 7, 3)).astype(np.float32)
    assert np.array_equal(nets.pixel_map(_dev(x), 0).cpu().numpy(), (x + 1) * 127.5)
    y = rng.uniform(0, 255, (5, 7, 3)).astype(np.float32)
    assert np.array_equal(nets.pixel_map(_dev(y), 1).cpu().numpy(), y / np.float32(127.5) - np.float32(1))
    u = rng.randint(0, 256, (4, 5, 3)).astype(np.uint8)
    assert np.array_equal(nets.u8_to_f32(_dev(u, torch.uint8)).cpu().numpy(), u.astype(np.float32))


def test_inception_v3_features_match_oracle():
    """InceptionV3 at the metric's input size (256 x 256, B = 2): 94 folded conv launches + pooling against the unfolded fp64
    restatement; 1e-3 relative on the 2048 features (network-output bar)."""
    from confignet_b200.metrics import nets, InceptionFeatureExtractor
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ex = InceptionFeatureExtractor((256, 256, 3), device="cuda:0", seed=3)
    imgs = np.random.RandomState(0).randint(0, 256, (2, 256, 256, 3)).astype(np.uint8)
    got = ex.get_features(imgs)
    x = torch.tensor(imgs.astype(np.float32) / np.float32(127.5) - np.float32(1), dtype=torch.float64)
    want = MO.inception_v3_features(O.to_torch(ex.raw_weights, dtype=torch.float64), x).numpy()
    assert got.shape == (2, 2048) and got.dtype == np.float32 and float(want.std()) > 1e-3
    assert _rel(got, want) < 1e-3
    # chunking leaves the features unchanged (per-sample network): one image at a time == the batch
    one = np.concatenate([ex.get_features(imgs[i:i + 1]) for i in range(2)])
    assert np.abs(one - got).max() <= 1e-5 * np.abs(got).max()
    # float images in 0..255 take the same path as uint8 ones
    assert np.array_equal(ex.get_features(imgs.astype(np.float32)), got)


@pytest.mark.parametrize("size", [128, 96])
def test_attribute_classifier_matches_oracle(size):
    """MobileNetV2 + head at the classifier's own input size, fed 256 x 256 uint8 generator outputs (resized on the device)
    and float images in [-1, 1] of the right size."""
    from confignet_b200.metrics import CelebaAttributeClassifier
    cfg = {"input_shape": [size, size, 3], "predicted_attributes": list(FK.ATTRIBUTES)}
    c = CelebaAttributeClassifier(cfg, device="cuda:0", seed=7)
    p = O.to_torch(c._raw, dtype=torch.float64)
    rng = np.random.RandomState(3)
    imgs = rng.randint(0, 256, (3, 256, 256, 3)).astype(np.uint8)
    got = c.predict_attributes(imgs)
    small = np.stack([MO.cv2_resize_linear(i, size, size) for i in imgs])
    x = torch.tensor(small.astype(np.float32) / np.float32(127.5) - np.float32(1), dtype=torch.float64)
    want = MO.attribute_classifier_forward(p, x).numpy()
    assert got.shape == (3, len(FK.ATTRIBUTES)) and float(want.std()) > 1e-3
    assert np.abs(got - want).max() < 1e-3                                                     # probabilities
    xf = rng.uniform(-1, 1, (2, size, size, 3)).astype(np.float32)
    got = c.predict_attributes(xf)
    back = ((xf + 1) * np.float32(127.5)) / np.float32(127.5) - np.float32(1)
    want = MO.attribute_classifier_forward(p, torch.tensor(back, dtype=torch.float64)).numpy()
    assert np.abs(got - want).max() < 1e-3
    assert np.abs(c.predict(back) - want).max() < 1e-3


def test_metric_classes_end_to_end(tmp_path):
    """calculate_metrics of both stages on tiny sets: InceptionMetrics / ControllabilityMetrics around the real networks.
    KID / FID equal the oracle's formulas on the oracle's InceptionV3 features of the very images the model generated."""
    import warnings
    from confignet_b200 import ConfigNetFirstStage, ConfigNet, netspec
    from confignet_b200.metrics import CelebaAttributeClassifier
    fm = {k: tuple(v) for k, v in netspec.default_facemodel_inputs().items()}
    rng = np.random.RandomState(4)

    class DS:
        pass
    ds = DS()
    ds.imgs = rng.randint(0, 256, (6, 256, 256, 3)).astype(np.uint8)
    ds.inception_features = rng.standard_normal((6, 2048)).astype(np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = ConfigNetFirstStage({"output_shape": (256, 256, 3), "batch_size": 2, "facemodel_inputs": fm}, device="cuda:0")
        from confignet_b200.metrics.metrics import InceptionMetrics
        np.random.seed(0)
        m._inception_metric_object = InceptionMetrics(m.config, ds, n_samples_for_metrics=5, device="cuda:0")
        m._generator_input_for_metrics = {"latent": m.sample_latent_vector(4), "rotation": m.sample_rotations(4)}
        m.calculate_metrics(str(tmp_path))
    assert m.metrics["training_step_number"] == [0] and len(m.metrics["kid"]) == len(m.metrics["fid"]) == 1
    imgs = m.generate_output_for_metrics()
    ex = m._inception_metric_object.inception_feature_extractor
    x = torch.tensor(imgs.astype(np.float32) / np.float32(127.5) - np.float32(1), dtype=torch.float64)
    feats = MO.inception_v3_features(O.to_torch(ex.raw_weights, dtype=torch.float64), x).numpy().astype(np.float32)
    gt = m._inception_metric_object.gt_inception_features
    kid, fid = MO.compute_KID(feats, gt), MO.compute_FID(feats, gt)
    assert abs(m.metrics["kid"][0] - kid) <= 5e-3 * abs(kid) and abs(m.metrics["fid"][0] - fid) <= 5e-3 * abs(fid)
    rows = np.loadtxt(tmp_path / "inception_metrics.txt", ndmin=2)
    assert rows.shape == (1, 3) and rows[0, 0] == 0
    # a metric-sized generate_images call walks the generator in chunks of 64: a chunk equals a call of its own size bit
    # for bit; against a call of another batch size (other reduction splits) the truncating uint8 cast may move a grey level
    lat, rot = m.sample_latent_vector(66), m.sample_rotations(66)
    big = m.generate_images(lat, rot)
    assert big.shape == (66, 256, 256, 3) and big.dtype == np.uint8
    assert np.array_equal(big[64:], m.generate_images(lat[64:], rot[64:]))
    d = np.abs(big[:3].astype(int) - m.generate_images(lat[:3], rot[:3]).astype(int))
    assert d.max() <= 1 and (d != 0).mean() < 1e-2

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m2 = ConfigNet({"output_shape": (256, 256, 3), "batch_size": 2, "facemodel_inputs": fm}, device="cuda:0")
        clf = CelebaAttributeClassifier({"input_shape": [128, 128, 3], "predicted_attributes": list(FK.ATTRIBUTES)}, device="cuda:0")
        synth = DS()
        synth.imgs = ds.imgs
        from sklearn.mixture import GaussianMixture
        m2.facemodel_param_distributions = {k: GaussianMixture(1, random_state=0).fit(rng.standard_normal((8, v[0]))) for k, v in fm.items()}
        np.random.seed(1)
        m2._inception_metric_object = InceptionMetrics(m2.config, ds, n_samples_for_metrics=5, device="cuda:0")
        m2._generator_input_for_metrics = {"input_images": ds.imgs[:3]}
        from confignet_b200.metrics.metrics import ControllabilityMetrics
        m2.controllability_metrics = ControllabilityMetrics(m2, clf)
        m2.calculate_metrics(str(tmp_path / "s2"))
    assert len(m2.metrics["kid"]) == 1 and len(m2.metrics["perceptual_loss"]) == 1 and np.isfinite(m2.metrics["perceptual_loss"][0])
    assert 0.0 <= m2.metrics["contr_attribute_means"][0][0] <= 1.0 and np.isfinite(m2.metrics["controllability"][0])
    assert sorted(os.listdir(tmp_path / "s2")) == ["controllability_metrics.json", "image_metrics.txt", "inception_metrics.txt"]
