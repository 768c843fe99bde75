"""Training-time metrics, CPU half: the oracle and the product's host logic against golden
vectors produced by EXECUTING the reference's metric code (scripts/make_golden_metrics_from_reference.py), the cv2.resize
restatement against OpenCV itself, BatchNorm folding, weight-file order and the InceptionV3 wiring (product graph against
the oracle's independent restatement, both on torch CPU)."""
import json
import os
import sys
import types

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import metrics_fakes as FK                                         # noqa: E402
from oracle import metrics_oracle as MO                            # noqa: E402
from oracle import confignet_oracle as O                           # noqa: E402
from confignet_b200.metrics import nets, inception_distance as ID  # noqa: E402
from confignet_b200.metrics import metrics as PM                   # noqa: E402
from confignet_b200.metrics import ControllabilityMetricConfigs, CelebaAttributeClassifier   # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "reference_metrics.npz"))
with open(os.path.join(HERE, "golden", "reference_metrics.json")) as _fp:
    META = json.load(_fp)


@pytest.mark.parametrize("tag", ["small", "wide", "f64"])
def test_fid_kid_match_the_executed_reference(tag):
    g, r = GOLD["feat_g_" + tag], GOLD["feat_r_" + tag]
    for mod in (MO, ID):
        assert abs(mod.compute_FID(g, r) - META["fid_" + tag]) <= 1e-9 * max(1.0, abs(META["fid_" + tag])), mod.__name__
        # float32 kernels: sklearn's safe_sparse_dot and np.dot call the same BLAS; allow the last bits of a different blocking
        assert abs(float(mod.compute_KID(g, r)) - META["kid_" + tag]) <= 2e-6 * max(1.0, abs(META["kid_" + tag])), mod.__name__
    assert ID.compute_KID(g, r).dtype == g.dtype or isinstance(ID.compute_KID(g, r), float)


class _Classifier:
    def __init__(self):
        self.config = {"predicted_attributes": list(FK.ATTRIBUTES), "input_shape": (8, 8, 3)}

    def predict_attributes(self, images):
        return FK.fake_predict_attributes(images)


@pytest.mark.parametrize("iters", [0, 2])
def test_controllability_metrics_host_logic_matches_the_executed_reference(tmp_path, iters):
    """metrics/metrics.py:15-199 driven with the shared fakes: the same face-model parameters reach the synthetic encoder,
    the same modified latents reach generate_images in the same order, and the same numbers land in
    controllability_metrics.json - for the encoder path and the per-image fine-tuning path."""
    assert [n for n, _ in ControllabilityMetricConfigs.all_configs()] == META["config_names"]
    model = FK.FakeModel()
    cm = PM.ControllabilityMetrics(model, _Classifier(), per_image_tuning_iters=iters)
    md = {"training_step_number": [0]}
    cm.update_and_log_metrics(GOLD["contr_input_images"], md, str(tmp_path))
    assert list(md.keys()) == META["contr_keys_iters%d" % iters]
    assert json.loads(json.dumps(model.log)) == META["contr_log_iters%d" % iters]
    with open(tmp_path / "controllability_metrics.json") as fp:
        got = json.load(fp)
    want = META["contr_json_iters%d" % iters]
    assert list(got.keys()) == list(want.keys())
    for k in want:
        assert np.allclose(np.asarray(got[k], np.float64), np.asarray(want[k], np.float64), rtol=1e-12, atol=1e-15), k
    if iters == 0:        # get_metrics(img_output_dir=...): the reference's per-image PNG dumps (metrics.py:141-154)
        cv2 = pytest.importorskip("cv2")
        cm.get_metrics(GOLD["contr_input_images"], img_output_dir=str(tmp_path / "imgs"))
        files = sorted(os.listdir(tmp_path / "imgs"))
        assert len(files) == 5 * (2 + 2 * 8) and "gt_img_0000.png" in files and "smile_config_img_not_set_0004.png" in files
        assert np.array_equal(cv2.imread(str(tmp_path / "imgs" / "gt_img_0003.png")), GOLD["contr_input_images"][3])
    pair = cm.get_metrics_for_attribute_pairs(FK.fake_predict_attributes(GOLD["contr_input_images"]),
                                              FK.fake_predict_attributes(GOLD["contr_input_images"][::-1]),
                                              ControllabilityMetricConfigs.smile_config)
    names = FK.ATTRIBUTES
    const = [i for i, n in enumerate(names) if n not in ("Narrow_Eyes", "Mouth_Slightly_Open", "Smiling")]
    assert pair == MO.attribute_pair_metrics(FK.fake_predict_attributes(GOLD["contr_input_images"]),
                                             FK.fake_predict_attributes(GOLD["contr_input_images"][::-1]), names.index("Smiling"), const)


def test_inception_metrics_host_logic_matches_the_executed_reference(tmp_path):
    """metrics/metrics.py:201-264: the sample draw (stream position), kid / fid histories over two updates, the text table"""
    ds = FK.FakeDataset()
    np.random.seed(11)
    # the constructor proper (no GPU needed: the device network is built at the first get_features call)
    with pytest.warns(UserWarning, match="stand-in"):
        im = PM.InceptionMetrics({"output_shape": (8, 8, 3)}, ds, n_samples_for_metrics=20, device="cpu")
    assert int(np.random.randint(0, 2 ** 31 - 1)) == META["inception_next_draw"]
    im.inception_feature_extractor = types.SimpleNamespace(get_features=FK.fake_inception_features)
    md = {"training_step_number": [0]}
    gen = GOLD["inception_generated"]
    im.update_and_log_metrics(gen, md, str(tmp_path))
    md["training_step_number"].append(1000)
    im.update_and_log_metrics(gen[::-1] // 2, md, str(tmp_path))
    want = META["inception_metrics_dict"]
    assert list(md.keys()) == list(want.keys())
    for k in want:
        assert np.allclose(np.asarray(md[k], np.float64), want[k], rtol=2e-6, atol=1e-9), k
    got_rows = np.loadtxt(tmp_path / "inception_metrics.txt", ndmin=2)
    want_rows = np.loadtxt(META["inception_metrics_txt"].splitlines(), ndmin=2)
    assert open(tmp_path / "inception_metrics.txt").readline() == META["inception_metrics_txt"].splitlines(True)[0]
    assert got_rows.shape == want_rows.shape == (2, 3) and np.allclose(got_rows, want_rows, rtol=2e-6)


def test_cv2_resize_restatement_against_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(0)
    for (h, w, oh, ow) in [(256, 256, 128, 128), (256, 256, 224, 224), (100, 130, 64, 64), (256, 256, 96, 160), (31, 57, 16, 24)]:
        img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        assert np.array_equal(cv2.resize(img, (ow, oh)), MO.cv2_resize_linear(img, oh, ow)), (h, w, oh, ow)     # shrinking: bit-exact
        imf = rng.uniform(0, 255, (h, w, 3)).astype(np.float32)
        assert np.abs(cv2.resize(imf, (ow, oh)) - MO.cv2_resize_linear(imf, oh, ow)).max() < 2e-2           # 1e-4 of the 0..255 range
    img = rng.randint(0, 256, (64, 64, 3)).astype(np.uint8)                                                   # enlarging: +-1 grey level
    d = np.abs(cv2.resize(img, (128, 128)).astype(int) - MO.cv2_resize_linear(img, 128, 128).astype(int))
    assert d.max() <= 1 and (d != 0).mean() < 0.01


def test_specs_folding_and_weight_orders():
    spec = nets.inception_v3_spec()
    assert sum(int(np.prod(s)) for s, _ in spec.values()) == 21802784                 # keras: InceptionV3 without top
    cspec = nets.attribute_classifier_spec(40)
    base = sum(int(np.prod(s)) for k, (s, _) in cspec.items() if not k.startswith(("batch_normalization/", "dense/")))
    assert base == 2257984                                                            # keras: MobileNetV2 (alpha 1) without top
    order = nets.attribute_classifier_keras_order(40)
    assert sorted(order) == sorted(cspec.keys()) and order[0] == "Conv1/kernel" and order[1] == "bn_Conv1/gamma"
    first_moving = next(i for i, k in enumerate(order) if k.endswith("/moving_mean"))
    assert all(not k.endswith(("/kernel", "/gamma", "/beta", "/depthwise_kernel")) for k in order[first_moving:-6])
    assert order[-6:] == ["batch_normalization/gamma", "batch_normalization/beta", "batch_normalization/moving_mean",
                          "batch_normalization/moving_variance", "dense/kernel", "dense/bias"]
    # folding: conv -> BatchNorm == conv with the folded kernel + bias
    rng = np.random.RandomState(1)
    k, gamma, beta = rng.standard_normal((3, 3, 4, 6)), rng.uniform(0.5, 1.5, 6), rng.standard_normal(6)
    mean, var = rng.standard_normal(6), rng.uniform(0.5, 2.0, 6)
    kf, bf = nets.fold_batchnorm(k, gamma, beta, mean, var, 1e-3)
    x = torch.tensor(rng.standard_normal((2, 7, 7, 4)))
    p = {"bn/gamma": torch.tensor(gamma), "bn/beta": torch.tensor(beta), "bn/moving_mean": torch.tensor(mean), "bn/moving_variance": torch.tensor(var)}
    want = MO.batchnorm(MO.conv2d(x, torch.tensor(k)), p, "bn", 1e-3)
    got = O.conv_same(x, torch.tensor(kf, dtype=torch.float64), torch.tensor(bf, dtype=torch.float64))
    assert (want - got).abs().max() < 1e-5
    kd = rng.standard_normal((3, 3, 6, 1))
    kdf, bdf = nets.fold_batchnorm(kd, gamma, beta, mean, var, 1e-3, depthwise=True)
    x = torch.tensor(rng.standard_normal((2, 7, 7, 6)))
    want = MO.batchnorm(MO.depthwise3x3(x, torch.tensor(kd), 1), p, "bn", 1e-3)
    got = MO.depthwise3x3(x, torch.tensor(kdf, dtype=torch.float64), 1) + torch.tensor(bdf, dtype=torch.float64)
    assert (want - got).abs().max() < 1e-5


class _CpuRunner:
    """the product's InceptionV3 graph walked with torch-CPU layers on the FOLDED parameters"""

    def __init__(self, p, x):
        self.p, self.input, self.n = p, x, 0

    def conv(self, x, filters, kh, kw, stride=1, valid=False):
        self.n += 1
        k, b = self.p["conv2d_%d/kernel" % self.n], self.p["conv2d_%d/bias" % self.n]
        return torch.relu(MO.conv2d(x, k, stride, valid) + b)

    def maxpool(self, x): return MO.maxpool_valid(x)
    def avgpool(self, x): return MO.avgpool_same_3x3(x)
    def concat(self, xs): return torch.cat(xs, dim=-1)
    def gap(self, x): return x.mean(dim=(1, 2))


def test_inception_graph_and_folding_equal_the_oracle_restatement():
    raw = nets.init_stand_in(nets.inception_v3_spec(), 3)
    x = torch.tensor(np.random.RandomState(0).uniform(-1, 1, (1, 139, 139, 3)))
    want = MO.inception_v3_features(O.to_torch(raw, dtype=torch.float64), x)
    got = nets.inception_v3_graph(_CpuRunner(O.to_torch(nets.fold_inception_params(raw), dtype=torch.float64), x))
    assert want.shape == (1, 2048) and float(want.std()) > 1e-3
    assert float((want - got).abs().max() / want.abs().max()) < 1e-5


def test_inception_weight_file_renumbering():
    raw = nets.init_stand_in(nets.inception_v3_spec(), 5)
    shifted = {}
    for k, v in raw.items():                    # a Keras process that had already built 188 conv / 203 BatchNorm layers
        layer, var = k.split("/")
        base, num = layer.rsplit("_", 1)
        shifted["%s_%d/%s" % (base, int(num) + (188 if base == "conv2d" else 203), var)] = v
    back = ID.load_inception_arrays(shifted)
    assert list(back.keys()) == list(raw.keys()) and all(np.array_equal(back[k], raw[k]) for k in raw)
    with pytest.raises(ValueError):
        ID.load_inception_arrays({k: v for k, v in shifted.items() if not k.startswith("conv2d_200/")})


def test_attribute_classifier_files_round_trip_without_a_gpu(tmp_path):
    """celeba_attribute_prediction.py:31-52: <name>.json (config + logs) and <name>.npy (object array, get_weights() order)"""
    cfg = {"input_shape": [128, 128, 3], "predicted_attributes": list(FK.ATTRIBUTES), "optimizer": {"lr": 0.001}, "batch_size": 32}
    c = CelebaAttributeClassifier(cfg, device="cpu")
    c.logs = {"loss": [0.5]}
    c.save(str(tmp_path), "model")
    w = np.load(tmp_path / "model.npy", allow_pickle=True)
    order = nets.attribute_classifier_keras_order(len(FK.ATTRIBUTES))
    assert w.dtype == object and len(w) == len(order) == 266 and w[0].shape == (3, 3, 3, 32) and w[-2].shape == (1280, len(FK.ATTRIBUTES))
    c2 = CelebaAttributeClassifier.load(str(tmp_path / "model.json"), device="cpu")
    assert c2.logs == c.logs and c2.config == cfg
    assert all(np.array_equal(a, b) for a, b in zip(c.get_weights(), c2.get_weights()))
    with pytest.raises(ValueError):
        c2.classifier.set_weights(list(w)[:-1])
    from confignet_b200._lib import CnError
    with pytest.raises((CnError, RuntimeError, AssertionError)):       # no CPU fallback: predicting without CUDA fails loudly
        c2.predict_attributes(np.zeros((1, 128, 128, 3), np.uint8))


# ------------------------------------------------------------------------------------------------ third-party cross-check
def _kernel_to_torch(k):
    return torch.tensor(np.transpose(k, (3, 2, 0, 1)).copy())


def test_oracle_inception_v3_equals_torchvision_with_tf_average_pooling():
    """keras-applications cannot be executed here, torchvision can: its Inception3 is the same published architecture
    (Szegedy et al. 2015: stem, 3 x InceptionA, InceptionB, 4 x InceptionC, InceptionD, 2 x InceptionE; BatchNorm eps 1e-3)
    and differs from the Keras model in ONE layer semantic - its 3x3 average pools count the zero padding, TensorFlow's
    SAME pooling does not.  With that one function patched and the same weights (BasicConv2d modules in registration order
    = Keras creation order; BatchNorm weight 1 = scale=False), the oracle's restatement must reproduce torchvision's
    network: an independent pin of the wiring (branch order, kernel shapes, strides, paddings, concatenation order)."""
    tv = pytest.importorskip("torchvision")
    from torchvision.models import inception as tvi
    raw = nets.init_stand_in(nets.inception_v3_spec(), 11)
    model = tvi.Inception3(num_classes=3, aux_logits=False, transform_input=False, init_weights=False).double().eval()
    blocks = [m for m in model.modules() if isinstance(m, tvi.BasicConv2d)]
    assert len(blocks) == 94
    with torch.no_grad():
        for i, b in enumerate(blocks, start=1):
            k = raw["conv2d_%d/kernel" % i]
            assert tuple(b.conv.weight.shape) == (k.shape[3], k.shape[2], k.shape[0], k.shape[1]), (i, b.conv.weight.shape, k.shape)
            b.conv.weight.copy_(_kernel_to_torch(k))
            q = "batch_normalization_%d" % i
            assert b.bn.eps == 1e-3
            b.bn.weight.fill_(1.0)
            b.bn.bias.copy_(torch.tensor(raw[q + "/beta"]))
            b.bn.running_mean.copy_(torch.tensor(raw[q + "/moving_mean"]))
            b.bn.running_var.copy_(torch.tensor(raw[q + "/moving_variance"]))
    x = torch.tensor(np.random.RandomState(1).uniform(-1, 1, (1, 151, 139, 3)))
    keep = tvi.F.avg_pool2d
    tvi.F.avg_pool2d = lambda t, kernel_size, stride=None, padding=0: keep(t, kernel_size, stride, padding, count_include_pad=False)
    try:
        with torch.no_grad():
            t = x.permute(0, 3, 1, 2)
            for name, layer in model.named_children():          # Inception3._forward without dropout / fc
                if name in ("AuxLogits", "dropout", "fc"):
                    continue
                t = layer(t)
            want = torch.flatten(t, 1)
    finally:
        tvi.F.avg_pool2d = keep
    got = MO.inception_v3_features(O.to_torch(raw, dtype=torch.float64), x)
    assert want.shape == got.shape == (1, 2048) and float(want.std()) > 1e-3
    assert float((want - got).abs().max() / want.abs().max()) < 1e-9


def test_oracle_mobilenet_v2_equals_torchvision_on_odd_sizes():
    """torchvision's MobileNetV2 is the architecture of the paper's table 2, as keras-applications' is; the two differ in
    the padding of the stride-2 layers (torch pads 1 on both sides, Keras pads by correct_pad() = TF SAME) and in the
    BatchNorm epsilon.  On an input whose size stays odd down the network (97 -> 49 -> 25 -> 13 -> 7) correct_pad() is (1, 1)
    too, so with eps set to 1e-3 and the same weights the oracle's restatement must reproduce torchvision's features:
    an independent pin of the block table, the expansion / depthwise / projection order and the residual rule."""
    pytest.importorskip("torchvision")
    from torchvision.models import mobilenetv2 as tvm
    n_attr = 5
    raw = nets.init_stand_in(nets.attribute_classifier_spec(n_attr), 12)
    model = tvm.MobileNetV2(num_classes=3).double().eval()
    convs = [m for m in model.features.modules() if isinstance(m, torch.nn.Conv2d)]
    bns = [m for m in model.features.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    layers = nets.mobilenet_v2_layers()
    assert len(convs) == len(bns) == len(layers) == 52
    with torch.no_grad():
        for (kind, cname, bname, cin, cout, stride, act, add), conv, bn in zip(layers, convs, bns):
            if kind == "dw":
                k = raw[cname + "/depthwise_kernel"]                      # (3,3,C,1) -> (C,1,3,3)
                w = torch.tensor(np.transpose(k, (2, 3, 0, 1)).copy())
                assert conv.groups == cin and conv.stride == (stride, stride)
            else:
                w = _kernel_to_torch(raw[cname + "/kernel"])
                assert conv.groups == 1 and conv.stride == (stride, stride)
            assert tuple(conv.weight.shape) == tuple(w.shape), (cname, conv.weight.shape, w.shape)
            conv.weight.copy_(w)
            bn.eps = 1e-3
            bn.weight.copy_(torch.tensor(raw[bname + "/gamma"])); bn.bias.copy_(torch.tensor(raw[bname + "/beta"]))
            bn.running_mean.copy_(torch.tensor(raw[bname + "/moving_mean"])); bn.running_var.copy_(torch.tensor(raw[bname + "/moving_variance"]))
    x = torch.tensor(np.random.RandomState(2).uniform(-1, 1, (2, 97, 97, 3)))
    with torch.no_grad():
        want = model.features(x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    got = MO.mobilenet_v2_features(O.to_torch(raw, dtype=torch.float64), x)
    assert tuple(got.shape) == tuple(want.shape) == (2, 4, 4, 1280) and float(want.std()) > 1e-3
    assert float((want - got).abs().max() / want.abs().max()) < 1e-9


class _ClassifierCpuLayers:
    """torch-CPU layers behind the product's attribute_classifier_graph, on the FOLDED parameters"""

    @staticmethod
    def conv(x, k, b, stride, relu6):
        y = O.conv_same(x, k, b, stride)
        return MO.relu6(y) if relu6 else y

    @staticmethod
    def dwconv(x, k, b, stride):
        xt = x
        if stride == 2:
            xt = MO._pad_hw(x, MO._correct_pad(x.shape[1]), MO._correct_pad(x.shape[2]))
        return MO.relu6(MO.depthwise3x3(xt, k.unsqueeze(-1), stride) + b)

    add = staticmethod(lambda a, b: a + b)
    gap = staticmethod(lambda x: x.mean(dim=(1, 2)))
    dense_sigmoid = staticmethod(lambda f, k, b: torch.sigmoid(f @ k + b))


@pytest.mark.parametrize("size", [96, 97])
def test_classifier_graph_and_folding_equal_the_oracle_restatement(size):
    """the product's MobileNetV2 + head wiring with BatchNorm folded into kernels / biases / the Dense layer, against the
    oracle's unfolded restatement (even size: TF-SAME (0,1) padding of the stride-2 layers; odd size: (1,1))"""
    raw = nets.init_stand_in(nets.attribute_classifier_spec(7), 21)
    x = torch.tensor(np.random.RandomState(3).uniform(-1, 1, (2, size, size, 3)))
    want = MO.attribute_classifier_forward(O.to_torch(raw, dtype=torch.float64), x)
    folded = O.to_torch(nets.fold_attribute_classifier_params(raw, 7), dtype=torch.float64)
    got = nets.attribute_classifier_graph(_ClassifierCpuLayers, folded, x)
    assert tuple(got.shape) == (2, 7) and float(want.std()) > 1e-4
    assert float((want - got).abs().max()) < 1e-6
