"""The two frozen third-party networks behind the training-time metrics, on this package's kernels:

  InceptionV3   keras.applications.inception_v3.InceptionV3(include_top=False, weights="imagenet", pooling="avg")
                (metrics/inception_distance.py:9-11) -> 2048 features per image for KID / FID
  MobileNetV2   keras.applications.MobileNetV2(include_top=False) + GlobalAveragePooling2D + BatchNormalization +
                Dropout + Dense(sigmoid) (metrics/celeba_attribute_prediction.py:54-62) -> attribute probabilities

keras-applications is library code outside /root/reference: the two architectures are restated from its published model
definitions ([KA-1.0.8] inception_v3.py, mobilenet_v2.py), not executed - the same standing as VGG19 / ResNet50 (DESIGN.md
section 4).  Both run forward only.  Every convolution is one launch of the implicit-GEMM family (tcgen05 where the shape
allows) with the inference BatchNorm folded into kernel and bias on the host (``fold_batchnorm``) and the ReLU / ReLU6 in the
epilogue; pooling, the depthwise convolutions and the resize are ``csrc/metrics.cu``.
"""
from collections import OrderedDict
import numpy as np
import torch

from .. import _lib as L
from .. import ops

INCEPTION_BN_EPS = 1e-3      # keras BatchNormalization default epsilon (conv2d_bn passes none) [KA-1.0.8 inception_v3.py]
MOBILENET_BN_EPS = 1e-3      # epsilon=1e-3, momentum=0.999 on every BatchNormalization [KA-1.0.8 mobilenet_v2.py]
HEAD_BN_EPS = 1e-3           # keras.layers.BatchNormalization() of the classifier head (celeba_attribute_prediction.py:60)


# ------------------------------------------------------------------------------------------------ InceptionV3
def inception_v3_graph(b):
    """The layer graph of InceptionV3 (include_top=False, pooling='avg') written against a small builder interface
    (conv / maxpool / avgpool / concat / gap), so that the parameter table and the device forward walk the same code."""
    c = b.conv
    x = c(b.input, 32, 3, 3, stride=2, valid=True)
    x = c(x, 32, 3, 3, valid=True)
    x = c(x, 64, 3, 3)
    x = b.maxpool(x)
    x = c(x, 80, 1, 1, valid=True)
    x = c(x, 192, 3, 3, valid=True)
    x = b.maxpool(x)
    for pool_filters in (32, 64, 64):                                  # mixed0..2: 35 x 35 (at 299 x 299 inputs)
        b1 = c(x, 64, 1, 1)
        b5 = c(c(x, 48, 1, 1), 64, 5, 5)
        b3 = c(c(c(x, 64, 1, 1), 96, 3, 3), 96, 3, 3)
        bp = c(b.avgpool(x), pool_filters, 1, 1)
        x = b.concat([b1, b5, b3, bp])
    b3 = c(x, 384, 3, 3, stride=2, valid=True)                         # mixed3
    bd = c(c(c(x, 64, 1, 1), 96, 3, 3), 96, 3, 3, stride=2, valid=True)
    x = b.concat([b3, bd, b.maxpool(x)])
    for f in (128, 160, 160, 192):                                     # mixed4..7: 17 x 17
        b1 = c(x, 192, 1, 1)
        b7 = c(c(c(x, f, 1, 1), f, 1, 7), 192, 7, 1)
        bd = c(c(c(c(c(x, f, 1, 1), f, 7, 1), f, 1, 7), f, 7, 1), 192, 1, 7)
        bp = c(b.avgpool(x), 192, 1, 1)
        x = b.concat([b1, b7, bd, bp])
    b3 = c(c(x, 192, 1, 1), 320, 3, 3, stride=2, valid=True)           # mixed8
    b7 = c(c(c(c(x, 192, 1, 1), 192, 1, 7), 192, 7, 1), 192, 3, 3, stride=2, valid=True)
    x = b.concat([b3, b7, b.maxpool(x)])
    for _ in range(2):                                                 # mixed9, mixed10: 8 x 8
        b1 = c(x, 320, 1, 1)
        t = c(x, 384, 1, 1)
        b3 = b.concat([c(t, 384, 1, 3), c(t, 384, 3, 1)])
        t = c(c(x, 448, 1, 1), 384, 3, 3)
        bd = b.concat([c(t, 384, 1, 3), c(t, 384, 3, 1)])
        bp = c(b.avgpool(x), 192, 1, 1)
        x = b.concat([b1, b3, bd, bp])
    return b.gap(x)


class _SpecBuilder:
    """walks a graph on channel counts only and records the Keras variables in creation order"""

    def __init__(self, cin=3):
        self.input, self.spec, self.n = cin, OrderedDict(), 0

    def conv(self, x, filters, kh, kw, stride=1, valid=False):
        self.n += 1
        self.spec["conv2d_%d/kernel" % self.n] = ((kh, kw, x, filters), "glorot")
        q = "batch_normalization_%d" % self.n                          # scale=False: no gamma
        self.spec[q + "/beta"] = ((filters,), "zeros")
        self.spec[q + "/moving_mean"] = ((filters,), "zeros")
        self.spec[q + "/moving_variance"] = ((filters,), "ones")
        return filters

    def maxpool(self, x): return x
    def avgpool(self, x): return x
    def concat(self, xs): return sum(xs)
    def gap(self, x): return x


def inception_v3_spec():
    b = _SpecBuilder()
    assert inception_v3_graph(b) == 2048 and b.n == 94
    return b.spec


def fold_batchnorm(kernel, gamma, beta, mean, var, eps, depthwise=False):
    """conv (no bias) -> inference BatchNorm == conv with kernel * s and bias beta - mean * s, s = gamma / sqrt(var + eps)
    per output channel (float64 on the host, once per weight load).  A depthwise kernel (3,3,C,1) scales along C."""
    s = (1.0 if gamma is None else np.asarray(gamma, np.float64)) / np.sqrt(np.asarray(var, np.float64) + eps)
    k = np.asarray(kernel, np.float64) * (s[:, None] if depthwise else s)
    return k.astype(np.float32), (np.asarray(beta, np.float64) - np.asarray(mean, np.float64) * s).astype(np.float32)


def fold_inception_params(raw):
    """raw Keras variables (inception_v3_spec names) -> the arrays the device holds: conv2d_i/kernel, conv2d_i/bias"""
    out = OrderedDict()
    for i in range(1, 95):
        q = "batch_normalization_%d" % i
        k, bias = fold_batchnorm(raw["conv2d_%d/kernel" % i], None, raw[q + "/beta"], raw[q + "/moving_mean"],
                                 raw[q + "/moving_variance"], INCEPTION_BN_EPS)
        out["conv2d_%d/kernel" % i], out["conv2d_%d/bias" % i] = k, bias
    return out


def pool2d(x, k, stride, same, mode):
    """MaxPooling2D((k,k), strides) VALID / AveragePooling2D((k,k), strides, padding='same') on an NHWC device tensor"""
    x = ops._chk(x)
    n, h, w, c = x.shape
    if same:
        oh, ow = -(-h // stride), -(-w // stride)
        pt = max((oh - 1) * stride + k - h, 0) // 2
        pl = max((ow - 1) * stride + k - w, 0) // 2
    else:
        oh, ow, pt, pl = (h - k) // stride + 1, (w - k) // stride + 1, 0, 0
    y = torch.empty((n, oh, ow, c), device=x.device, dtype=torch.float32)
    L.call("cn_pool2d_fwd", ops._p(x), n, h, w, c, k, k, stride, pt, pl, oh, ow, mode, ops._p(y), c, ops._stream())
    return y


class _DeviceRunner:
    """walks a graph on device tensors; ``p`` holds the folded kernels / biases"""

    def __init__(self, p, x, act=L.ACT_RELU):
        self.p, self.input, self.n, self.act = p, x, 0, act

    def conv(self, x, filters, kh, kw, stride=1, valid=False):
        self.n += 1
        q = "conv2d_%d" % self.n
        return ops.conv_act(x, self.p[q + "/kernel"], self.p[q + "/bias"], stride=stride, act=self.act, pad=0 if valid else -1)

    def maxpool(self, x): return pool2d(x, 3, 2, False, L.POOL_MAX)
    def avgpool(self, x): return pool2d(x, 3, 1, True, L.POOL_AVG_VALID)
    def concat(self, xs): return torch.cat(xs, dim=-1)          # plumbing: a strided device copy, no arithmetic
    def gap(self, x): return ops.global_avg_pool(x)


def inception_v3_features(p, x):
    """x: (B,H,W,3) float32 in [-1,1] ('tf'-mode preprocess_input already applied) -> (B,2048)"""
    with torch.no_grad():
        return inception_v3_graph(_DeviceRunner(p, x))


# ------------------------------------------------------------------------------------------------ MobileNetV2 (alpha = 1)
# (filters, stride, expansion) of the 17 inverted residual blocks, block_id = position [KA-1.0.8 mobilenet_v2.py]
MOBILENET_V2_BLOCKS = [(16, 1, 1), (24, 2, 6), (24, 1, 6), (32, 2, 6), (32, 1, 6), (32, 1, 6), (64, 2, 6), (64, 1, 6), (64, 1, 6),
                       (64, 1, 6), (96, 1, 6), (96, 1, 6), (96, 1, 6), (160, 2, 6), (160, 1, 6), (160, 1, 6), (320, 1, 6)]


def _bn4(spec, name, c):
    spec[name + "/gamma"] = ((c,), "ones")
    spec[name + "/beta"] = ((c,), "zeros")
    spec[name + "/moving_mean"] = ((c,), "zeros")
    spec[name + "/moving_variance"] = ((c,), "ones")


def mobilenet_v2_layers():
    """[(kind, conv name, bn name, cin, cout, stride, act, block input for the residual add or None)] in layer order.
    Stride-2 layers: keras-applications pads with correct_pad() = ((k//2 - (1 - size % 2), k//2), ...) and convolves VALID;
    for k = 3, s = 2 that is exactly TF SAME ((0,1) on even sizes, (1,1) on odd ones), which is what the kernels compute."""
    layers = [("conv3", "Conv1", "bn_Conv1", 3, 32, 2, "relu6", False)]
    cin = 32
    for bid, (f, s, e) in enumerate(MOBILENET_V2_BLOCKS):
        pre = "block_%d_" % bid if bid else "expanded_conv_"
        mid = cin * e
        if bid:
            layers.append(("conv1", pre + "expand", pre + "expand_BN", cin, mid, 1, "relu6", False))
        layers.append(("dw", pre + "depthwise", pre + "depthwise_BN", mid, mid, s, "relu6", False))
        layers.append(("conv1", pre + "project", pre + "project_BN", mid, f, 1, None, cin == f and s == 1))
        cin = f
    layers.append(("conv1", "Conv_1", "Conv_1_bn", cin, 1280, 1, "relu6", False))
    return layers


def attribute_classifier_spec(n_attributes):
    """MobileNetV2 base + head BatchNormalization(1280) + Dense(n_attributes), per-layer order"""
    s = OrderedDict()
    for kind, cname, bname, cin, cout, stride, act, add in mobilenet_v2_layers():
        if kind == "dw":
            s[cname + "/depthwise_kernel"] = ((3, 3, cin, 1), "glorot")
        else:
            k = 3 if kind == "conv3" else 1
            s[cname + "/kernel"] = ((k, k, cin, cout), "glorot")
        _bn4(s, bname, cout)
    _bn4(s, "batch_normalization", 1280)
    s["dense/kernel"] = ((1280, n_attributes), "glorot")
    s["dense/bias"] = ((n_attributes,), "zeros")
    return s


def attribute_classifier_keras_order(n_attributes):
    """Names in the order ``CelebaAttributeClassifier.classifier.get_weights()`` lists them (the order of the released
    .npy, celeba_attribute_prediction.py:31-37,48-50) under the reference's pinned TensorFlow 2.1: Sequential.get_weights
    concatenates ``layer.weights`` over [base_model, pooling, BatchNormalization, Dropout, Dense], and the NESTED model's
    ``weights`` is trainable_weights + non_trainable_weights (all kernels / gammas / betas in layer order, then every
    moving mean / variance) - the rule DESIGN.md section 4 pins for the nested ResNet50."""
    names = list(attribute_classifier_spec(n_attributes).keys())
    head = [k for k in names if k.startswith("batch_normalization/") or k.startswith("dense/")]
    base = [k for k in names if k not in head]
    moving = lambda k: k.endswith("/moving_mean") or k.endswith("/moving_variance")
    return [k for k in base if not moving(k)] + [k for k in base if moving(k)] + head


def fold_attribute_classifier_params(raw, n_attributes):
    """raw Keras variables -> device arrays: <conv>/kernel + <conv>/bias with the BatchNorm folded in; the head's
    BatchNorm (on the pooled features, ahead of the Dense layer) folds into the Dense kernel and bias."""
    out = OrderedDict()
    for kind, cname, bname, cin, cout, stride, act, add in mobilenet_v2_layers():
        kname = cname + ("/depthwise_kernel" if kind == "dw" else "/kernel")
        k, bias = fold_batchnorm(raw[kname], raw[bname + "/gamma"], raw[bname + "/beta"], raw[bname + "/moving_mean"],
                                 raw[bname + "/moving_variance"], MOBILENET_BN_EPS, depthwise=kind == "dw")
        out[cname + "/kernel"] = k.reshape(3, 3, cin) if kind == "dw" else k
        out[cname + "/bias"] = bias
    q = "batch_normalization"
    s = np.asarray(raw[q + "/gamma"], np.float64) / np.sqrt(np.asarray(raw[q + "/moving_variance"], np.float64) + HEAD_BN_EPS)
    t = np.asarray(raw[q + "/beta"], np.float64) - np.asarray(raw[q + "/moving_mean"], np.float64) * s
    w = np.asarray(raw["dense/kernel"], np.float64)
    out["dense/kernel"] = (w * s[:, None]).astype(np.float32)
    out["dense/bias"] = (np.asarray(raw["dense/bias"], np.float64) + t @ w).astype(np.float32)
    return out


def dwconv3x3(x, wk, bias, stride, act):
    x = ops._chk(x)
    n, h, w, c = x.shape
    y = torch.empty((n, -(-h // stride), -(-w // stride), c), device=x.device, dtype=torch.float32)
    L.call("cn_dwconv3x3_fwd", ops._p(x), ops._p(wk), ops._p(bias), n, h, w, c, stride, act, 0.0, ops._p(y), ops._stream())
    return y


def residual_add(a, b):
    """keras.layers.Add of a MobileNetV2 block"""
    out = torch.empty_like(a)
    L.call("cn_axpby", ops._p(ops._chk(a)), ops._p(ops._chk(b)), 1.0, 1.0, ops._p(out), a.numel(), ops._stream())
    return out


def act_ext_(x, act):
    """in place: ReLU6's clamp behind a conv that ran with ReLU in its epilogue / the sigmoid of the head.  The conv
    epilogues carry four activation codes only (csrc/common.cuh: a fifth one in the tcgen05 kernel's unrolled epilogue
    costs the training step 13 %)."""
    L.call("cn_act_ext", ops._p(x), ops._p(x), x.numel(), act, ops._stream())
    return x


def attribute_classifier_graph(r, p, x):
    """MobileNetV2 + head written against a small layer interface (conv / dwconv / add / gap / dense_sigmoid), so that the
    device forward and a torch-CPU check of the wiring + folding (tests/test_metrics_cpu.py) walk the same code"""
    block_in = x
    for kind, cname, bname, cin, cout, stride, act, add in mobilenet_v2_layers():
        if cname.endswith("_expand") or cname == "expanded_conv_depthwise":
            block_in = x                                       # first layer of an inverted residual block
        if kind == "dw":
            x = r.dwconv(x, p[cname + "/kernel"], p[cname + "/bias"], stride)            # + ReLU6
        else:
            x = r.conv(x, p[cname + "/kernel"], p[cname + "/bias"], stride, relu6=act == "relu6")
            if add:
                x = r.add(x, block_in)
    return r.dense_sigmoid(r.gap(x), p["dense/kernel"], p["dense/bias"])


class _ClassifierDeviceLayers:
    """the B200 kernels behind attribute_classifier_graph"""

    @staticmethod
    def conv(x, k, b, stride, relu6):
        # ReLU6 = the conv's fused ReLU + an in-place clamp: the conv epilogues carry four activation codes only
        y = ops.conv_act(x, k, b, stride=stride, act=L.ACT_RELU if relu6 else L.ACT_NONE)
        return act_ext_(y, L.ACT_RELU6) if relu6 else y

    @staticmethod
    def dwconv(x, k, b, stride):
        return dwconv3x3(x, k, b, stride, L.ACT_RELU6)

    add = staticmethod(residual_add)
    gap = staticmethod(ops.global_avg_pool)

    @staticmethod
    def dense_sigmoid(feat, k, b):
        return act_ext_(ops.conv_act(feat, k, b), L.ACT_SIGMOID)


def attribute_classifier_forward(p, x):
    """x: (B,H,W,3) float32 in [-1,1] -> (B, n_attributes) sigmoid probabilities (Dropout is the identity at inference)"""
    with torch.no_grad():
        return attribute_classifier_graph(_ClassifierDeviceLayers, p, x)


# ------------------------------------------------------------------------------------------------ image plumbing
def resize_images(x, oh, ow):
    """cv2.resize(img, (ow, oh)) per image (default INTER_LINEAR) on a device batch, uint8 or float32"""
    x = x.contiguous()
    n, h, w, c = x.shape
    y = torch.empty((n, oh, ow, c), device=x.device, dtype=x.dtype)
    is_u8 = {torch.uint8: 1, torch.float32: 0}[x.dtype]
    L.call("cn_resize_bilinear", x.data_ptr(), n, h, w, c, oh, ow, is_u8, y.data_ptr(), ops._stream())
    return y


def pixel_map(x, mode):
    x = ops._chk(x)
    y = torch.empty_like(x)
    L.call("cn_pixel_map", ops._p(x), ops._p(y), x.numel(), mode, ops._stream())
    return y


def u8_to_f32(x):
    x = x.contiguous()
    y = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    L.call("cn_u8_to_f32", x.data_ptr(), ops._p(y), x.numel(), ops._stream())
    return y


def init_stand_in(spec, seed):
    """Seeded stand-ins for the pretrained weights (unavailable offline): He-normal kernels, non-trivial moving
    statistics, so every kernel runs on live activations and the parity tests exercise the folding."""
    from .. import netspec
    p = netspec.init_params(spec, seed, vgg_like=True)
    rng = np.random.RandomState(seed + 1)
    for k in p:
        if k.endswith("/beta"):
            p[k] = (0.1 * rng.standard_normal(p[k].shape)).astype(np.float32)
        elif k.endswith("/gamma"):
            p[k] = rng.uniform(0.8, 1.2, p[k].shape).astype(np.float32)
    return p
