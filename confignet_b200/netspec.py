"""Parameter specifications (names, Keras-layout shapes, initialisers) of the ConfigNet networks.

Pure NumPy.  One table per network, in the order Keras' ``get_weights()`` returns them for the
reference's subclassed models (attribute-assignment order, kernel before bias, gamma before beta).
The order is checked against the reference's own constructors executed under [TF-2.1]'s layer-tracking
rules (scripts/make_golden_models_from_reference.py -> tests/golden/reference_weight_order.json,
tests/test_host_cpu.py::test_weight_order_matches_reference_constructors); the tracking rules themselves
are restated from TensorFlow 2.1 (not installable here), so a released ``models.zip`` (SURVEY.md section
8c) remains the final check.  The real encoder's checkpoint order differs from its table order: see
``real_encoder_keras_order``.

Reference:
  generator      confignet/dnn_models/hologan_generator.py:12-101
  discriminator  confignet/dnn_models/hologan_discriminator.py:10-46
  regressor      confignet/dnn_models/hologan_discriminator.py:66-97
  blocks         confignet/dnn_models/building_blocks.py:11-173
  synth encoder  confignet/dnn_models/synthetic_encoder.py:10-34
  latent discr   confignet/confignet_first_stage.py:269-274
  VGG19          keras.applications.vgg19 (perceptual_loss.py:19-24), truncated at block4_conv2
"""
from collections import OrderedDict
import numpy as np

# Default first-stage latent layout (confignet_first_stage.py:63-76), latent dims only.
DEFAULT_FACEMODEL_LATENT_DIMS = {
    "texture_embedding": 30, "geometry_identity_params": 30, "blendshape_values": 30,
    "beard_style_embedding": 7, "eyebrow_style_embedding": 7, "lower_eyelash_style": 2,
    "upper_eyelash_style": 2, "head_hair_style_embedding": 9, "eye_color": 3,
    "head_hair_color": 3, "hdri_embedding": 20, "bone_rotations:left_eye": 2,
}
# Input dims of the reference's test dataset (tests/test_assets/meta_0000000_000.json via
# neural_renderer_dataset.py:175-222); SURVEY.md section 8d fixes the benchmark to these.
TEST_FACEMODEL_INPUT_DIMS = {
    "beard_style_embedding": 9, "blendshape_values": 62, "bone_rotations:left_eye": 3,
    "eye_color": 4, "eyebrow_style_embedding": 44, "geometry_identity_params": 53,
    "hdri_embedding": 50, "head_hair_color": 3, "head_hair_style_embedding": 18,
    "lower_eyelash_style": 3, "texture_embedding": 50, "upper_eyelash_style": 3,
}


def default_facemodel_inputs():
    """OrderedDict name -> (input_dim, latent_dim), sorted by key (confignet_first_stage.py:115-116)."""
    d = {k: (TEST_FACEMODEL_INPUT_DIMS[k], DEFAULT_FACEMODEL_LATENT_DIMS[k]) for k in DEFAULT_FACEMODEL_LATENT_DIMS}
    return OrderedDict(sorted(d.items(), key=lambda t: t[0]))


def _dense(spec, prefix, n_in, n_out, kernel_init="glorot", bias_init="zeros"):
    spec[prefix + "/kernel"] = ((n_in, n_out), kernel_init)
    spec[prefix + "/bias"] = ((n_out,), bias_init)


def _conv(spec, prefix, ksize, c_in, c_out):
    spec[prefix + "/kernel"] = (tuple(ksize) + (c_in, c_out), "glorot")
    spec[prefix + "/bias"] = ((c_out,), "zeros")


def _mlp(spec, prefix, num_layers, n_in, n_hidden, n_out):
    # building_blocks.py:152-173
    cur = n_in
    for i in range(num_layers - 1):
        _dense(spec, "%s/dense%d" % (prefix, i), cur, n_hidden)
        cur = n_hidden
    _dense(spec, "%s/dense%d" % (prefix, num_layers - 1), cur, n_out)


def generator_spec(latent_dim=145, output_res=256, n_adain_mlp_units=128, n_adain_mlp_layers=2):
    s = OrderedDict()
    _dense(s, "learned_input", 1, 4 * 4 * 4 * 512, kernel_init="zeros", bias_init="ones")

    def conv_adain(name, ksize, c_in, c_out):
        _conv(s, name + "/conv", ksize, c_in, c_out)
        _mlp(s, name + "/adain", n_adain_mlp_layers, latent_dim, n_adain_mlp_units, 2 * c_out)

    conv_adain("map_3d_0", (3, 3, 3), 512, 256)
    conv_adain("map_3d_1", (3, 3, 3), 256, 128)
    _conv(s, "map_3d_post/conv0", (3, 3, 3), 128, 64)
    _conv(s, "map_3d_post/conv1", (3, 3, 3), 64, 64)
    _conv(s, "projection_conv", (1, 1), 16 * 64, 512)
    conv_adain("map_2d_0", (4, 4), 512, 256)
    conv_adain("map_2d_1", (4, 4), 256, 64)
    conv_adain("map_2d_2", (4, 4), 64, 32)
    last = 32
    if output_res > 128:
        conv_adain("map_2d_2b", (4, 4), 32, 32)
    if output_res > 256:
        conv_adain("map_2d_2c", (4, 4), 32, 16)
        last = 16
    _conv(s, "map_final", (4, 4), last, 3)
    return s


def discr_channels(n_layers=5, base=48, max_maps=512):
    return [min(base * (2 ** i), max_maps) for i in range(n_layers)]


def discriminator_spec(output_res=256, n_layers=5, base=48, max_maps=512, ksize=3, from_rgb=True):
    s = OrderedDict()
    if from_rgb:
        _conv(s, "initial_1x1_conv", (1, 1), 3, 3)
    chans = discr_channels(n_layers, base, max_maps)
    c_in = 3
    for i, c in enumerate(chans):
        _conv(s, "block%d/conv" % i, (ksize, ksize), c_in, c)
        s["block%d/in/gamma" % i] = ((c,), "ones")
        s["block%d/in/beta" % i] = ((c,), "zeros")
        c_in = c
    for i, c in enumerate(chans):
        _dense(s, "style%d" % i, 2 * c, 1)
    out = output_res // (2 ** n_layers)
    # hologan_discriminator.py:43: min(e*max//2, max) * out_h * out_w with e = 2**n_layers
    n_lin = min((2 ** n_layers) * max_maps // 2, max_maps) * out * out
    _dense(s, "disc_map", n_lin, 1)
    return s


def latent_regressor_spec(latent_dim=145, output_res=256, n_layers=5, base=48, max_maps=512, ksize=3, from_rgb=True):
    s = OrderedDict()
    if from_rgb:
        _conv(s, "initial_1x1_conv", (1, 1), 3, 3)
    chans = discr_channels(n_layers, base, max_maps)
    c_in = 3
    for i, c in enumerate(chans):
        _conv(s, "block%d/conv" % i, (ksize, ksize), c_in, c)
        s["block%d/in/gamma" % i] = ((c,), "ones")
        s["block%d/in/beta" % i] = ((c,), "zeros")
        c_in = c
    out = output_res // (2 ** n_layers)
    n_lin = min((2 ** n_layers) * max_maps // 2, max_maps) * out * out
    _dense(s, "latent_predictor", n_lin, latent_dim + 3)
    return s


def synthetic_encoder_spec(facemodel_inputs, num_layers=2):
    s = OrderedDict()
    for name, (n_in, n_out) in facemodel_inputs.items():
        _mlp(s, "mlp_" + name, num_layers, n_in, n_in, n_out)
    return s


def latent_discriminator_spec(latent_dim=145, n_layers=4):
    s = OrderedDict()
    _mlp(s, "mlp", n_layers, latent_dim, latent_dim, 1)
    return s


# Keras VGG19 layers up to block4_conv2 (layer idx 13; InputLayer is idx 0).
# ("conv", name, cin, cout) or ("pool", name)
VGG19_LAYERS = [
    ("conv", "block1_conv1", 3, 64), ("conv", "block1_conv2", 64, 64), ("pool", "block1_pool"),
    ("conv", "block2_conv1", 64, 128), ("conv", "block2_conv2", 128, 128), ("pool", "block2_pool"),
    ("conv", "block3_conv1", 128, 256), ("conv", "block3_conv2", 256, 256),
    ("conv", "block3_conv3", 256, 256), ("conv", "block3_conv4", 256, 256), ("pool", "block3_pool"),
    ("conv", "block4_conv1", 256, 512), ("conv", "block4_conv2", 512, 512),
]
VGG19_USED_LAYER_IDXS = [1, 2, 8, 13]   # perceptual_loss.py:21


def vgg19_spec():
    s = OrderedDict()
    for l in VGG19_LAYERS:
        if l[0] == "conv":
            _conv(s, l[1], (3, 3), l[2], l[3])
    return s


def init_params(spec, seed, vgg_like=False):
    """Seeded NumPy initialisation: Glorot-uniform kernels (Keras default), zero biases.

    ``vgg_like`` draws He-normal kernels and small positive biases so a random (non-pretrained)
    VGG / ResNet keeps activations alive through ReLUs; pretrained weights are unavailable offline.
    Names ending in moving_variance / moving_mean get mildly random positive / zero-centred values in
    ``vgg_like`` mode so the inference BatchNorm is exercised with non-trivial statistics.
    """
    rng = np.random.RandomState(seed)
    out = OrderedDict()
    for name, (shape, kind) in spec.items():
        if vgg_like and name.endswith("/moving_variance"):
            a = rng.uniform(0.5, 1.5, shape).astype(np.float32)
        elif vgg_like and name.endswith("/moving_mean"):
            a = (0.1 * rng.standard_normal(shape)).astype(np.float32)
        elif kind == "zeros":
            a = np.zeros(shape, np.float32)
        elif kind == "ones":
            a = np.ones(shape, np.float32)
        elif kind == "glorot":
            recept = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
            fan_in, fan_out = recept * shape[-2], recept * shape[-1]
            if vgg_like:
                a = (rng.standard_normal(shape) * np.sqrt(2.0 / fan_in)).astype(np.float32)
            else:
                lim = np.sqrt(6.0 / (fan_in + fan_out))
                a = rng.uniform(-lim, lim, shape).astype(np.float32)
        else:
            raise ValueError(kind)
        out[name] = a
    return out


def perturb_params(params, seed, scale=0.05):
    """Adds small noise to every array (so zero-initialised biases / kernels are exercised in tests)."""
    rng = np.random.RandomState(seed)
    out = OrderedDict()
    for k, v in params.items():
        out[k] = (v + scale * rng.standard_normal(v.shape)).astype(np.float32)
    return out


# ------------------------------------------------------------------------------------------------
# second stage: RealEncoder = keras-applications ResNet50 (include_top=False, pooling="avg") + two Dense heads
# (dnn_models/real_encoder.py:9-34); VGG16 (VGGFace weights) for the fine-tuning perceptual loss
# (perceptual_loss.py:26-41); LatentGAN MLPs (latent_gan.py:88-109)
# ------------------------------------------------------------------------------------------------
RESNET50_STAGES = [(64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)]     # (filters, blocks, stride of the first block)


def _bn(spec, prefix, c):
    spec[prefix + "/gamma"] = ((c,), "ones")
    spec[prefix + "/beta"] = ((c,), "zeros")
    spec[prefix + "/moving_mean"] = ((c,), "zeros")
    spec[prefix + "/moving_variance"] = ((c,), "ones")


def resnet50_spec():
    """Layer order as keras lists it for ResNet50 v1 (conv / bn pairs; inside a block: 1, 2, then the shortcut
    '0' next to '3').  keras-applications is library code outside /root/reference, so this one order is restated from
    its published model summary, not executed."""
    s = OrderedDict()
    _conv(s, "conv1_conv", (7, 7), 3, 64)
    _bn(s, "conv1_bn", 64)
    c_in = 64
    for si, (f, blocks, _) in enumerate(RESNET50_STAGES, start=2):
        for b in range(1, blocks + 1):
            p = "conv%d_block%d" % (si, b)
            _conv(s, p + "_1_conv", (1, 1), c_in, f); _bn(s, p + "_1_bn", f)
            _conv(s, p + "_2_conv", (3, 3), f, f); _bn(s, p + "_2_bn", f)
            if b == 1:
                _conv(s, p + "_0_conv", (1, 1), c_in, 4 * f)
            _conv(s, p + "_3_conv", (1, 1), f, 4 * f)
            if b == 1:
                _bn(s, p + "_0_bn", 4 * f)
            _bn(s, p + "_3_bn", 4 * f)
            c_in = 4 * f
    return s


def real_encoder_spec(latent_dim=145):
    s = OrderedDict()
    for k, v in resnet50_spec().items():
        s["resnet/" + k] = v
    _dense(s, "rotation_regressor", 2048, 3)
    _dense(s, "feature_to_latent_mlp", 2048, latent_dim)
    return s


def real_encoder_keras_order(latent_dim=145):
    """Names in the order ``RealEncoder.get_weights()`` lists them under the reference's pinned TensorFlow 2.1
    (setup/requirements.txt:8).  [TF-2.1] Network.get_weights concatenates ``layer.weights`` over the tracked
    attributes (resnet, rotation_regressor, feature_to_latent_mlp - real_encoder.py:13-18), and the NESTED ResNet50's
    ``weights`` is ``trainable_weights + non_trainable_weights``: all kernels / biases / gammas / betas in layer order
    first, then every BatchNormalization's moving mean / variance.  The flat HBM layout keeps real_encoder_spec's
    per-layer order; only the checkpoint interchange (get_weights / set_weights) is permuted.  Checked against the
    reference's executed constructor in tests/golden/reference_weight_order.json."""
    names = list(real_encoder_spec(latent_dim).keys())
    resnet = [k for k in names if k.startswith("resnet/")]
    heads = [k for k in names if not k.startswith("resnet/")]
    return [k for k in resnet if is_trainable(k)] + [k for k in resnet if not is_trainable(k)] + heads


def is_trainable(name):
    """keras: BatchNormalization moving statistics are non-trainable weights."""
    return not (name.endswith("/moving_mean") or name.endswith("/moving_variance"))


# Keras VGG16 layers up to block4_conv2 (layer idx 12; InputLayer is idx 0)
VGG16_LAYERS = [
    ("conv", "block1_conv1", 3, 64), ("conv", "block1_conv2", 64, 64), ("pool", "block1_pool"),
    ("conv", "block2_conv1", 64, 128), ("conv", "block2_conv2", 128, 128), ("pool", "block2_pool"),
    ("conv", "block3_conv1", 128, 256), ("conv", "block3_conv2", 256, 256), ("conv", "block3_conv3", 256, 256),
    ("pool", "block3_pool"),
    ("conv", "block4_conv1", 256, 512), ("conv", "block4_conv2", 512, 512),
]
VGG16_USED_LAYER_IDXS = [1, 2, 8, 12]   # perceptual_loss.py:35


def vgg16_spec():
    s = OrderedDict()
    for l in VGG16_LAYERS:
        if l[0] == "conv":
            _conv(s, l[1], (3, 3), l[2], l[3])
    return s


def latent_gan_mlp_spec(latent_dim, num_layers=3, hidden_multiplier=1.5, num_out=None):
    """MLPSimple(num_layers, latent_dim, int(latent_dim*multiplier), num_out) (latent_gan.py:88-109)."""
    s = OrderedDict()
    _mlp(s, "mlp", num_layers, latent_dim, int(latent_dim * hidden_multiplier), latent_dim if num_out is None else num_out)
    return s


def init_real_encoder_params(latent_dim, seed):
    """Random stand-in for the ImageNet ResNet50 + freshly initialised heads (pretrained weights are unavailable
    offline): He-normal kernels damped on the residual branches so activations neither die nor explode over the
    16 blocks, non-trivial moving statistics, small head kernels so the tanh rotation head is not saturated."""
    p = init_params(real_encoder_spec(latent_dim), seed, vgg_like=True)
    for k in p:
        if k.endswith("_3_conv/kernel"):
            p[k] = (p[k] * 0.25).astype(np.float32)
        elif k.endswith("conv1_conv/kernel"):
            p[k] = (p[k] * 0.01).astype(np.float32)          # 'caffe' inputs are O(100)
        elif k.endswith("_conv/kernel"):
            p[k] = (p[k] * 0.8).astype(np.float32)
    p["rotation_regressor/kernel"] = (p["rotation_regressor/kernel"] * 0.5).astype(np.float32)
    return p
