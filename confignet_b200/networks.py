"""ConfigNet networks as functions over parameter dictionaries (names/layouts from netspec.py), every
layer a launch of our sm_100a kernels through ops.py.  Mirrors, function by function, the reference's
Keras models so that the parity tests read like the reference's structure:

  generator_forward            HologanGenerator.call          dnn_models/hologan_generator.py:129-174
  discriminator_forward        HologanDiscriminator.call      dnn_models/hologan_discriminator.py:48-64
  latent_regressor_forward     HologanLatentRegressor.call    dnn_models/hologan_discriminator.py:99-113
  synthetic_encoder_forward    SyntheticDataEncoder.__call__  dnn_models/synthetic_encoder.py:50-60
  latent_discriminator_forward MLPSimple                      confignet_first_stage.py:269-274
  vgg19_activations / perceptual_loss   PerceptualLoss        perceptual_loss.py:43-82
  loss functions                                              losses.py:7-90
"""
from collections import OrderedDict
import numpy as np
import torch
from . import ops
from . import _lib as L
from .netspec import (VGG19_LAYERS, VGG19_USED_LAYER_IDXS, VGG16_LAYERS, VGG16_USED_LAYER_IDXS, RESNET50_STAGES)


# ------------------------------------------------------------------------------------------------ host helpers
def euler_angles_to_matrix_np(angles):
    """confignet_utils.py:122-145 in float32 NumPy (rotations are host inputs in stage 1)."""
    a = np.asarray(angles, np.float32).reshape(-1, 3)
    s, c = np.sin(a), np.cos(a)
    m = np.stack([
        c[:, 2] * c[:, 1], -s[:, 2], c[:, 2] * s[:, 1],
        s[:, 0] * s[:, 1] + c[:, 0] * c[:, 1] * s[:, 2], c[:, 0] * c[:, 2], c[:, 0] * s[:, 2] * s[:, 1] - c[:, 1] * s[:, 0],
        c[:, 1] * s[:, 0] * s[:, 2] - c[:, 0] * s[:, 1], c[:, 2] * s[:, 0], c[:, 0] * c[:, 1] + s[:, 0] * s[:, 1] * s[:, 2],
    ], axis=-1)
    return m.astype(np.float32)


def _as_dev(x, dev, dtype=torch.float32):
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=dtype)
    return torch.as_tensor(np.ascontiguousarray(np.asarray(x)), device="cpu").to(device=dev, dtype=dtype)


# ------------------------------------------------------------------------------------------------ MLPs
def mlp_fused(x, p, prefix, num_layers, alpha):
    """MLPSimple (building_blocks.py:152-173), first-order: LeakyReLU fused in the GEMM epilogue."""
    for i in range(num_layers - 1):
        x = ops.conv_act(x, p["%s/dense%d/kernel" % (prefix, i)], p["%s/dense%d/bias" % (prefix, i)],
                         act=L.ACT_LRELU, alpha=alpha)
    i = num_layers - 1
    return ops.conv_act(x, p["%s/dense%d/kernel" % (prefix, i)], p["%s/dense%d/bias" % (prefix, i)])


def mlp_diff(x, p, prefix, num_layers, alpha):
    """MLPSimple, differentiable to second order (latent discriminator under R1)."""
    for i in range(num_layers - 1):
        x = ops.dense(x, p["%s/dense%d/kernel" % (prefix, i)], p["%s/dense%d/bias" % (prefix, i)])
        x = ops.lrelu(x, alpha)
    i = num_layers - 1
    return ops.dense(x, p["%s/dense%d/kernel" % (prefix, i)], p["%s/dense%d/bias" % (prefix, i)])


# ------------------------------------------------------------------------------------------------ generator
def conv_adain(x, z, p, prefix, upsample, n_mlp_layers=2):
    """Conv{2,3}dAdaIn.call (building_blocks.py:37-44,73-80): [upsample ->] conv -> LeakyReLU(0.3) -> AdaIN.
    The upsample of the PREVIOUS layer output (hologan_generator.py:139-170) is fused into the conv gather,
    the LeakyReLU into the conv epilogue and its derivative into the AdaIN backward kernel."""
    a = ops.conv_act(x, p[prefix + "/conv/kernel"], p[prefix + "/conv/bias"], upsample=upsample,
                     act=L.ACT_LRELU, alpha=0.3, grad_is_preact=True)
    sb = mlp_fused(z, p, prefix + "/adain", n_mlp_layers, alpha=0.2)       # hologan_generator.py:21
    return ops.adain(a, sb, mask_alpha=0.3)


OUTPUT_ACTIVATIONS = {"tanh": L.ACT_TANH, None: L.ACT_NONE, "linear": L.ACT_NONE, "relu": L.ACT_RELU}


def output_activation(name):
    """gen_output_activation (confignet_first_stage.py:44,246; passed by the reference to keras Conv2D(activation=...)) ->
    the epilogue activation of map_final; anything the kernels do not implement fails loudly instead of silently being tanh."""
    if name not in OUTPUT_ACTIVATIONS:
        raise ValueError("gen_output_activation %r is not supported (supported: %s)" % (name, sorted(map(str, OUTPUT_ACTIVATIONS))))
    return OUTPUT_ACTIVATIONS[name]


def generator_forward(p, z, rotation, output_res=256, zs=None, n_mlp_layers=2, out_act=L.ACT_TANH):
    """HologanGenerator.call.  z: (B, latent) device tensor (or ``zs`` = the 5 per-block latents);
    rotation: (B,3) host array / tensor of Euler angles in radians."""
    if zs is None:
        zs = [z] * 5
    dev = zs[0].device
    B = zs[0].shape[0]
    if isinstance(rotation, torch.Tensor) and rotation.is_cuda:
        # predicted / optimised rotations (confignet_second_stage.py:169-170,348): differentiable, on the device
        R = ops.euler_to_matrix(rotation.reshape(-1, 3).to(torch.float32))
    else:
        rot = rotation.detach().cpu().numpy() if isinstance(rotation, torch.Tensor) else rotation
        R = torch.from_numpy(euler_angles_to_matrix_np(rot)).to(dev)
    # Dense(1 -> 32768) applied to zeros (hologan_generator.py:24-27,133-136): the bias, broadcast
    x = p["learned_input/bias"].reshape(1, 4, 4, 4, 512).expand(B, 4, 4, 4, 512)
    x = conv_adain(x, zs[0], p, "map_3d_0", 2, n_mlp_layers)
    x = conv_adain(x, zs[1], p, "map_3d_1", 2, n_mlp_layers)
    x = ops.rotate3d(x, R)
    x = ops.conv_act(x, p["map_3d_post/conv0/kernel"], p["map_3d_post/conv0/bias"], act=L.ACT_LRELU, alpha=0.3)
    x = ops.conv_act(x, p["map_3d_post/conv1/kernel"], p["map_3d_post/conv1/bias"], act=L.ACT_LRELU, alpha=0.3)
    x = x.reshape(B, x.shape[1], x.shape[2], x.shape[3] * x.shape[4])
    x = ops.conv_act(x, p["projection_conv/kernel"], p["projection_conv/bias"], act=L.ACT_LRELU, alpha=0.2)
    x = conv_adain(x, zs[2], p, "map_2d_0", 1, n_mlp_layers)
    x = conv_adain(x, zs[3], p, "map_2d_1", 2, n_mlp_layers)
    x = conv_adain(x, zs[4], p, "map_2d_2", 2, n_mlp_layers)
    if output_res > 128:
        x = conv_adain(x, zs[4], p, "map_2d_2b", 2, n_mlp_layers)
    if output_res > 256:
        x = conv_adain(x, zs[4], p, "map_2d_2c", 2, n_mlp_layers)
    return ops.conv_act(x, p["map_final/kernel"], p["map_final/bias"], upsample=2, act=out_act)


# ------------------------------------------------------------------------------------------------ discriminators
def discr_block(x, p, prefix, return_styles):
    """DiscrBlock.call (building_blocks.py:97-111)."""
    c = ops.conv(x, p[prefix + "/conv/kernel"], p[prefix + "/conv/bias"], stride=2)
    if return_styles:          # both consumers of c as one autograd node: their input gradients leave in one pass
        return ops.discr_norm(c, p[prefix + "/in/gamma"], p[prefix + "/in/beta"], 0.3)
    return ops.lrelu_instance_norm(c, p[prefix + "/in/gamma"], p[prefix + "/in/beta"], 0.3), None


def discriminator_forward(p, img, n_layers=5):
    x = img
    if "initial_1x1_conv/kernel" in p:
        x = ops.conv(x, p["initial_1x1_conv/kernel"], p["initial_1x1_conv/bias"])
    out = OrderedDict()
    for i in range(n_layers):
        x, style = discr_block(x, p, "block%d" % i, True)
        out["discr_style_%d" % i] = ops.dense(style, p["style%d/kernel" % i], p["style%d/bias" % i])
    x = x.reshape(x.shape[0], -1)
    out["discr_final"] = ops.dense(x, p["disc_map/kernel"], p["disc_map/bias"])
    return out


def latent_regressor_forward(p, img, n_layers=5):
    x = img
    if "initial_1x1_conv/kernel" in p:
        x = ops.conv(x, p["initial_1x1_conv/kernel"], p["initial_1x1_conv/bias"])
    for i in range(n_layers):
        x, _ = discr_block(x, p, "block%d" % i, False)
    x = x.reshape(x.shape[0], -1)
    return ops.dense(x, p["latent_predictor/kernel"], p["latent_predictor/bias"])


def synthetic_encoder_forward(p, inputs, facemodel_inputs, num_layers=2):
    names = list(facemodel_inputs.keys())
    if not isinstance(inputs, (list, tuple)):
        cols, used = [], 0
        for n in names:
            d = facemodel_inputs[n][0]
            cols.append(inputs[:, used:used + d]); used += d
        inputs = cols
    outs = [mlp_fused(x, p, "mlp_" + n, num_layers, alpha=0.3) for n, x in zip(names, inputs)]
    return torch.cat(outs, dim=1)


def latent_discriminator_forward(p, z, n_layers=4):
    return mlp_diff(z, p, "mlp", n_layers, alpha=0.3)


# ------------------------------------------------------------------------------------------------ perceptual loss
def _vgg_activations(p, img, layers, used, face):
    x = ops.vgg_preprocess(img, face=face)
    acts = []
    for idx, layer in enumerate(layers, start=1):
        if layer[0] == "conv":
            x = ops.conv_act(x, p[layer[1] + "/kernel"], p[layer[1] + "/bias"], act=L.ACT_RELU)
        else:
            x = ops.maxpool2(x)
        if idx in used:
            acts.append(x)
    return acts


def vgg19_activations(p, img):
    """Activations of Keras VGG19 layers [1,2,8,13] for images in [-1,1] (perceptual_loss.py:43-59)."""
    return _vgg_activations(p, img, VGG19_LAYERS, VGG19_USED_LAYER_IDXS, False)


def vggface_activations(p, img):
    """Activations of Keras VGG16 layers [1,2,8,12] with the VGGFace preprocessing (perceptual_loss.py:26-41,54-56)."""
    return _vgg_activations(p, img, VGG16_LAYERS, VGG16_USED_LAYER_IDXS, True)


def _vgg_tapped_loss(p, img, targets, layers, used, face):
    """The VGG pass of the side that carries the gradient, every used layer tapped by its squared-difference term against
    the other side's activation (ops.conv_act_sqdiff: loss gradient + accumulation + ReLU derivative in one pass)."""
    x = ops.vgg_preprocess(img, face=face)
    total, j = None, 0
    for idx, layer in enumerate(layers, start=1):
        if layer[0] == "conv":
            w, b = p[layer[1] + "/kernel"], p[layer[1] + "/bias"]
            if idx in used:
                t = targets[j]; j += 1
                x, term = ops.conv_act_sqdiff(x, w, b, t, 1.0 / t.numel())
                total = term if total is None else total + term
            else:
                x = ops.conv_act(x, w, b, act=L.ACT_RELU)
        else:
            x = ops.maxpool2(x)
    return total


def perceptual_loss(p_vgg, predicted, data, model_type="imagenet"):
    """PerceptualLoss.loss: sum over the 4 layers of the batch-wide MSE.  Gradient flows to both arguments
    that require it (in ConfigNet only one of them does).  model_type "imagenet" = VGG19, "VGGFace" = VGG16."""
    fwd = vgg19_activations if model_type == "imagenet" else vggface_activations
    layers, used, face = ((VGG19_LAYERS, VGG19_USED_LAYER_IDXS, False) if model_type == "imagenet"
                          else (VGG16_LAYERS, VGG16_USED_LAYER_IDXS, True))
    if predicted.requires_grad != data.requires_grad and torch.is_grad_enabled():
        # ConfigNet's case: one side is a constant.  Its activations first, then the other side with tapped layers.
        live, const = (predicted, data) if predicted.requires_grad else (data, predicted)
        with torch.no_grad():
            targets = fwd(p_vgg, const)
        if all(layers[i - 1][0] == "conv" for i in used) and all(t.numel() % 4 == 0 for t in targets):
            return _vgg_tapped_loss(p_vgg, live, targets, layers, used, face)

    def acts(t):
        if t.requires_grad:
            return fwd(p_vgg, t)
        with torch.no_grad():
            return fwd(p_vgg, t)
    a_p, a_d = acts(predicted), acts(data)
    total = None
    for x, y in zip(a_p, a_d):
        if not x.requires_grad and y.requires_grad:
            x, y = y, x
        term = ops.reduce_sum(x, ops.RED_SQDIFF, y=y, scale=1.0 / x.numel())
        total = term if total is None else total + term
    return total


# ------------------------------------------------------------------------------------------------ losses
def gan_g_loss(scores):
    return ops.reduce_sum(scores, ops.RED_SOFTPLUS, sign=-1.0, scale=1.0 / scores.numel())


def gan_d_loss(label, scores):
    """labels are all-ones (real) or all-zeros (fake) on ConfigNet's path (losses.py:22-23)."""
    return ops.reduce_sum(scores, ops.RED_SOFTPLUS, sign=(-1.0 if label == 1 else 1.0), scale=1.0 / scores.numel())


def eye_loss(gt, gen, eye_masks):
    """losses.py:13-18.  eye_masks: (B,H,W) host array or device tensor.  The per-image 1/(1+sum(mask)) and
    the 1/B of the batch mean are folded into one per-pixel weight so a single weighted reduction kernel
    computes the whole loss (the weight itself is B*H*W floats of bookkeeping)."""
    m = _as_dev(eye_masks, gen.device)
    denom = 1.0 + m.sum(dim=(1, 2), keepdim=True)
    wgt = (m * m / denom / m.shape[0]).contiguous()
    return ops.reduce_sum(gen, ops.RED_SQDIFF, y=gt, wgt=wgt, wdiv=gen.shape[-1])


def gradient_regularization(out, x):
    """losses.py:75-82: 10 * 0.5 * mean_n sum (d sum(out) / dx)^2, kept differentiable."""
    with ops.input_grad_only():
        g, = torch.autograd.grad(out, x, grad_outputs=torch.ones_like(out), create_graph=True)
    return ops.reduce_sum(g, ops.RED_SQ, scale=10 * 0.5 / x.shape[0])


def compute_discriminator_loss(p_d, real, fake, n_layers=5):
    real = real.detach().requires_grad_(True)
    out_real = discriminator_forward(p_d, real, n_layers)
    out_fake = discriminator_forward(p_d, fake.detach(), n_layers)
    losses = OrderedDict()
    for i, o in enumerate(out_real.values()):
        losses["GAN_loss_real_%d" % i] = gan_d_loss(1, o)
    for i, o in enumerate(out_fake.values()):
        losses["GAN_loss_fake_%d" % i] = gan_d_loss(0, o)
    for i, o in enumerate(out_real.values()):
        losses["gp_loss_%d" % i] = gradient_regularization(o, real)
    losses["loss_sum"] = _sum(losses.values())
    return losses


def compute_latent_discriminator_loss(p_ld, real_latents, fake_latents, n_layers=4):
    real = real_latents.detach().requires_grad_(True)
    o_real = latent_discriminator_forward(p_ld, real, n_layers)
    o_fake = latent_discriminator_forward(p_ld, fake_latents.detach(), n_layers)
    losses = OrderedDict()
    losses["GAN_loss_real"] = gan_d_loss(1, o_real)
    losses["GAN_loss_fake"] = gan_d_loss(0, o_fake)
    losses["gp_loss"] = gradient_regularization(o_real, real)
    losses["loss_sum"] = _sum(losses.values())
    return losses


def latent_regression_loss(p_lr, imgs, labels, n_layers=5):
    out = latent_regressor_forward(p_lr, imgs, n_layers)
    return ops.reduce_sum(out, ops.RED_SQDIFF, y=labels, scale=1.0 / out.numel())


def _sum(vals):
    """tf.reduce_sum(list(losses.values())) (confignet_first_stage.py:553, losses.py:44): one stack + one sum instead of a
    chain of ~20 scalar additions (each a launch of its own on the device)."""
    vals = [v.reshape(()) for v in vals]
    return vals[0] if len(vals) == 1 else torch.stack(vals).sum()


# ------------------------------------------------------------------------------------------------ second stage
def resnet50_forward(p, x, prefix="resnet/"):
    """keras-applications ResNet50 (include_top=False, pooling="avg") on 'caffe'-preprocessed images, BatchNorm with
    its moving statistics (real_encoder.py:13,24-27).  Every conv is a launch of the implicit-GEMM family; BN, the
    residual add and the ReLU are one fused kernel."""
    def bn(t, name, residual=None, relu=True):
        q = prefix + name
        return ops.bn_act(t, p[q + "/gamma"], p[q + "/beta"], p[q + "/moving_mean"], p[q + "/moving_variance"], residual, relu)

    def conv(t, name, stride=1, pad=-1):
        q = prefix + name
        return ops.conv_act(t, p[q + "/kernel"], p[q + "/bias"], stride=stride, pad=pad)

    x = conv(x, "conv1_conv", stride=2, pad=3)          # ZeroPadding2D(3) + 7x7/s2 VALID
    x = bn(x, "conv1_bn")
    x = ops.maxpool3s2(x)                               # ZeroPadding2D(1) + 3x3/s2
    for si, (f, blocks, stride) in enumerate(RESNET50_STAGES, start=2):
        for b in range(1, blocks + 1):
            q = "conv%d_block%d" % (si, b)
            s = stride if b == 1 else 1
            shortcut = bn(conv(x, q + "_0_conv", stride=s), q + "_0_bn", relu=False) if b == 1 else x
            y = bn(conv(x, q + "_1_conv", stride=s), q + "_1_bn")
            y = bn(conv(y, q + "_2_conv"), q + "_2_bn")
            x = bn(conv(y, q + "_3_conv"), q + "_3_bn", residual=shortcut, relu=True)
    return ops.global_avg_pool(x)


def real_encoder_forward(p, img, rotation_range_multiplier):
    """RealEncoder.call (real_encoder.py:23-34): img in [-1,1] -> (embedding (B, latent), rotation (B, 3) radians)."""
    x = ops.vgg_preprocess(img)                          # resnet50.preprocess_input = 'caffe' mode
    feat = resnet50_forward(p, x)
    raw = ops.conv_act(feat, p["rotation_regressor/kernel"], p["rotation_regressor/bias"], act=L.ACT_TANH)
    rotation = ops.col_scale(raw, rotation_range_multiplier)
    embedding = ops.conv_act(feat, p["feature_to_latent_mlp/kernel"], p["feature_to_latent_mlp/bias"])
    return embedding, rotation


def gan_d_loss_mixed(scores_zero, scores_one):
    """GAN_D_loss with labels 0 for the first group and 1 for the second (confignet_second_stage.py:163-199):
    the mean over both groups of softplus(s) resp. softplus(-s)."""
    n = scores_zero.numel() + scores_one.numel()
    return (ops.reduce_sum(scores_zero, ops.RED_SOFTPLUS, sign=1.0, scale=1.0 / n) +
            ops.reduce_sum(scores_one, ops.RED_SOFTPLUS, sign=-1.0, scale=1.0 / n))


def normalized_latent_regression_loss(p_lr, imgs, labels, weight, n_layers=5):
    """compute_normalized_latent_regression_loss (confignet_second_stage.py:93-107)."""
    out = latent_regressor_forward(p_lr, imgs, n_layers)
    return ops.norm_latent_loss(out, labels, weight, 3)
