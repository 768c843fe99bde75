"""CPU ORACLE - TEST INFRASTRUCTURE ONLY.  Never imported by the product path (confignet_b200/).

A torch-CPU restatement of the reference's training hot path (generator, discriminators,
latent regressor, synthetic encoder, VGG19 perceptual loss, losses, Keras-Adam, EMA and the
D / G training steps), following the reference files cited on each function plus the
TensorFlow-2.1 semantics listed in SURVEY.md section 8c (marked [TF-2.1] below).

PINNING: TensorFlow 2.1 is not installable in this environment and the released ``models/`` directory (needed by
every golden .npz in /root/reference/tests/test_assets) is absent, so parity is UNPINNED AGAINST TENSORFLOW ITSELF.
The oracle is pinned instead to the reference's own code EXECUTED here on stand-ins for TensorFlow (scripts/
make_golden_*_from_reference.py, vectors under tests/golden/, checked by tests/test_host_cpu.py): the host logic
bit-exactly; transform_3d_grid_tf, euler_angles_to_matrix, get_layer_style, the loss formulas and
InstanceNormalization.call to 1e-12; HologanGenerator / HologanDiscriminator / HologanLatentRegressor / MLPSimple /
SyntheticDataEncoder and compute_discriminator_loss (nested tape, R1) to 1e-13; the discriminator, synth-discriminator,
latent-discriminator and generator training steps incl. the Adam update to 3e-13.  The stand-ins restate (do not
execute) the elementary tf / Keras semantics marked [TF-2.1] below and the keras-applications networks.  Further
checks: NumPy-loop restatements of conv-SAME / rotate / norms on tiny shapes, invariants readable from the reference
source, fp64 autograd of the hand-derived formulas.

Gradients come from torch.autograd on CPU (fp32 for the timed baseline, fp64 as gold).
Allowed importers: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference).
"""
from collections import OrderedDict
import math
import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------- helpers


def to_torch(params, dtype=torch.float32, requires_grad=False):
    out = OrderedDict()
    for k, v in params.items():
        t = torch.as_tensor(np.asarray(v)).to(dtype).clone()
        t.requires_grad_(requires_grad)
        out[k] = t
    return out


def same_pad(in_size, k, s):
    """[TF-2.1] padding='same': total = max((ceil(in/s)-1)*s + k - in, 0); before = total//2."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    return total // 2, total - total // 2


def conv_same(x, kernel, bias=None, stride=1):
    """Keras Conv2D/Conv3D, padding='same', channels-last.  x: (N,*spatial,Cin); kernel: (*k,Cin,Cout).
    Cross-correlation [TF-2.1]."""
    nd = x.dim() - 2
    perm_in = (0, nd + 1) + tuple(range(1, nd + 1))
    xc = x.permute(*perm_in)
    w = kernel.permute(nd + 1, nd, *range(nd))
    pads = []
    for d in reversed(range(nd)):          # F.pad wants last dim first
        b, a = same_pad(x.shape[1 + d], kernel.shape[d], stride)
        pads += [b, a]
    xc = F.pad(xc, pads)
    y = (F.conv2d if nd == 2 else F.conv3d)(xc, w, bias, stride=stride)
    perm_out = (0,) + tuple(range(2, nd + 2)) + (1,)
    return y.permute(*perm_out)


def upsample_nearest2(x):
    """keras.layers.UpSampling2D/3D default (nearest, x2) on channels-last."""
    for d in range(1, x.dim() - 1):
        x = x.repeat_interleave(2, dim=d)
    return x


def lrelu(x, alpha):
    return torch.where(x >= 0, x, x * alpha)


def dense(x, p, prefix):
    return x @ p[prefix + "/kernel"] + p[prefix + "/bias"]


def mlp_simple(x, p, prefix, num_layers, alpha, alpha_last=None):
    """building_blocks.py:152-173."""
    for i in range(num_layers - 1):
        x = lrelu(dense(x, p, "%s/dense%d" % (prefix, i)), alpha)
    x = dense(x, p, "%s/dense%d" % (prefix, num_layers - 1))
    if alpha_last is not None:
        x = lrelu(x, alpha_last)
    return x


# ----------------------------------------------------------------------------- generator


def adain(x, z, p, prefix, n_mlp_layers=2):
    """building_blocks.py:135-149.  LayerNormalization(axis=spatial, center=False, scale=False):
    [TF-2.1] non-fused path = nn.moments + nn.batch_normalization, eps 1e-3, population variance."""
    c = x.shape[-1]
    zz = mlp_simple(z, p, prefix, n_mlp_layers, alpha=0.2)          # hologan_generator.py:21
    axes = tuple(range(1, x.dim() - 1))
    mean = x.mean(dim=axes, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=axes, keepdim=True)
    xn = (x - mean) * torch.rsqrt(var + 1e-3)
    bshape = (x.shape[0],) + (1,) * (x.dim() - 2) + (c,)
    scale = zz[:, :c].reshape(bshape)
    bias = zz[:, c:].reshape(bshape)
    return xn * (scale + 1) + bias


def conv_adain(x, z, p, prefix):
    """Conv{2,3}dAdaIn.call building_blocks.py:37-44,73-80: conv -> LeakyReLU(0.3) -> AdaIN."""
    x = conv_same(x, p[prefix + "/conv/kernel"], p[prefix + "/conv/bias"])
    x = lrelu(x, 0.3)
    return adain(x, z, p, prefix + "/adain")


def euler_angles_to_matrix(angles):
    """confignet_utils.py:122-145."""
    angles = angles.reshape(-1, 3)
    s, c = torch.sin(angles), torch.cos(angles)
    a11 = c[:, 2] * c[:, 1]
    a12 = -s[:, 2]
    a13 = c[:, 2] * s[:, 1]
    a21 = s[:, 0] * s[:, 1] + c[:, 0] * c[:, 1] * s[:, 2]
    a22 = c[:, 0] * c[:, 2]
    a23 = c[:, 0] * s[:, 2] * s[:, 1] - c[:, 1] * s[:, 0]
    a31 = c[:, 1] * s[:, 0] * s[:, 2] - c[:, 0] * s[:, 1]
    a32 = c[:, 2] * s[:, 0]
    a33 = c[:, 0] * c[:, 1] + s[:, 0] * s[:, 1] * s[:, 2]
    return torch.stack([a11, a12, a13, a21, a22, a23, a31, a32, a33], dim=-1).reshape(-1, 3, 3)


def transform_3d_grid(grid, transform):
    """confignet_utils.py:63-120: trilinear resampling of (B,S,S,S,C) under a per-sample 3x3."""
    B, S = grid.shape[0], grid.shape[1]
    center = (S - 1) / 2
    r = torch.arange(S, dtype=grid.dtype)
    xs, ys, zs = torch.meshgrid(r, r, r, indexing="ij")
    coords = torch.stack([xs.flatten(), ys.flatten(), zs.flatten()])          # (3, S^3)
    tc = transform.to(grid.dtype) @ (coords - center)[None] + center          # (B,3,S^3)
    tc = tc.clamp(0, S - 1)
    fl = tc.floor().clamp(0, S - 1)
    ce = (fl + 1).clamp(0, S - 1)
    fi, ci = fl.long(), ce.long()
    diffs = (tc - fl).unsqueeze(-1)                                            # (B,3,S^3,1)
    b = torch.arange(B)[:, None].expand(B, S ** 3)

    def g(i0, i1, i2):
        return grid[b, i0, i1, i2]                                             # (B,S^3,C)
    c000 = g(fi[:, 0], fi[:, 1], fi[:, 2]); c100 = g(ci[:, 0], fi[:, 1], fi[:, 2])
    c101 = g(ci[:, 0], fi[:, 1], ci[:, 2]); c001 = g(fi[:, 0], fi[:, 1], ci[:, 2])
    c010 = g(fi[:, 0], ci[:, 1], fi[:, 2]); c110 = g(ci[:, 0], ci[:, 1], fi[:, 2])
    c111 = g(ci[:, 0], ci[:, 1], ci[:, 2]); c011 = g(fi[:, 0], ci[:, 1], ci[:, 2])
    d0, d1, d2 = diffs[:, 0], diffs[:, 1], diffs[:, 2]
    c00 = c000 * (1 - d0) + c100 * d0
    c01 = c001 * (1 - d0) + c101 * d0
    c10 = c010 * (1 - d0) + c110 * d0
    c11 = c011 * (1 - d0) + c111 * d0
    c0 = c00 * (1 - d1) + c10 * d1
    c1 = c01 * (1 - d1) + c11 * d1
    out = c0 * (1 - d2) + c1 * d2
    return out.reshape(grid.shape)


def generator_forward(p, z, rotation, output_res=256, zs=None):
    """HologanGenerator.call hologan_generator.py:129-174.  ``zs`` optionally gives the 5 per-block
    latents (z_3d_0, z_3d_1, z_2d_0, z_2d_1, z_2d_2) of build_input_dict :109-127."""
    if zs is None:
        zs = [z] * 5
    B = zs[0].shape[0]
    zeros = torch.zeros(B, 1, dtype=zs[0].dtype)
    x = dense(zeros, p, "learned_input").reshape(B, 4, 4, 4, 512)
    x = upsample_nearest2(x)
    x = conv_adain(x, zs[0], p, "map_3d_0")
    x = upsample_nearest2(x)
    x = conv_adain(x, zs[1], p, "map_3d_1")
    x = transform_3d_grid(x, euler_angles_to_matrix(rotation))
    x = lrelu(conv_same(x, p["map_3d_post/conv0/kernel"], p["map_3d_post/conv0/bias"]), 0.3)
    x = lrelu(conv_same(x, p["map_3d_post/conv1/kernel"], p["map_3d_post/conv1/bias"]), 0.3)
    x = x.reshape(B, x.shape[1], x.shape[2], x.shape[3] * x.shape[4])
    x = lrelu(conv_same(x, p["projection_conv/kernel"], p["projection_conv/bias"]), 0.2)   # tf.nn.leaky_relu
    x = conv_adain(x, zs[2], p, "map_2d_0")
    x = upsample_nearest2(x)
    x = conv_adain(x, zs[3], p, "map_2d_1")
    x = upsample_nearest2(x)
    x = conv_adain(x, zs[4], p, "map_2d_2")
    x = upsample_nearest2(x)
    if output_res > 128:
        x = conv_adain(x, zs[4], p, "map_2d_2b")
        x = upsample_nearest2(x)
    if output_res > 256:
        x = conv_adain(x, zs[4], p, "map_2d_2c")
        x = upsample_nearest2(x)
    x = torch.tanh(conv_same(x, p["map_final/kernel"], p["map_final/bias"]))
    return x


def to_uint8_images(imgs):
    """generate_images post-process confignet_first_stage.py:636-637 (clip, scale, truncate)."""
    imgs = np.clip(np.asarray(imgs), -1.0, 1.0)
    return ((imgs + 1) * 127.5).astype(np.uint8)


# ----------------------------------------------------------------------------- discriminator


def layer_style(x, eps=1e-6):
    """get_layer_style confignet_utils.py:147-159 -> concat(mean, std) flattened to (B, 2C)."""
    axes = tuple(range(1, x.dim() - 1))
    mean = x.mean(dim=axes)
    std = torch.sqrt(((x - x.mean(dim=axes, keepdim=True)) ** 2).mean(dim=axes) + eps)
    return torch.cat([mean, std], dim=-1)


def instance_norm_std(x, gamma, beta, eps=1e-3):
    """InstanceNormalization.call instance_normalization.py:108-131: eps added to the STD."""
    axes = tuple(range(1, x.dim() - 1))
    mean = x.mean(dim=axes, keepdim=True)
    std = torch.sqrt(((x - mean) ** 2).mean(dim=axes, keepdim=True)) + eps     # K.std = population
    return (x - mean) / std * gamma + beta


def discr_block(x, p, prefix, return_styles):
    """DiscrBlock.call building_blocks.py:97-111."""
    x = conv_same(x, p[prefix + "/conv/kernel"], p[prefix + "/conv/bias"], stride=2)
    style = layer_style(x) if return_styles else None
    x = lrelu(x, 0.3)
    x = instance_norm_std(x, p[prefix + "/in/gamma"], p[prefix + "/in/beta"])
    return x, style


def discriminator_forward(p, img, n_layers=5):
    """HologanDiscriminator.call hologan_discriminator.py:48-64 -> OrderedDict of 6 logits (B,1)."""
    x = img
    if "initial_1x1_conv/kernel" in p:
        x = conv_same(x, p["initial_1x1_conv/kernel"], p["initial_1x1_conv/bias"])
    out = OrderedDict()
    for i in range(n_layers):
        x, style = discr_block(x, p, "block%d" % i, True)
        out["discr_style_%d" % i] = dense(style, p, "style%d" % i)
    x = x.reshape(x.shape[0], -1)
    out["discr_final"] = dense(x, p, "disc_map")
    return out


def latent_regressor_forward(p, img, n_layers=5):
    """HologanLatentRegressor.call hologan_discriminator.py:99-113."""
    x = img
    if "initial_1x1_conv/kernel" in p:
        x = conv_same(x, p["initial_1x1_conv/kernel"], p["initial_1x1_conv/bias"])
    for i in range(n_layers):
        x, _ = discr_block(x, p, "block%d" % i, False)
    x = x.reshape(x.shape[0], -1)
    return dense(x, p, "latent_predictor")


def synthetic_encoder_forward(p, inputs, facemodel_inputs, num_layers=2):
    """SyntheticDataEncoder.__call__ synthetic_encoder.py:36-60.  ``inputs``: list (one array per
    parameter, in facemodel_inputs order) or one concatenated matrix (split by input dims)."""
    names = list(facemodel_inputs.keys())
    if not isinstance(inputs, (list, tuple)):
        cols, used = [], 0
        for n in names:
            d = facemodel_inputs[n][0]
            cols.append(inputs[:, used:used + d]); used += d
        inputs = cols
    outs = [mlp_simple(x, p, "mlp_" + n, num_layers, alpha=0.3) for n, x in zip(names, inputs)]
    return torch.cat(outs, dim=1)


def latent_discriminator_forward(p, z, n_layers=4):
    """MLPSimple(4, latent, latent, 1, LeakyReLU(0.3)) confignet_first_stage.py:269-274."""
    return mlp_simple(z, p, "mlp", n_layers, alpha=0.3)


def facemodel_param_idxs_in_latent(facemodel_inputs, name):
    """get_facemodel_param_idxs_in_latent confignet_first_stage.py:217-227 (integer, bit-exact)."""
    names = list(facemodel_inputs.keys())
    dims = list(facemodel_inputs.values())
    i = names.index(name)
    start = int(np.sum([x[1] for x in dims[:i]]))
    return range(start, start + dims[i][1])


# ----------------------------------------------------------------------------- perceptual loss

VGG19_LAYERS = [
    ("conv", "block1_conv1"), ("conv", "block1_conv2"), ("pool", "block1_pool"),
    ("conv", "block2_conv1"), ("conv", "block2_conv2"), ("pool", "block2_pool"),
    ("conv", "block3_conv1"), ("conv", "block3_conv2"), ("conv", "block3_conv3"),
    ("conv", "block3_conv4"), ("pool", "block3_pool"),
    ("conv", "block4_conv1"), ("conv", "block4_conv2"),
]
VGG19_USED = [1, 2, 8, 13]          # Keras layer indices, InputLayer = 0 (perceptual_loss.py:21)
CAFFE_MEAN_BGR = (103.939, 116.779, 123.68)


def vgg19_preprocess(img):
    """perceptual_loss.py:50-59: (x+1)*127.5 then keras vgg19.preprocess_input ('caffe' mode
    [TF-2.1]: reverse channel order, subtract BGR means)."""
    x = (img + 1) * 127.5
    x = x.flip(-1)
    return x - torch.tensor(CAFFE_MEAN_BGR, dtype=img.dtype)


def vgg19_activations(p, x):
    acts = []
    for idx, (kind, name) in enumerate(VGG19_LAYERS, start=1):
        if kind == "conv":
            x = torch.relu(conv_same(x, p[name + "/kernel"], p[name + "/bias"]))
        else:
            x = F.max_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
        if idx in VGG19_USED:
            acts.append(x)
    return acts


def perceptual_loss(p_vgg, predicted, data):
    """PerceptualLoss.loss perceptual_loss.py:43-82: sum over 4 layers of the batch-wide MSE."""
    a_p = vgg19_activations(p_vgg, vgg19_preprocess(predicted))
    a_d = vgg19_activations(p_vgg, vgg19_preprocess(data))
    total = 0
    for x, y in zip(a_p, a_d):
        total = total + ((x.reshape(-1) - y.reshape(-1)) ** 2).mean()
    return total


# ----------------------------------------------------------------------------- losses


def gan_g_loss(scores):
    """losses.py:7-8."""
    return F.softplus(-scores).mean()


def gan_d_loss(labels, scores):
    """losses.py:10-11."""
    return (labels * F.softplus(-scores) + (1.0 - labels) * F.softplus(scores)).mean()


def eye_loss(gt, gen, eye_masks):
    """losses.py:13-18.  eye_masks: (B,H,W) numeric array."""
    m = torch.as_tensor(np.asarray(eye_masks)).to(gt.dtype)
    diff = (gt - gen) * m.unsqueeze(-1)
    per_img = (diff ** 2).sum(dim=(1, 2, 3)) / (1 + m.sum(dim=(1, 2)))
    return per_img.mean()


def gradient_regularization(out, x):
    """losses.py:75-82.  tape.gradient(out, x) == d(sum(out))/dx; kept differentiable (R1 is
    inside the outer tape, confignet_first_stage.py:469-472)."""
    g, = torch.autograd.grad(out.sum(), x, create_graph=True)
    r1 = (g ** 2).sum(dim=tuple(range(1, g.dim()))).mean()
    return 10 * 0.5 * r1


def compute_discriminator_loss(p_d, real, fake, n_layers=5):
    """losses.py:20-47 -> OrderedDict of 18 terms + loss_sum."""
    real = real.detach().clone().requires_grad_(True)
    out_real = discriminator_forward(p_d, real, n_layers)
    out_fake = discriminator_forward(p_d, fake, n_layers)
    losses = OrderedDict()
    ones = torch.ones(real.shape[0], 1, dtype=real.dtype)
    zeros = torch.zeros(fake.shape[0], 1, dtype=real.dtype)
    for i, o in enumerate(out_real.values()):
        losses["GAN_loss_real_%d" % i] = gan_d_loss(ones, o)
    for i, o in enumerate(out_fake.values()):
        losses["GAN_loss_fake_%d" % i] = gan_d_loss(zeros, o)
    for i, o in enumerate(out_real.values()):
        losses["gp_loss_%d" % i] = gradient_regularization(o, real)
    losses["loss_sum"] = sum(losses.values())
    return losses


def compute_latent_discriminator_loss(p_ld, real_latents, fake_latents, n_layers=4):
    """losses.py:49-73."""
    real = real_latents.detach().clone().requires_grad_(True)
    o_real = latent_discriminator_forward(p_ld, real, n_layers)
    o_fake = latent_discriminator_forward(p_ld, fake_latents, n_layers)
    losses = OrderedDict()
    losses["GAN_loss_real"] = gan_d_loss(torch.ones_like(o_real), o_real)
    losses["GAN_loss_fake"] = gan_d_loss(torch.zeros_like(o_fake), o_fake)
    losses["gp_loss"] = gradient_regularization(o_real, real)
    losses["loss_sum"] = sum(losses.values())
    return losses


def latent_regression_loss(p_lr, imgs, labels, n_layers=5):
    """losses.py:85-90 ([TF-2.1] mean_squared_error = mean over last axis, then mean)."""
    out = latent_regressor_forward(p_lr, imgs, n_layers)
    return ((labels - out) ** 2).mean(dim=-1).mean()


# ----------------------------------------------------------------------------- optimiser / EMA


class KerasAdam:
    """[TF-2.1] keras.optimizers.Adam (non-amsgrad): t = iterations+1;
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v updates; theta -= lr_t*m/(sqrt(v)+eps), eps=1e-7.
    ``iterations`` is per optimizer object: the three discriminators share one
    (confignet_first_stage.py:601,608-610)."""

    def __init__(self, lr=0.0004, beta_1=0.0, beta_2=0.9, epsilon=1e-7):
        self.lr, self.b1, self.b2, self.eps = lr, beta_1, beta_2, epsilon
        self.iterations = 0
        self.state = {}

    def apply_gradients(self, grads_and_vars):
        t = self.iterations + 1
        lr_t = self.lr * math.sqrt(1 - self.b2 ** t) / (1 - self.b1 ** t)
        with torch.no_grad():
            for g, v in grads_and_vars:
                st = self.state.setdefault(id(v), (torch.zeros_like(v), torch.zeros_like(v)))
                m, vv = st
                m.mul_(self.b1).add_(g, alpha=1 - self.b1)
                vv.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                v.sub_(lr_t * m / (vv.sqrt() + self.eps))
        self.iterations += 1


def update_smoothed_weights(p_smoothed, p_train, alpha=0.999):
    """confignet_first_stage.py:393-400."""
    with torch.no_grad():
        for k in p_smoothed:
            p_smoothed[k].copy_(alpha * p_smoothed[k] + (1 - alpha) * p_train[k])


# ----------------------------------------------------------------------------- training steps

DEFAULT_LOSS_WEIGHTS = dict(image_loss_weight=0.00005, eye_loss_weight=5.0,
                            domain_adverserial_loss_weight=5.0, latent_regression_weight=10.0,
                            latent_regressor_rot_weight=5.0)


def discriminator_step_losses(p_d, p_g, real_imgs, latents, rotations, output_res=256):
    """discriminator_training_step loss part, confignet_first_stage.py:438-450,466-470:
    fakes come from G outside the tape."""
    with torch.no_grad():
        fake = generator_forward(p_g, latents, rotations, output_res)
    return compute_discriminator_loss(p_d, real_imgs, fake)


def synth_discriminator_step_losses(p_sd, p_g, p_se, facemodel_inputs, real_imgs, facemodel_params,
                                    rotations, output_res=256):
    """synth_discriminator_training_step confignet_first_stage.py:452-464,478-482."""
    with torch.no_grad():
        z = synthetic_encoder_forward(p_se, facemodel_params, facemodel_inputs)
        fake = generator_forward(p_g, z, rotations, output_res)
    return compute_discriminator_loss(p_sd, real_imgs, fake)


def latent_discriminator_step_losses(p_ld, p_se, facemodel_inputs, real_latents, facemodel_params):
    """latent_discriminator_training_step confignet_first_stage.py:490-498."""
    with torch.no_grad():
        fake = synthetic_encoder_forward(p_se, facemodel_params, facemodel_inputs)
    return compute_latent_discriminator_loss(p_ld, real_latents, fake)


def generator_step_losses(p_g, p_lr, p_se, p_d, p_sd, p_ld, p_vgg, facemodel_inputs, batch,
                          weights=None, output_res=256):
    """generator_training_step (stage 1) confignet_first_stage.py:506-554.
    batch: dict with facemodel_params (list), synth_rotations, gt_imgs (float, [-1,1]),
    eye_masks, real_latents, real_rotations (torch tensors / arrays)."""
    w = dict(DEFAULT_LOSS_WEIGHTS); w.update(weights or {})
    losses = OrderedDict()
    synth_latents = synthetic_encoder_forward(p_se, batch["facemodel_params"], facemodel_inputs)
    out_synth = generator_forward(p_g, synth_latents, batch["synth_rotations"], output_res)
    out_real = generator_forward(p_g, batch["real_latents"], batch["real_rotations"], output_res)
    losses["image_loss"] = w["image_loss_weight"] * perceptual_loss(p_vgg, batch["gt_imgs"], out_synth)
    losses["eye_loss"] = w["eye_loss_weight"] * eye_loss(batch["gt_imgs"], out_synth, batch["eye_masks"])
    for i, o in enumerate(discriminator_forward(p_sd, out_synth).values()):
        losses["GAN_loss_synth_%d" % i] = gan_g_loss(o)
    for i, o in enumerate(discriminator_forward(p_d, out_real).values()):
        losses["GAN_loss_real_%d" % i] = gan_g_loss(o)
    losses["latent_GAN_loss"] = w["domain_adverserial_loss_weight"] * gan_g_loss(
        latent_discriminator_forward(p_ld, synth_latents))
    stacked_latents = torch.cat([synth_latents, batch["real_latents"]], dim=0)
    stacked_imgs = torch.cat([out_synth, out_real], dim=0)
    stacked_rot = torch.cat([batch["synth_rotations"], batch["real_rotations"]], dim=0)
    labels = torch.cat([stacked_latents, w["latent_regressor_rot_weight"] * stacked_rot], dim=-1)
    losses["latent_regression_loss"] = w["latent_regression_weight"] * latent_regression_loss(p_lr, stacked_imgs, labels)
    losses["loss_sum"] = sum(losses.values())
    return losses


def grads_of(loss, params):
    """tape.gradient(loss, trainable_weights): list aligned with params.values() (None -> zeros)."""
    ps = list(params.values())
    gs = torch.autograd.grad(loss, ps, allow_unused=True)
    return [torch.zeros_like(p) if g is None else g for g, p in zip(gs, ps)]


# ----------------------------------------------------------------------------- whole training iteration (CPU)


class OracleFirstStage:
    """CPU restatement of the first-stage step loop (confignet_first_stage.py:466-560,597-617) used as the
    timed CPU baseline (bench.py) - same step functions, same synthetic inputs, torch-CPU fp32."""

    def __init__(self, facemodel_inputs, output_res=256, seed=1234, dtype=torch.float32):
        from confignet_b200 import netspec          # parameter tables only (pure NumPy, no CUDA)
        self.fm, self.res, self.dtype = facemodel_inputs, output_res, dtype
        mk = lambda spec, s, **kw: to_torch(netspec.init_params(spec, s, **kw), dtype=dtype, requires_grad=True)
        self.p_se = mk(netspec.synthetic_encoder_spec(facemodel_inputs, 2), seed + 1)
        self.p_d = mk(netspec.discriminator_spec(output_res), seed + 2)
        self.p_sd = mk(netspec.discriminator_spec(output_res), seed + 3)
        self.p_ld = mk(netspec.latent_discriminator_spec(145, 4), seed + 4)
        self.p_lr = mk(netspec.latent_regressor_spec(145, output_res), seed + 5)
        self.p_g = mk(netspec.generator_spec(145, output_res), seed + 6)
        self.p_gs = to_torch({k: v.detach().numpy() for k, v in self.p_g.items()}, dtype=dtype)
        self.p_vgg = to_torch(netspec.init_params(netspec.vgg19_spec(), seed + 7, vgg_like=True), dtype=dtype)
        self.d_opt, self.g_opt = KerasAdam(), KerasAdam()

    def _t(self, a):
        return torch.as_tensor(np.asarray(a)).to(self.dtype)

    def discriminator_step(self, real_u8, latents, rotations):
        real = self._t(real_u8) / 127.5 - 1.0
        losses = discriminator_step_losses(self.p_d, self.p_g, real, self._t(latents), self._t(rotations), self.res)
        self.d_opt.apply_gradients(zip(grads_of(losses["loss_sum"], self.p_d), self.p_d.values()))
        return losses

    def synth_discriminator_step(self, real_u8, facemodel_params, rotations):
        """synth_discriminator_training_step confignet_first_stage.py:452-464,478-488 (shared discriminator optimizer)"""
        real = self._t(real_u8) / 127.5 - 1.0
        losses = synth_discriminator_step_losses(self.p_sd, self.p_g, self.p_se, self.fm, real,
                                                 [self._t(a) for a in facemodel_params], self._t(rotations), self.res)
        self.d_opt.apply_gradients(zip(grads_of(losses["loss_sum"], self.p_sd), self.p_sd.values()))
        return losses

    def latent_discriminator_step(self, real_latents, facemodel_params):
        """latent_discriminator_training_step confignet_first_stage.py:490-504 (shared discriminator optimizer)"""
        losses = latent_discriminator_step_losses(self.p_ld, self.p_se, self.fm, self._t(real_latents),
                                                  [self._t(a) for a in facemodel_params])
        self.d_opt.apply_gradients(zip(grads_of(losses["loss_sum"], self.p_ld), self.p_ld.values()))
        return losses

    def generate_images(self, latents, rotations):
        """generate_images confignet_first_stage.py:633-639: smoothed generator, clip, truncating uint8 cast"""
        with torch.no_grad():
            imgs = generator_forward(self.p_gs, self._t(latents), self._t(rotations), self.res)
        return to_uint8_images(imgs.numpy())

    def generator_step(self, facemodel_params, synth_rot, gt_u8, eye_masks, real_latents, real_rot):
        batch = dict(facemodel_params=[self._t(a) for a in facemodel_params], synth_rotations=self._t(synth_rot),
                     gt_imgs=self._t(gt_u8) / 127.5 - 1.0, eye_masks=eye_masks, real_latents=self._t(real_latents),
                     real_rotations=self._t(real_rot))
        losses = generator_step_losses(self.p_g, self.p_lr, self.p_se, self.p_d, self.p_sd, self.p_ld, self.p_vgg,
                                       self.fm, batch, output_res=self.res)
        allp = OrderedDict()
        for pre, p in (("g/", self.p_g), ("lr/", self.p_lr), ("se/", self.p_se)):
            for k, v in p.items():
                allp[pre + k] = v
        self.g_opt.apply_gradients(zip(grads_of(losses["loss_sum"], allp), allp.values()))
        update_smoothed_weights(self.p_gs, self.p_g)
        return losses
