"""CPU ORACLE (part 2) - TEST INFRASTRUCTURE ONLY.  Never imported by the product path (confignet_b200/).

torch-CPU restatement of the second-stage / fine-tuning / LatentGAN pieces of the reference:

  resnet50_forward, real_encoder_forward   dnn_models/real_encoder.py:9-34 + keras-applications ResNet50 v1 [TF-2.1]
  vggface_*                                perceptual_loss.py:26-41,54-56 (VGG16 truncated at layer 12)
  normalized_latent_regression_loss        confignet_second_stage.py:93-107
  stage2_generator_step_losses             confignet_second_stage.py:149-218
  stage2_discriminator_step_losses         confignet_first_stage.py:466-476 + confignet_second_stage.py:119-130
  stage2_latent_discriminator_step_losses  confignet_second_stage.py:132-147
  fine_tune_losses                         confignet_second_stage.py:360-390
  latent_gan_*                             latent_gan.py:117-165

PINNING (see oracle/confignet_oracle.py): unpinned against TensorFlow itself; normalized_regression, the stage-2
latent-discriminator / generator steps (with a stand-in encoder) and the latent_gan_* steps are pinned to the reference's own code executed on TensorFlow stand-ins (tests/golden/reference_float_logic.npz,
reference_steps.npz); the ResNet50 / VGG16 pieces (keras-applications, not reference code) restate SURVEY.md section 8c
items 9-11 and are pinned (<= 1e-9, fp64) to torchvision's ResNet50 (strides moved to the v1 position) and VGG16 run here
on the same weights (tests/test_third_party_pins_cpu.py).  Allowed importers: tests/, __graft_entry__.smoke(), bench.py.
"""
from collections import OrderedDict
import numpy as np
import torch
import torch.nn.functional as F

from . import confignet_oracle as O

BN_EPS = 1.001e-5        # keras-applications ResNet50 BatchNormalization epsilon [TF-2.1]
RESNET50_STAGES = [(64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)]
VGGFACE_MEAN = (93.5940, 104.7624, 129.1863)       # perceptual_loss.py:55
VGG16_LAYERS = [("conv", "block1_conv1"), ("conv", "block1_conv2"), ("pool", "block1_pool"),
                ("conv", "block2_conv1"), ("conv", "block2_conv2"), ("pool", "block2_pool"),
                ("conv", "block3_conv1"), ("conv", "block3_conv2"), ("conv", "block3_conv3"), ("pool", "block3_pool"),
                ("conv", "block4_conv1"), ("conv", "block4_conv2")]
VGG16_USED = [1, 2, 8, 12]     # perceptual_loss.py:35


def conv_valid_padded(x, kernel, bias, stride, pad):
    """ZeroPadding2D(pad) + Conv2D(padding='valid') on channels-last (ResNet50 stem, [TF-2.1] item 10)."""
    xc = F.pad(x.permute(0, 3, 1, 2), [pad, pad, pad, pad])
    y = F.conv2d(xc, kernel.permute(3, 2, 0, 1), bias, stride=stride)
    return y.permute(0, 2, 3, 1)


def batchnorm_inference(x, p, name):
    """keras BatchNormalization with its moving statistics (training=False inside the manual tape, section 3.4)."""
    g, b = p[name + "/gamma"], p[name + "/beta"]
    m, v = p[name + "/moving_mean"], p[name + "/moving_variance"]
    return (x - m) / torch.sqrt(v + BN_EPS) * g + b


def maxpool_3x3_s2_pad1(x):
    """ZeroPadding2D(1) + MaxPooling2D(3, strides=2): the padded zeros take part in the maximum."""
    xc = F.pad(x.permute(0, 3, 1, 2), [1, 1, 1, 1], value=0.0)
    return F.max_pool2d(xc, 3, 2).permute(0, 2, 3, 1)


def resnet50_forward(p, x, prefix="resnet/"):
    """keras-applications ResNet50 v1, include_top=False, pooling='avg' ([TF-2.1] item 10): stride on the first 1x1
    of each stage's first block, projection shortcut on the first block of every stage."""
    def conv(t, name, stride=1):
        return O.conv_same(t, p[prefix + name + "/kernel"], p[prefix + name + "/bias"], stride)

    x = conv_valid_padded(x, p[prefix + "conv1_conv/kernel"], p[prefix + "conv1_conv/bias"], 2, 3)
    x = torch.relu(batchnorm_inference(x, p, prefix + "conv1_bn"))
    x = maxpool_3x3_s2_pad1(x)
    for si, (f, blocks, stride) in enumerate(RESNET50_STAGES, start=2):
        for b in range(1, blocks + 1):
            q = "conv%d_block%d" % (si, b)
            s = stride if b == 1 else 1
            if b == 1:
                shortcut = batchnorm_inference(conv(x, q + "_0_conv", s), p, prefix + q + "_0_bn")
            else:
                shortcut = x
            y = torch.relu(batchnorm_inference(conv(x, q + "_1_conv", s), p, prefix + q + "_1_bn"))
            y = torch.relu(batchnorm_inference(conv(y, q + "_2_conv"), p, prefix + q + "_2_bn"))
            y = batchnorm_inference(conv(y, q + "_3_conv"), p, prefix + q + "_3_bn")
            x = torch.relu(shortcut + y)
    return x.mean(dim=(1, 2))


def rotation_range_multiplier(rotation_ranges=((-30, 30), (-10, 10), (0, 0))):
    """real_encoder.py:20-21."""
    return np.pi * np.array([rotation_ranges[0][1], rotation_ranges[1][1], rotation_ranges[2][1]], np.float64) / 180.0


def real_encoder_forward(p, img, rotation_ranges=((-30, 30), (-10, 10), (0, 0))):
    """RealEncoder.call real_encoder.py:23-34."""
    x = O.vgg19_preprocess(img)                      # resnet50.preprocess_input is the same 'caffe' mode
    feat = resnet50_forward(p, x)
    raw = torch.tanh(feat @ p["rotation_regressor/kernel"] + p["rotation_regressor/bias"])
    rot = raw * torch.as_tensor(rotation_range_multiplier(rotation_ranges)).to(img.dtype)
    emb = feat @ p["feature_to_latent_mlp/kernel"] + p["feature_to_latent_mlp/bias"]
    return emb, rot


def vggface_preprocess(img):
    """perceptual_loss.py:51-56: (x+1)*127.5 - face means, no channel flip."""
    return (img + 1) * 127.5 - torch.tensor(VGGFACE_MEAN, dtype=img.dtype)


def vgg16_activations(p, x):
    acts = []
    for idx, (kind, name) in enumerate(VGG16_LAYERS, start=1):
        if kind == "conv":
            x = torch.relu(O.conv_same(x, p[name + "/kernel"], p[name + "/bias"]))
        else:
            x = F.max_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
        if idx in VGG16_USED:
            acts.append(x)
    return acts


def face_reco_loss(p_vgg16, predicted, data):
    """PerceptualLoss(model_type='VGGFace').loss, reduced by face_reco_loss (confignet_second_stage.py:88-91)."""
    a_p = vgg16_activations(p_vgg16, vggface_preprocess(predicted))
    a_d = vgg16_activations(p_vgg16, vggface_preprocess(data))
    total = 0
    for x, y in zip(a_p, a_d):
        total = total + ((x.reshape(-1) - y.reshape(-1)) ** 2).mean()
    return total


def normalized_latent_regression_loss(p_lr, generator_outputs, labels, weight=10.0):
    """compute_normalized_latent_regression_loss confignet_second_stage.py:93-107."""
    out = O.latent_regressor_forward(p_lr, generator_outputs)
    return normalized_regression(out, labels, weight)


def normalized_regression(out, labels, weight):
    den = torch.sqrt(labels.var(dim=0, unbiased=False, keepdim=True) + 1e-3)
    den = torch.cat((den[:, :-3], torch.ones((1, 3), dtype=out.dtype)), dim=1)
    out_n = out.mean(dim=0) + (out - out.mean(dim=0)) / den
    lab_n = labels.mean(dim=0) + (labels - labels.mean(dim=0)) / den
    return ((lab_n - out_n) ** 2).mean(dim=-1).mean() * weight


def stage2_discriminator_step_losses(p_d, p_g, p_enc, real_imgs, input_imgs, output_res=256):
    """discriminator_training_step of ConfigNet: confignet_first_stage.py:466-470 over the overridden
    get_discriminator_batch (confignet_second_stage.py:119-130) - the fakes are reconstructions of encoded training
    images, generator(encode_images(input_imgs)), built outside the tape."""
    with torch.no_grad():
        latent, rotation = real_encoder_forward(p_enc, input_imgs)
        fake = O.generator_forward(p_g, latent, rotation, output_res)
    return O.compute_discriminator_loss(p_d, real_imgs, fake)


def stage2_latent_discriminator_step_losses(p_ld, p_enc, p_se, facemodel_inputs, real_imgs, facemodel_params):
    """latent_discriminator_training_step confignet_second_stage.py:132-147: real latents come from the encoder."""
    with torch.no_grad():
        real_latents, _ = real_encoder_forward(p_enc, real_imgs)
        fake = O.synthetic_encoder_forward(p_se, facemodel_params, facemodel_inputs)
    return O.compute_latent_discriminator_loss(p_ld, real_latents, fake)


def stage2_generator_step_losses(p_g, p_lr, p_se, p_enc, p_d, p_sd, p_ld, p_vgg, facemodel_inputs, batch, weights=None,
                                 output_res=256):
    """generator_training_step (stage 2) confignet_second_stage.py:149-211.
    batch: facemodel_params, synth_rotations, synth_imgs, eye_masks, real_imgs."""
    w = dict(O.DEFAULT_LOSS_WEIGHTS); w.update(weights or {})
    losses = OrderedDict()
    synth_latents = O.synthetic_encoder_forward(p_se, batch["facemodel_params"], facemodel_inputs)
    out_synth = O.generator_forward(p_g, synth_latents, batch["synth_rotations"], output_res)
    real_latents, real_rot = real_encoder_forward(p_enc, batch["real_imgs"])
    out_real = O.generator_forward(p_g, real_latents, real_rot, output_res)
    losses["image_loss_synth"] = w["image_loss_weight"] * O.perceptual_loss(p_vgg, batch["synth_imgs"], out_synth)
    losses["image_loss_real"] = w["image_loss_weight"] * O.perceptual_loss(p_vgg, batch["real_imgs"], out_real)
    losses["eye_loss"] = w["eye_loss_weight"] * O.eye_loss(batch["synth_imgs"], out_synth, batch["eye_masks"])
    for i, o in enumerate(O.discriminator_forward(p_sd, out_synth).values()):
        losses["GAN_loss_synth_%d" % i] = O.gan_g_loss(o)
    for i, o in enumerate(O.discriminator_forward(p_d, out_real).values()):
        losses["GAN_loss_real_%d" % i] = O.gan_g_loss(o)
    ld_synth = O.latent_discriminator_forward(p_ld, synth_latents)
    ld_real = O.latent_discriminator_forward(p_ld, real_latents)
    n_real, n_synth = ld_real.shape[0], ld_synth.shape[0]
    labels_da = torch.cat((torch.zeros(n_real, 1), torch.ones(n_synth, 1)), dim=0).to(ld_real.dtype)
    losses["latent_GAN_loss"] = w["domain_adverserial_loss_weight"] * O.gan_d_loss(labels_da, torch.cat((ld_real, ld_synth), dim=0))
    if w["latent_regression_weight"] > 0.0:
        stacked_latents = torch.cat((synth_latents, real_latents), dim=0)
        stacked_imgs = torch.cat((out_synth, out_real), dim=0)
        stacked_rot = torch.cat((batch["synth_rotations"], real_rot), dim=0)
        labels = torch.cat((stacked_latents, w["latent_regressor_rot_weight"] * stacked_rot), dim=-1)
        losses["latent_regression_loss"] = normalized_latent_regression_loss(p_lr, stacked_imgs, labels, w["latent_regression_weight"])
    losses["loss_sum"] = sum(losses.values())
    return losses


def fine_tune_losses(p_gft, p_lr, p_d, p_ld, p_vgg, p_vgg16, input_images, pre_expr, expr, post_expr, rotations,
                     weights=None, output_res=256):
    """One iteration of fine_tune_on_img's loss (confignet_second_stage.py:360-390).  pre/expr/post embeddings and
    rotations are the optimised variables; pre/post are (1, k) and tiled over the images."""
    w = dict(O.DEFAULT_LOSS_WEIGHTS); w.update(weights or {})
    n = input_images.shape[0]
    emb = torch.cat((pre_expr.expand(n, -1), expr, post_expr.expand(n, -1)), dim=1)
    out = O.generator_forward(p_gft, emb, rotations, output_res)
    losses = OrderedDict()
    losses["image_loss_real"] = 0.5 * w["image_loss_weight"] * O.perceptual_loss(p_vgg, input_images, out)
    losses["face_reco_loss"] = 0.5 * w["image_loss_weight"] * face_reco_loss(p_vgg16, out, input_images)
    for i, o in enumerate(O.discriminator_forward(p_d, out).values()):
        losses["GAN_loss_real_%d" % i] = O.gan_g_loss(o)
    ld = O.latent_discriminator_forward(p_ld, emb)
    losses["latent_GAN_loss"] = w["domain_adverserial_loss_weight"] * O.gan_d_loss(torch.ones(1, 1, dtype=ld.dtype), ld)
    labels = torch.cat((emb, w["latent_regressor_rot_weight"] * rotations), dim=-1)
    losses["latent_regression_loss"] = normalized_latent_regression_loss(p_lr, out, labels, w["latent_regression_weight"])
    losses["loss_sum"] = sum(losses.values())
    return losses


# ----------------------------------------------------------------------------- LatentGAN (latent_gan.py)
def latent_gan_mlp(p, x, num_layers=3):
    """MLPSimple(non_linear=LeakyReLU (alpha 0.3 [TF-2.1]), non_linear_last=None) building_blocks.py:152-173."""
    return O.mlp_simple(x, p, "mlp", num_layers, 0.3)


def latent_gan_discriminator_losses(p_d, p_g, real_embeddings, input_latents, num_layers=3):
    """discriminator_training_step latent_gan.py:117-149."""
    with torch.no_grad():
        fake = latent_gan_mlp(p_g, input_latents, num_layers)
    real = real_embeddings.detach().requires_grad_(True)
    o_real = latent_gan_mlp(p_d, real, num_layers)
    o_fake = latent_gan_mlp(p_d, fake, num_layers)
    losses = OrderedDict()
    losses["GAN_loss_real"] = O.gan_d_loss(torch.ones_like(o_real), o_real)
    losses["GAN_loss_fake"] = O.gan_d_loss(torch.zeros_like(o_fake), o_fake)
    losses["gp_loss"] = O.gradient_regularization(o_real, real)
    losses["loss_sum"] = sum(losses.values())
    return losses


def latent_gan_generator_losses(p_d, p_g, input_latents, num_layers=3):
    """generator_training_step latent_gan.py:151-165."""
    losses = OrderedDict()
    losses["gan_loss"] = O.gan_g_loss(latent_gan_mlp(p_d, latent_gan_mlp(p_g, input_latents, num_layers), num_layers))
    losses["loss_sum"] = sum(losses.values())
    return losses


# ----------------------------------------------------------------------------- whole iterations (CPU baselines of bench.py)
class OracleSecondStage(O.OracleFirstStage):
    """CPU restatement of the second-stage step loop (confignet_second_stage.py:119-218,268-299) on the first-stage
    networks plus the RealEncoder; same step functions as the parity tests, torch-CPU fp32: the timed CPU baseline of
    bench.py --config 4 and (fine_tune) --config 5."""

    def __init__(self, facemodel_inputs, output_res=256, seed=1234, dtype=torch.float32, image_loss_weight=0.0005):
        super().__init__(facemodel_inputs, output_res, seed, dtype)
        from confignet_b200 import netspec          # parameter tables only (pure NumPy, no CUDA)
        enc = netspec.init_real_encoder_params(145, seed + 8)
        self.p_enc = OrderedDict((k, torch.tensor(v, dtype=dtype, requires_grad=netspec.is_trainable(k))) for k, v in enc.items())
        self.p_vgg16 = O.to_torch(netspec.init_params(netspec.vgg16_spec(), seed + 9, vgg_like=True), dtype=dtype)
        self.weights = dict(O.DEFAULT_LOSS_WEIGHTS); self.weights["image_loss_weight"] = image_loss_weight

    def discriminator_step(self, real_u8, input_u8):
        losses = stage2_discriminator_step_losses(self.p_d, self.p_g, self.p_enc, self._t(real_u8) / 127.5 - 1.0,
                                                  self._t(input_u8) / 127.5 - 1.0, self.res)
        self.d_opt.apply_gradients(zip(O.grads_of(losses["loss_sum"], self.p_d), self.p_d.values()))
        return losses

    def latent_discriminator_step(self, real_u8, facemodel_params):
        losses = stage2_latent_discriminator_step_losses(self.p_ld, self.p_enc, self.p_se, self.fm, self._t(real_u8) / 127.5 - 1.0,
                                                         [self._t(a) for a in facemodel_params])
        self.d_opt.apply_gradients(zip(O.grads_of(losses["loss_sum"], self.p_ld), self.p_ld.values()))
        return losses

    def generator_step(self, facemodel_params, synth_rot, synth_u8, eye_masks, real_u8):
        batch = dict(facemodel_params=[self._t(a) for a in facemodel_params], synth_rotations=self._t(synth_rot),
                     synth_imgs=self._t(synth_u8) / 127.5 - 1.0, eye_masks=eye_masks, real_imgs=self._t(real_u8) / 127.5 - 1.0)
        losses = stage2_generator_step_losses(self.p_g, self.p_lr, self.p_se, self.p_enc, self.p_d, self.p_sd, self.p_ld, self.p_vgg,
                                              self.fm, batch, weights=self.weights, output_res=self.res)
        allp = OrderedDict()
        for pre, p in (("g/", self.p_g), ("lr/", self.p_lr), ("se/", self.p_se), ("enc/", self.p_enc)):
            for k, v in p.items():
                if v.requires_grad:
                    allp[pre + k] = v
        self.g_opt.apply_gradients(zip(O.grads_of(losses["loss_sum"], allp), allp.values()))
        O.update_smoothed_weights(self.p_gs, self.p_g)
        return losses

    def fine_tune(self, imgs_u8, n_iters):
        """fine_tune_on_img confignet_second_stage.py:321-403: encoder prediction, shared pre/post parts from the mean
        embedding, per-image expression part and rotations, Adam(1e-4), n_iters iterations."""
        imgs = self._t(imgs_u8) / 127.5 - 1.0
        with torch.no_grad():
            e0, r0 = real_encoder_forward(self.p_enc, imgs)
        names = list(self.fm.keys())
        lo = sum(self.fm[n][1] for n in names[:names.index("blendshape_values")])
        hi = lo + self.fm["blendshape_values"][1]
        mean_e = e0.mean(dim=0, keepdim=True)
        pre, expr, post, rots = [t.clone().requires_grad_(True) for t in (mean_e[:, :lo], e0[:, lo:hi], mean_e[:, hi:], r0)]
        opt = O.KerasAdam(lr=1e-4, beta_1=0.9, beta_2=0.999)
        p_ft = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in self.p_gs.items())
        for _ in range(n_iters):
            l = fine_tune_losses(p_ft, self.p_lr, self.p_d, self.p_ld, self.p_vgg, self.p_vgg16, imgs, pre, expr, post, rots,
                                 weights=self.weights, output_res=self.res)
            tv = list(p_ft.values()) + [pre, post, rots, expr]
            gs = torch.autograd.grad(l["loss_sum"], tv, allow_unused=True)
            opt.apply_gradients(zip([torch.zeros_like(v) if q is None else q for q, v in zip(gs, tv)], tv))
        return l


class OracleLatentGAN:
    """latent_gan.py:117-174,232-247 on CPU: one discriminator step, one generator step (one optimizer), EMA."""

    def __init__(self, latent_dim=145, seed=4321, dtype=torch.float32):
        from confignet_b200 import netspec
        mk = lambda spec, s: O.to_torch(netspec.init_params(spec, s), dtype=dtype, requires_grad=True)
        self.p_g = mk(netspec.latent_gan_mlp_spec(latent_dim), seed)
        self.p_d = mk(netspec.latent_gan_mlp_spec(latent_dim, num_out=1), seed + 1)
        self.p_gs = O.to_torch({k: v.detach().numpy() for k, v in self.p_g.items()}, dtype=dtype)
        self.opt = O.KerasAdam(lr=5e-5)
        self.dtype = dtype

    def step(self, real_embeddings, z_d, z_g):
        t = lambda a: torch.as_tensor(np.asarray(a)).to(self.dtype)
        ld = latent_gan_discriminator_losses(self.p_d, self.p_g, t(real_embeddings), t(z_d))
        self.opt.apply_gradients(zip(O.grads_of(ld["loss_sum"], self.p_d), self.p_d.values()))
        lg = latent_gan_generator_losses(self.p_d, self.p_g, t(z_g))
        self.opt.apply_gradients(zip(O.grads_of(lg["loss_sum"], self.p_g), self.p_g.values()))
        O.update_smoothed_weights(self.p_gs, self.p_g)
        return ld, lg
