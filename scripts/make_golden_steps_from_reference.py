"""Generates tests/golden/reference_steps.npz by EXECUTING the reference's own training-step methods -
ConfigNetFirstStage.discriminator_training_step, synth_discriminator_training_step,
latent_discriminator_training_step, generator_training_step (confignet/confignet_first_stage.py:438-560) incl. their
NumPy batch assembly, nested GradientTapes and optimizer.apply_gradients - from /root/reference on the torch-backed
TensorFlow stand-in (scripts/tf_torch_shim.py, float64) with a Keras-Adam stand-in that states the [TF-2.1] update.
Networks are the reference's own classes with the seeded oracle parameters copied in; batch size 2 at 256x256.

The one thing the reference cannot run here is its perceptual loss (keras.applications VGG19 + ImageNet weights): the
G step gets a stand-in ``perceptual_loss.loss(gt, gen) = 1e4 * mean((gt - gen)^2)`` and the oracle is given the same
stand-in - every other term, weight, label and the choice of which network sees which images is the reference's.

For every step the script replays the NumPy draws in the reference's order, feeds the oracle's step functions
(oracle/confignet_oracle.py) and its Keras-Adam, and requires equal loss dictionaries and equal updated weights.
tests/test_host_cpu.py::test_oracle_steps_match_reference_steps repeats the oracle half against the committed vectors.

    python scripts/make_golden_steps_from_reference.py
"""
import importlib
import math
import os
import sys
import types
from collections import OrderedDict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "reference_steps.npz")

import tf_torch_shim as S                                         # noqa: E402
tf = S.install()


# ---- additions the step methods need: trainable_weights, a Keras-Adam stand-in
def _collect(obj, seen, out):
    if isinstance(obj, S.Layer):
        if id(obj) in seen:
            return
        seen.add(id(obj))
        for _, w in obj.weights:
            out.append(w)
        for v in vars(obj).values():
            _collect(v, seen, out)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            _collect(v, seen, out)
    elif isinstance(obj, dict):
        for v in obj.values():
            _collect(v, seen, out)


S.Layer.trainable_weights = property(lambda self: (lambda o: (_collect(self, set(), o), o)[1])([]))


class Adam:
    """[TF-2.1] keras.optimizers.Adam, amsgrad=False: t = iterations + 1; lr_t = lr sqrt(1 - b2^t) / (1 - b1^t);
    m <- b1 m + (1 - b1) g; v <- b2 v + (1 - b2) g^2; theta <- theta - lr_t m / (sqrt(v) + eps), eps = 1e-7."""

    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, amsgrad=False, **kw):
        assert not amsgrad
        self.lr, self.b1, self.b2, self.eps, self.iterations, self.slots = lr, beta_1, beta_2, epsilon, 0, {}

    def apply_gradients(self, grads_and_vars):
        t = self.iterations + 1
        lr_t = self.lr * math.sqrt(1 - self.b2 ** t) / (1 - self.b1 ** t)
        with torch.no_grad():
            for g, v in grads_and_vars:
                if g is None:
                    continue
                m, s = self.slots.setdefault(id(v), (torch.zeros_like(v), torch.zeros_like(v)))
                m.mul_(self.b1).add_((1 - self.b1) * g)
                s.mul_(self.b2).add_((1 - self.b2) * g * g)
                v.sub_(lr_t * m / (s.sqrt() + self.eps))
        self.iterations += 1


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
        return sys.modules.get(self.__name__ + "." + name, _Anything)


for m in ["matplotlib", "matplotlib.pyplot", "transformations", "azureml", "azureml.core", "azureml.core.run",
          "tensorflow.keras.utils", "tensorflow.keras.applications"]:
    parts = m.split(".")
    for i in range(1, len(parts) + 1):
        sys.modules.setdefault(".".join(parts[:i]), _Stub(".".join(parts[:i])))
sys.path.insert(0, REF)
for name, sub in (("confignet", "confignet"), ("confignet.dnn_models", "confignet/dnn_models")):
    pkg = types.ModuleType(name)
    pkg.__path__ = [os.path.join(REF, sub)]
    sys.modules[name] = pkg
for heavy in ["confignet.perceptual_loss", "confignet.metrics", "confignet.metrics.metrics"]:
    sys.modules[heavy] = _Stub(heavy)
sys.modules["confignet"].confignet_utils = importlib.import_module("confignet.confignet_utils")
gen_mod = importlib.import_module("confignet.dnn_models.hologan_generator")
dis_mod = importlib.import_module("confignet.dnn_models.hologan_discriminator")
enc_mod = importlib.import_module("confignet.dnn_models.synthetic_encoder")
blocks = importlib.import_module("confignet.dnn_models.building_blocks")
fs = importlib.import_module("confignet.confignet_first_stage")

from confignet_b200 import netspec                               # noqa: E402
from oracle import confignet_oracle as O                         # noqa: E402

FM = netspec.default_facemodel_inputs()
SEEDS = dict(g=201, d=202, sd=203, ld=204, lr=205, se=206)
B, RES, N_IMGS = 2, 256, 5


def seeded_params(spec, seed):
    arrays = netspec.perturb_params(netspec.init_params(spec, seed), seed + 1000, 0.05)
    return O.to_torch(arrays, dtype=torch.float64, requires_grad=True)


def put(layer, p, prefix, names=("kernel", "bias")):
    for n in names:
        w, v = getattr(layer, n), p[prefix + "/" + n]
        assert tuple(w.shape) == tuple(v.shape), (prefix, n)
        with torch.no_grad():
            w.copy_(v)


def dense_layers(mlp_simple):
    return [l for l in mlp_simple.map.layers if type(l).__name__ == "Dense"]


def make_dataset(seed):
    r = np.random.RandomState(seed)
    ds = types.SimpleNamespace()
    ds.imgs = r.randint(0, 256, (N_IMGS, RES, RES, 3)).astype(np.uint8)
    ds.eye_masks = (r.rand(N_IMGS, RES, RES) < 0.01).astype(np.uint8)
    ds.metadata_inputs = {k: r.rand(N_IMGS, d[0]).astype(np.float32) for k, d in FM.items()}
    rot = np.zeros((N_IMGS, 3), np.float32)
    rot[:, 0] = r.uniform(-0.5, 0.5, N_IMGS); rot[:, 1] = r.uniform(-0.17, 0.17, N_IMGS)
    ds.metadata_inputs["rotations"] = rot
    return ds


real_set, synth_set = make_dataset(31), make_dataset(32)

# ------------------------------------------------------------------------------------------------ the reference model object
cfg = {"output_shape": (RES, RES, 3), "batch_size": B, "facemodel_inputs": {k: tuple(v) for k, v in FM.items()}}
model = fs.ConfigNetFirstStage(cfg, initialize=False)
assert list(model.config["facemodel_inputs"].keys()) == list(FM.keys())
P = dict(g=seeded_params(netspec.generator_spec(145, RES), SEEDS["g"]), d=seeded_params(netspec.discriminator_spec(RES), SEEDS["d"]),
         sd=seeded_params(netspec.discriminator_spec(RES), SEEDS["sd"]), ld=seeded_params(netspec.latent_discriminator_spec(145, 4), SEEDS["ld"]),
         lr=seeded_params(netspec.latent_regressor_spec(145, RES), SEEDS["lr"]), se=seeded_params(netspec.synthetic_encoder_spec(FM, 2), SEEDS["se"]))
disc_args = dict(img_shape=(RES, RES), num_resample=5, disc_kernel_size=3, disc_expansion_factor=48,
                 disc_max_feature_maps=512, initial_from_rgb_layer_in_discr=True)
model.generator = gen_mod.HologanGenerator(**model._get_generator_kwargs())
model.discriminator = dis_mod.HologanDiscriminator(**disc_args)
model.synth_discriminator = dis_mod.HologanDiscriminator(**disc_args)
model.latent_regressor = dis_mod.HologanLatentRegressor(model.config["latent_dim"], **disc_args)
model.latent_discriminator = blocks.MLPSimple(num_layers=model.config["n_latent_discr_layers"], num_in=145, num_hidden=145,
                                              num_out=1, non_linear=S.LeakyReLU, non_linear_last=None)
model.synthetic_encoder = enc_mod.SyntheticDataEncoder(synthetic_encoder_inputs=model.config["facemodel_inputs"],
                                                       num_layers=model.config["num_synth_encoder_layers"])
PERC_SCALE = 1e4
model.perceptual_loss = types.SimpleNamespace(loss=lambda gt, gen: PERC_SCALE * ((S.T(gt) - S.T(gen)) ** 2).mean())
O.perceptual_loss = lambda p_vgg, gt, gen: PERC_SCALE * ((gt - gen) ** 2).mean()          # the same stand-in for the oracle

with torch.no_grad():                                             # build every layer
    z0, r0 = torch.zeros(1, 145, dtype=torch.float64), torch.zeros(1, 3, dtype=torch.float64)
    img0 = model.generator([z0, r0])
    model.discriminator(img0); model.synth_discriminator(img0); model.latent_regressor(img0); model.latent_discriminator(z0)
    model.synthetic_encoder([torch.zeros(1, d[0], dtype=torch.float64) for d in FM.values()])


def load_generator(G, p):
    put(G.learned_input_layer, p, "learned_input")
    for blk, attr in (("map_3d_0", "map_3d"), ("map_3d_1", "map_3d"), ("map_2d_0", "map_2d"), ("map_2d_1", "map_2d"),
                      ("map_2d_2", "map_2d"), ("map_2d_2b", "map_2d")):
        b = getattr(G, blk)
        put(getattr(b, attr).layers[0], p, blk + "/conv")
        d0, d1 = dense_layers(b.adain.adain_mlp)
        put(d0, p, blk + "/adain/dense0"); put(d1, p, blk + "/adain/dense1")
    put(G.map_3d_post.layers[0], p, "map_3d_post/conv0"); put(G.map_3d_post.layers[2], p, "map_3d_post/conv1")
    put(G.projection_conv, p, "projection_conv"); put(G.map_final, p, "map_final")


def load_trunk(m, p):
    put(m.initial_1x1_conv, p, "initial_1x1_conv")
    for i, blk in enumerate(m.conv_blocks):
        put(blk.map_2d, p, "block%d/conv" % i)
        put(blk.instance_norm, p, "block%d/in" % i, names=("gamma", "beta"))


def load_disc(m, p):
    load_trunk(m, p)
    for i, sc in enumerate(m.style_classifiers):
        put(sc, p, "style%d" % i)
    put(m.disc_map, p, "disc_map")


load_generator(model.generator, P["g"])
load_disc(model.discriminator, P["d"]); load_disc(model.synth_discriminator, P["sd"])
load_trunk(model.latent_regressor, P["lr"]); put(model.latent_regressor.latent_predictor, P["lr"], "latent_predictor")
for j, d in enumerate(dense_layers(model.latent_discriminator)):
    put(d, P["ld"], "mlp/dense%d" % j)
for name in FM:
    d0, d1 = dense_layers(model.synthetic_encoder.per_facemodel_input_mlps[name])
    put(d0, P["se"], "mlp_%s/dense0" % name); put(d1, P["se"], "mlp_%s/dense1" % name)


def ref_weight(net, name):
    """the reference layer tensor that carries oracle parameter `name` of network `net` (spot checks after the update)"""
    m = {"d": model.discriminator, "sd": model.synth_discriminator, "lr": model.latent_regressor, "g": model.generator}[net]
    if net == "g":
        return {"map_3d_1/conv/kernel": m.map_3d_1.map_3d.layers[0].kernel, "map_final/kernel": m.map_final.kernel,
                "map_2d_1/adain/dense1/bias": dense_layers(m.map_2d_1.adain.adain_mlp)[1].bias,
                "learned_input/bias": m.learned_input_layer.bias}[name]
    return {"block0/conv/kernel": m.conv_blocks[0].map_2d.kernel, "block3/in/gamma": m.conv_blocks[3].instance_norm.gamma,
            "style2/kernel": getattr(m, "style_classifiers", [None] * 3)[2].kernel if hasattr(m, "style_classifiers") else None,
            "latent_predictor/bias": getattr(m, "latent_predictor", types.SimpleNamespace(bias=None)).bias}[name]


d_opt_ref, g_opt_ref = Adam(**model.config["optimizer"]), Adam(**model.config["optimizer"])
d_opt_orc, g_opt_orc = O.KerasAdam(), O.KerasAdam()
assert (d_opt_ref.lr, d_opt_ref.b1, d_opt_ref.b2) == (d_opt_orc.lr, d_opt_orc.b1, d_opt_orc.b2)
out = {}


def check(tag, l_ref, l_orc, pairs):
    assert list(l_ref.keys()) == list(l_orc.keys()), (tag, list(l_ref.keys()), list(l_orc.keys()))
    err = max(abs(float(l_ref[k]) - float(l_orc[k])) / max(1.0, abs(float(l_ref[k]))) for k in l_ref)
    werr = max(float((a.detach() - b.detach()).abs().max()) for a, b in pairs)
    print("%-8s %2d loss terms, max rel diff %.2e; updated weights max abs diff %.2e; loss_sum %.6f" %
          (tag, len(l_ref), err, werr, float(l_ref["loss_sum"])))
    assert err < 1e-9 and werr < 1e-10
    out[tag + "_keys"] = np.array(list(l_ref.keys()))
    out[tag + "_vals"] = np.array([float(v) for v in l_ref.values()])
    for i, (a, _) in enumerate(pairs):
        out["%s_w%d" % (tag, i)] = sub(a)


def sub(t):
    """<= 2048 evenly strided elements of a tensor (the fixture stays small)"""
    v = t.detach().numpy().ravel()
    return v[::max(1, -(-v.size // 2048))].copy()


def T64(a):
    return torch.as_tensor(np.asarray(a)).to(torch.float64)


def draw_flipped_real(ds):
    """get_discriminator_batch's first draws (confignet_first_stage.py:440-443, confignet_utils.py:198-204)"""
    idx = np.random.randint(0, ds.imgs.shape[0], B)
    real = np.copy(ds.imgs[idx]).astype(np.float32) / 127.5 - 1.0
    flips = np.random.randint(0, 2, size=B)
    for i in range(B):
        if flips[i]:
            real[i] = real[i][:, ::-1]
    return real


def draw_synth(ds, n):
    """sample_synthetic_dataset (confignet_first_stage.py:425-435)"""
    idx = np.random.randint(0, ds.imgs.shape[0], n)
    return ([ds.metadata_inputs[k][idx] for k in FM], ds.metadata_inputs["rotations"][idx].astype(np.float32),
            np.copy(ds.imgs[idx]).astype(np.float32), np.copy(ds.eye_masks[idx]))


# ------------------------------------------------------------------------------------------------ D step
np.random.seed(41)
l_ref = model.discriminator_training_step(real_set, d_opt_ref)
np.random.seed(41)
real = draw_flipped_real(real_set)
lat = np.random.normal(0, 1, (B, 145))
rot = model.sample_rotations(B)
l_orc = O.discriminator_step_losses(P["d"], P["g"], T64(real), T64(lat), T64(rot), RES)
d_opt_orc.apply_gradients(zip(O.grads_of(l_orc["loss_sum"], P["d"]), P["d"].values()))
check("d", l_ref, l_orc, [(ref_weight("d", n), P["d"][n]) for n in ("block0/conv/kernel", "block3/in/gamma", "style2/kernel")])

# ------------------------------------------------------------------------------------------------ synth-D step (same optimizer)
np.random.seed(42)
l_ref = model.synth_discriminator_training_step(synth_set, d_opt_ref)
np.random.seed(42)
real = draw_flipped_real(synth_set)
fm_p, srot, _, _ = draw_synth(synth_set, B)
l_orc = O.synth_discriminator_step_losses(P["sd"], P["g"], P["se"], FM, T64(real), [T64(a) for a in fm_p], T64(srot), RES)
d_opt_orc.apply_gradients(zip(O.grads_of(l_orc["loss_sum"], P["sd"]), P["sd"].values()))
check("synth_d", l_ref, l_orc, [(ref_weight("sd", n), P["sd"][n]) for n in ("block0/conv/kernel", "block3/in/gamma", "style2/kernel")])

# ------------------------------------------------------------------------------------------------ latent-D step (same optimizer)
np.random.seed(43)
l_ref = model.latent_discriminator_training_step(synth_set, d_opt_ref)
np.random.seed(43)
real_lat = np.random.normal(0, 1, (B, 145))
fm_p, _, _, _ = draw_synth(synth_set, B)
l_orc = O.latent_discriminator_step_losses(P["ld"], P["se"], FM, T64(real_lat), [T64(a) for a in fm_p])
d_opt_orc.apply_gradients(zip(O.grads_of(l_orc["loss_sum"], P["ld"]), P["ld"].values()))
ld_layers = dense_layers(model.latent_discriminator)
check("latent_d", l_ref, l_orc, [(ld_layers[0].kernel, P["ld"]["mlp/dense0/kernel"]), (ld_layers[3].bias, P["ld"]["mlp/dense3/bias"])])
assert d_opt_ref.iterations == d_opt_orc.iterations == 3              # the three discriminators share one step counter

# ------------------------------------------------------------------------------------------------ G step
np.random.seed(44)
l_ref = model.generator_training_step(real_set, synth_set, g_opt_ref)
np.random.seed(44)
ns = B // 2
fm_p, srot, gt, masks = draw_synth(synth_set, ns)
gt = gt / 127.5 - 1.0
real_lat = np.random.normal(0, 1, (B - ns, 145))
real_rot = model.sample_rotations(B - ns)
batch = dict(facemodel_params=[T64(a) for a in fm_p], synth_rotations=T64(srot), gt_imgs=T64(gt), eye_masks=masks,
             real_latents=T64(real_lat), real_rotations=T64(real_rot))
l_orc = O.generator_step_losses(P["g"], P["lr"], P["se"], P["d"], P["sd"], P["ld"], None, FM, batch, output_res=RES)
allp = OrderedDict()
for pre, p in (("g/", P["g"]), ("lr/", P["lr"]), ("se/", P["se"])):
    for k, v in p.items():
        allp[pre + k] = v
g_opt_orc.apply_gradients(zip(O.grads_of(l_orc["loss_sum"], allp), allp.values()))
se_layers = dense_layers(model.synthetic_encoder.per_facemodel_input_mlps["blendshape_values"])
check("g", l_ref, l_orc, [(ref_weight("g", "map_3d_1/conv/kernel"), P["g"]["map_3d_1/conv/kernel"]),
                          (ref_weight("g", "map_final/kernel"), P["g"]["map_final/kernel"]),
                          (ref_weight("g", "map_2d_1/adain/dense1/bias"), P["g"]["map_2d_1/adain/dense1/bias"]),
                          (ref_weight("g", "learned_input/bias"), P["g"]["learned_input/bias"]),
                          (ref_weight("lr", "latent_predictor/bias"), P["lr"]["latent_predictor/bias"]),
                          (se_layers[1].kernel, P["se"]["mlp_blendshape_values/dense1/kernel"])])

# ------------------------------------------------------------------------------------------------ stage-2 steps (confignet_second_stage.py:132-218)
# ConfigNet's own latent-discriminator and generator steps on the SAME network objects (their weights now differ from
# the seeds by one update - the oracle's dictionaries moved in lockstep) and the same optimizers (second / fourth
# iteration: Keras' bias correction is exercised at t > 1).  RealEncoder is keras-applications ResNet50 + two Dense
# heads and cannot run here: both sides get the same small differentiable stand-in encoder
#   f = 2x2 average-pooled image (12 values); latent = f A; rotation = tanh(f Br) * (pi/180 * [30, 10, 0]).
sys.modules.setdefault("cv2", _Stub("cv2"))
sys.modules["confignet"].ConfigNetFirstStage = fs.ConfigNetFirstStage        # what confignet/__init__.py exports
second = importlib.import_module("confignet.confignet_second_stage")
from oracle import confignet_oracle_stage2 as O2                  # noqa: E402
m2 = second.ConfigNet(dict(cfg, image_loss_weight=5e-4), initialize=False)
for attr in ("generator", "discriminator", "synth_discriminator", "latent_regressor", "latent_discriminator",
             "synthetic_encoder", "perceptual_loss"):
    setattr(m2, attr, getattr(model, attr))
MULT = torch.tensor(np.pi * np.array([30.0, 10.0, 0.0]) / 180.0)


def enc_fn(A, Br, imgs):
    f = imgs.reshape(imgs.shape[0], 2, RES // 2, 2, RES // 2, 3).mean(dim=(2, 4)).reshape(imgs.shape[0], 12)
    return f @ A, torch.tanh(f @ Br) * MULT


class EncStandIn(S.Model):
    def __init__(self):
        S.Model.__init__(self)
        r = np.random.RandomState(61)
        self.A = self.add_weight(shape=(12, 145), name="A")
        self.Br = self.add_weight(shape=(12, 3), name="Br")
        with torch.no_grad():
            self.A.copy_(torch.tensor(r.randn(12, 145))); self.Br.copy_(torch.tensor(r.randn(12, 3)))

    def call(self, imgs):
        return enc_fn(self.A, self.Br, S.T(imgs))


m2.encoder = EncStandIn()
p_enc = OrderedDict((("A", m2.encoder.A.detach().clone().requires_grad_(True)), ("Br", m2.encoder.Br.detach().clone().requires_grad_(True))))
O2.real_encoder_forward = lambda p, imgs, *a, **k: enc_fn(p["A"], p["Br"], imgs)
W2 = dict(O.DEFAULT_LOSS_WEIGHTS)
for k in W2:
    W2[k] = m2.config[k]
assert W2["image_loss_weight"] == 5e-4


def draw_random_batch(ds, n):
    """sample_random_batch_of_images (confignet_second_stage.py:108-116)"""
    idx = np.random.randint(0, ds.imgs.shape[0], n)
    imgs = np.copy(ds.imgs[idx]).astype(np.float32) / 127.5 - 1.0
    flips = np.random.randint(0, 2, size=n)
    for i in range(n):
        if flips[i]:
            imgs[i] = imgs[i][:, ::-1]
    return imgs


np.random.seed(47)
l_ref = m2.latent_discriminator_training_step(real_set, synth_set, d_opt_ref)
np.random.seed(47)
rimgs = draw_random_batch(real_set, B)
fm_p, _, _, _ = draw_synth(synth_set, B)
l_orc = O2.stage2_latent_discriminator_step_losses(P["ld"], p_enc, P["se"], FM, T64(rimgs), [T64(a) for a in fm_p])
d_opt_orc.apply_gradients(zip(O.grads_of(l_orc["loss_sum"], P["ld"]), P["ld"].values()))
check("s2_latent_d", l_ref, l_orc, [(ld_layers[0].kernel, P["ld"]["mlp/dense0/kernel"]), (ld_layers[3].bias, P["ld"]["mlp/dense3/bias"])])

np.random.seed(48)
l_ref = m2.generator_training_step(real_set, synth_set, g_opt_ref)
np.random.seed(48)
fm_p, srot, simgs, masks = draw_synth(synth_set, B // 2)
simgs = simgs / 127.5 - 1.0
rimgs = draw_random_batch(real_set, B - B // 2)
batch = dict(facemodel_params=[T64(a) for a in fm_p], synth_rotations=T64(srot), synth_imgs=T64(simgs), eye_masks=masks,
             real_imgs=T64(rimgs))
l_orc = O2.stage2_generator_step_losses(P["g"], P["lr"], P["se"], p_enc, P["d"], P["sd"], P["ld"], None, FM, batch, weights=W2,
                                        output_res=RES)
allp = OrderedDict()
for pre, p in (("g/", P["g"]), ("lr/", P["lr"]), ("se/", P["se"]), ("enc/", p_enc)):
    for k, v in p.items():
        allp[pre + k] = v
g_opt_orc.apply_gradients(zip(O.grads_of(l_orc["loss_sum"], allp), allp.values()))
check("s2_g", l_ref, l_orc, [(ref_weight("g", "map_3d_1/conv/kernel"), P["g"]["map_3d_1/conv/kernel"]),
                             (ref_weight("g", "map_final/kernel"), P["g"]["map_final/kernel"]),
                             (ref_weight("lr", "latent_predictor/bias"), P["lr"]["latent_predictor/bias"]),
                             (se_layers[1].kernel, P["se"]["mlp_blendshape_values/dense1/kernel"]),
                             (m2.encoder.A, p_enc["A"]), (m2.encoder.Br, p_enc["Br"])])
assert g_opt_ref.iterations == g_opt_orc.iterations == 2 and d_opt_ref.iterations == d_opt_orc.iterations == 4

# the image-discriminator step of stage 2: the base class' discriminator_training_step (confignet_first_stage.py:466-476)
# calls get_discriminator_batch, which ConfigNet OVERRIDES (confignet_second_stage.py:119-130): fakes are
# generator(encode_images(training images)); NumPy stream: image rows, flips, input image rows
S.Model.predict = lambda self, x: tuple(v.detach().numpy() for v in self(x)) if isinstance(self(x), tuple) else self(x).detach().numpy()
np.random.seed(49)
l_ref = m2.discriminator_training_step(real_set, d_opt_ref)
np.random.seed(49)
rimgs = draw_random_batch(real_set, B)
in_idx = np.random.randint(0, real_set.imgs.shape[0], B)
in_imgs = real_set.imgs[in_idx].astype(np.float32) / 127.5 - 1.0
l_orc = O2.stage2_discriminator_step_losses(P["d"], P["g"], p_enc, T64(rimgs), T64(in_imgs), output_res=RES)
d_opt_orc.apply_gradients(zip(O.grads_of(l_orc["loss_sum"], P["d"]), P["d"].values()))
check("s2_d", l_ref, l_orc, [(ref_weight("d", n), P["d"][n]) for n in ("block0/conv/kernel", "block3/in/gamma", "style2/kernel")])
assert d_opt_ref.iterations == d_opt_orc.iterations == 5

# ------------------------------------------------------------------------------------------------ fine_tune_on_img (confignet_second_stage.py:321-403)
# Two iterations on two images: the embedding slicing (shared pre/post-expression parts from the MEAN embedding,
# per-image expression part), the loss terms, the list of trained variables and Adam(lr=1e-4, Keras defaults).  The two
# perceptual networks are asymmetric stand-ins here (so that swapped arguments would show):
#   perceptual_loss.loss(gt, gen) = 1e4 mean((gt - 0.7 gen)^2);  perceptual_loss_face_reco.loss(gen, gt) = 2e3 mean((gen - 0.5 gt)^2)
S.Model.predict = lambda self, x: tuple(v.detach().numpy() for v in self(x)) if isinstance(self(x), tuple) else self(x).detach().numpy()
S.Model.get_weights = lambda self: [w.detach().numpy().copy() for w in self.trainable_weights]


def _set_weights(self, ws):
    tw = self.trainable_weights
    assert len(tw) == len(ws)
    for w, v in zip(tw, ws):
        with torch.no_grad():
            w.copy_(torch.as_tensor(v))


S.Model.set_weights = _set_weights
tf.keras.optimizers = types.SimpleNamespace(Adam=Adam)
second.keras = tf.keras
m2.perceptual_loss = types.SimpleNamespace(loss=lambda gt, gen: 1e4 * ((S.T(gt) - 0.7 * S.T(gen)) ** 2).mean())
m2.perceptual_loss_face_reco = types.SimpleNamespace(loss=lambda gen, gt: 2e3 * ((S.T(gen) - 0.5 * S.T(gt)) ** 2).mean())
O.perceptual_loss = lambda p_vgg, gt, gen: 1e4 * ((gt - 0.7 * gen) ** 2).mean()
O2.face_reco_loss = lambda p_vgg16, gen, gt: 2e3 * ((gen - 0.5 * gt) ** 2).mean()
m2.generator_smoothed = gen_mod.HologanGenerator(**m2._get_generator_kwargs())
with torch.no_grad():
    m2.generator_smoothed([z0, r0])
p_gs = seeded_params(netspec.generator_spec(145, RES), 221)
load_generator(m2.generator_smoothed, p_gs)
ft_imgs = np.random.RandomState(62).randint(0, 256, (2, RES, RES, 3)).astype(np.uint8)
emb_ref, rot_ref = m2.fine_tune_on_img(ft_imgs, n_iters=2)
# the oracle's loop
imgs = T64(ft_imgs / 127.5 - 1.0)
with torch.no_grad():
    e0, r0_ = enc_fn(p_enc["A"], p_enc["Br"], imgs)
lo, hi = 7, 37                                                     # blendshape_values in the sorted latent layout
mean_e = e0.mean(dim=0, keepdim=True)
pre, expr, post, rots = [t.clone().requires_grad_(True) for t in (mean_e[:, :lo], e0[:, lo:hi], mean_e[:, hi:], r0_)]
opt = O.KerasAdam(lr=1e-4, beta_1=0.9, beta_2=0.999)
p_ft = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in p_gs.items())
for _ in range(2):
    # the reference returns the tiled pre / post parts built inside the LAST tape (before the last update) next to the
    # updated expression part (confignet_second_stage.py:361-364,402)
    pre_t, post_t = pre.detach().clone(), post.detach().clone()
    l_orc = O2.fine_tune_losses(p_ft, P["lr"], P["d"], P["ld"], None, None, imgs, pre, expr, post, rots, weights=W2, output_res=RES)
    tv = list(p_ft.values()) + [pre, post, rots, expr]
    gs = torch.autograd.grad(l_orc["loss_sum"], tv, allow_unused=True)
    opt.apply_gradients(zip([torch.zeros_like(v) if g_ is None else g_ for g_, v in zip(gs, tv)], tv))
emb_orc = torch.cat((pre_t.expand(2, -1), expr, post_t.expand(2, -1)), dim=1).detach().numpy()
e_emb, e_rot = np.abs(emb_ref - emb_orc).max(), np.abs(rot_ref - rots.detach().numpy()).max()
e_w = float((m2.generator_fine_tuned.map_final.kernel.detach() - p_ft["map_final/kernel"].detach()).abs().max())
print("fine_tune 2 iterations on 2 images: embeddings max abs diff %.2e, rotations %.2e, map_final kernel %.2e; last loss_sum %.6f"
      % (e_emb, e_rot, e_w, float(l_orc["loss_sum"])))
assert e_emb < 1e-10 and e_rot < 1e-10 and e_w < 1e-10
out["ft_emb"], out["ft_rot"] = emb_ref, rot_ref
out["ft_w_map_final"] = sub(m2.generator_fine_tuned.map_final.kernel)
out["ft_w_map_3d_0"] = sub(m2.generator_fine_tuned.map_3d_0.map_3d.layers[0].kernel)

# ------------------------------------------------------------------------------------------------ LatentGAN steps (latent_gan.py:117-165)
S.Model.predict = lambda self, x: self(x).detach().numpy()          # keras Model.predict: forward without a tape -> NumPy
S.Model.set_weights = lambda self, ws: [w.data.copy_(torch.as_tensor(v)) for w, v in zip(self.trainable_weights, ws)] and None
sys.modules["confignet.metrics"] = _Stub("confignet.metrics"); sys.modules["confignet.metrics.metrics"] = _Stub("confignet.metrics.metrics")
lg = importlib.import_module("confignet.latent_gan")
gan = lg.LatentGAN({"latent_dim": 145, "batch_size": 8})
p_lg = seeded_params(netspec.latent_gan_mlp_spec(145), 211)
p_ldg = seeded_params(netspec.latent_gan_mlp_spec(145, num_out=1), 212)
with torch.no_grad():
    gan.generator(torch.zeros(1, 145, dtype=torch.float64)); gan.discriminator(torch.zeros(1, 145, dtype=torch.float64))
for j, d in enumerate(dense_layers(gan.generator)):
    put(d, p_lg, "mlp/dense%d" % j)
for j, d in enumerate(dense_layers(gan.discriminator)):
    put(d, p_ldg, "mlp/dense%d" % j)
gt_emb = np.random.RandomState(51).randn(40, 145)
opt_ref_d, opt_ref_g = Adam(**gan.config["optimizer"]), Adam(**gan.config["optimizer"])
opt_orc_d, opt_orc_g = O.KerasAdam(lr=5e-5), O.KerasAdam(lr=5e-5)
np.random.seed(45)
l_ref = gan.discriminator_training_step(gt_emb, opt_ref_d)
np.random.seed(45)
zin = np.random.normal(0, 1, (8, 145))
idx = np.random.randint(0, gt_emb.shape[0], 8)
l_orc = O2.latent_gan_discriminator_losses(p_ldg, p_lg, T64(gt_emb[idx]), T64(zin))
opt_orc_d.apply_gradients(zip(O.grads_of(l_orc["loss_sum"], p_ldg), p_ldg.values()))
dl = dense_layers(gan.discriminator)
check("lgan_d", l_ref, l_orc, [(dl[0].kernel, p_ldg["mlp/dense0/kernel"]), (dl[2].bias, p_ldg["mlp/dense2/bias"])])
np.random.seed(46)
l_ref = gan.generator_training_step(opt_ref_g)
np.random.seed(46)
zin = np.random.normal(0, 1, (8, 145))
l_orc = O2.latent_gan_generator_losses(p_ldg, p_lg, T64(zin))
opt_orc_g.apply_gradients(zip(O.grads_of(l_orc["loss_sum"], p_lg), p_lg.values()))
gl = dense_layers(gan.generator)
check("lgan_g", l_ref, l_orc, [(gl[0].kernel, p_lg["mlp/dense0/kernel"]), (gl[2].kernel, p_lg["mlp/dense2/kernel"])])

np.savez_compressed(OUT, **out)
print("wrote", OUT, "%d arrays, %.0f KB" % (len(out), os.path.getsize(OUT) / 1024))
