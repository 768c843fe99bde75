"""-m gpu: second-stage / fine-tune / LatentGAN pieces of the CUDA path against the CPU oracle (fp64 gold) on
identical seeded weights and inputs.  Tolerances as in test_networks_gpu.py: 1e-4..2e-4 for single operators,
1e-3 for network outputs and every loss term, relative L2 <= 3e-2 for whole-network parameter gradients."""
from collections import OrderedDict
import numpy as np
import pytest
import torch

from confignet_b200 import netspec
from oracle import confignet_oracle as O
from oracle import confignet_oracle_stage2 as O2
from parity_utils import nerr, make_params, grads_cpu, grads_gpu, compare_grads

pytestmark = pytest.mark.gpu
FM = netspec.default_facemodel_inputs()


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _rot(n, seed):
    rng = np.random.RandomState(seed)
    r = np.zeros((n, 3), np.float32)
    r[:, 0] = np.pi * rng.uniform(-30, 30, n) / 180
    r[:, 1] = np.pi * rng.uniform(-10, 10, n) / 180
    return r


def _encoder_pair(dev, seed):
    from confignet_b200.runtime import ParamGroup
    arrays = netspec.init_real_encoder_params(145, seed)
    p_cpu = O.to_torch(arrays, dtype=torch.float64, requires_grad=True)
    for k, v in p_cpu.items():
        if not netspec.is_trainable(k):
            v.requires_grad_(False)
    return p_cpu, ParamGroup(arrays, dev, trainable=netspec.is_trainable)


# ------------------------------------------------------------------------------------------------ operators
@pytest.mark.parametrize("res,relu", [(False, True), (True, True), (True, False), (False, False)])
def test_bn_act(dev, res, relu):
    from confignet_b200 import ops
    rng = np.random.RandomState(0)
    x = rng.randn(3, 5, 7, 24).astype(np.float32)
    r = rng.randn(3, 5, 7, 24).astype(np.float32) if res else None
    par = [rng.uniform(0.5, 1.5, 24), rng.randn(24), rng.randn(24) * 0.3, rng.uniform(0.5, 2.0, 24)]
    gy = rng.randn(3, 5, 7, 24).astype(np.float32)
    tx = torch.tensor(x).double().requires_grad_(True)
    tp = [torch.tensor(a).double().requires_grad_(i < 2) for i, a in enumerate(par)]
    tr = torch.tensor(r).double().requires_grad_(True) if res else None
    y = (tx - tp[2]) / torch.sqrt(tp[3] + O2.BN_EPS) * tp[0] + tp[1]
    if res:
        y = y + tr
    if relu:
        y = torch.relu(y)
    ins = [tx, tp[0], tp[1]] + ([tr] if res else [])
    ref = torch.autograd.grad(y, ins, torch.tensor(gy).double())
    gx = torch.tensor(x, device=dev).requires_grad_(True)
    gp = [torch.tensor(a.astype(np.float32), device=dev).requires_grad_(i < 2) for i, a in enumerate(par)]
    gr = torch.tensor(r, device=dev).requires_grad_(True) if res else None
    out = ops.bn_act(gx, gp[0], gp[1], gp[2], gp[3], gr, relu)
    got = torch.autograd.grad(out, [gx, gp[0], gp[1]] + ([gr] if res else []), torch.tensor(gy, device=dev))
    assert nerr(out, y) <= 1e-5
    for a, b in zip(got, ref):
        assert nerr(a, b) <= 1e-4


@pytest.mark.parametrize("shape", [(2, 8, 8, 4), (1, 7, 9, 3), (2, 16, 16, 64)])
def test_maxpool3s2_and_avgpool(dev, shape):
    from confignet_b200 import ops
    rng = np.random.RandomState(1)
    x = rng.randn(*shape).astype(np.float32)
    x[0, :3, :3] = -np.abs(x[0, :3, :3])           # corner windows where the padded zero is the maximum
    tx = torch.tensor(x).double().requires_grad_(True)
    y = O2.maxpool_3x3_s2_pad1(tx)
    gy = rng.randn(*y.shape).astype(np.float32)
    ref, = torch.autograd.grad(y, tx, torch.tensor(gy).double())
    gx = torch.tensor(x, device=dev).requires_grad_(True)
    out = ops.maxpool3s2(gx)
    got, = torch.autograd.grad(out, gx, torch.tensor(gy, device=dev))
    assert out.shape == y.shape and torch.equal(out.detach().cpu().double(), y.detach())
    assert nerr(got, ref) <= 1e-6
    ya = tx.mean(dim=(1, 2))
    ga = rng.randn(*ya.shape).astype(np.float32)
    refa, = torch.autograd.grad(ya, tx, torch.tensor(ga).double())
    outa = ops.global_avg_pool(gx)
    gota, = torch.autograd.grad(outa, gx, torch.tensor(ga, device=dev))
    assert nerr(outa, ya) <= 1e-5 and nerr(gota, refa) <= 1e-6


def test_stem_conv_explicit_padding(dev):
    """ZeroPadding2D(3) + 7x7/s2 VALID (ResNet50 conv1) forward / dgrad / wgrad through the C ABI."""
    from confignet_b200 import ops
    rng = np.random.RandomState(2)
    x = rng.randn(2, 32, 32, 3).astype(np.float32)
    w = (rng.randn(7, 7, 3, 64) * 0.1).astype(np.float32); b = rng.randn(64).astype(np.float32)
    tx, tw, tb = [torch.tensor(a).double().requires_grad_(True) for a in (x, w, b)]
    y = O2.conv_valid_padded(tx, tw, tb, 2, 3)
    gy = rng.randn(*y.shape).astype(np.float32)
    ref = torch.autograd.grad(y, (tx, tw, tb), torch.tensor(gy).double())
    gx, gw, gb = [torch.tensor(a, device=dev).requires_grad_(True) for a in (x, w, b)]
    out = ops.conv_act(gx, gw, gb, stride=2, pad=3)
    got = torch.autograd.grad(out, (gx, gw, gb), torch.tensor(gy, device=dev))
    assert out.shape == y.shape and nerr(out, y) <= 2e-4
    for a, r in zip(got, ref):
        assert nerr(a, r) <= 2e-4


def test_euler_and_rotation_gradient(dev):
    """euler_angles_to_matrix + transform_3d_grid_tf differentiated wrt the ANGLES (the encoder's rotation output is
    trained through the resampler, confignet_second_stage.py:169-170) and wrt the volume."""
    from confignet_b200 import ops
    rng = np.random.RandomState(3)
    B, S, C = 3, 16, 8
    ang = _rot(B, 7); ang[:, 2] = [0.05, -0.1, 0.2]
    # a smooth volume: the trilinear resampler's derivative wrt the coordinates is piecewise constant in the data
    grid = rng.randn(B, S, S, S, C).astype(np.float32)
    ta = torch.tensor(ang).double().requires_grad_(True)
    tg = torch.tensor(grid).double().requires_grad_(True)
    y = O.transform_3d_grid(tg, O.euler_angles_to_matrix(ta))
    gy = rng.randn(*y.shape).astype(np.float32)
    ra, rg = torch.autograd.grad(y, (ta, tg), torch.tensor(gy).double())
    ga = torch.tensor(ang, device=dev).requires_grad_(True)
    gg = torch.tensor(grid, device=dev).requires_grad_(True)
    R = ops.euler_to_matrix(ga)
    assert nerr(R, O.euler_angles_to_matrix(ta).reshape(B, 9)) <= 1e-6
    out = ops.rotate3d(gg, R)
    a, g = torch.autograd.grad(out, (ga, gg), torch.tensor(gy, device=dev))
    assert nerr(out, y) <= 1e-5
    assert nerr(g, rg) <= 1e-5
    assert nerr(a, ra) <= 2e-4, (a, ra)


@pytest.mark.parametrize("B,J", [(2, 148), (7, 148), (32, 148), (5, 10)])
def test_norm_latent_loss(dev, B, J):
    from confignet_b200 import ops
    rng = np.random.RandomState(4)
    o, l = rng.randn(B, J).astype(np.float32), (rng.randn(B, J) * 1.5 + 0.3).astype(np.float32)
    to, tl = torch.tensor(o).double().requires_grad_(True), torch.tensor(l).double().requires_grad_(True)
    ref = O2.normalized_regression(to, tl, 10.0)
    ro, rl = torch.autograd.grad(ref, (to, tl))
    go, gl = torch.tensor(o, device=dev).requires_grad_(True), torch.tensor(l, device=dev).requires_grad_(True)
    out = ops.norm_latent_loss(go, gl, 10.0, 3)
    a, b = torch.autograd.grad(out * 1.5, (go, gl))
    assert abs(float(out) - float(ref)) <= 1e-5 * abs(float(ref))
    assert nerr(a, ro * 1.5) <= 1e-4 and nerr(b, rl * 1.5) <= 1e-4


def test_vggface_preprocess(dev):
    from confignet_b200 import ops
    x = torch.rand(2, 4, 4, 3) * 2 - 1
    tx = x.double().requires_grad_(True)
    y = O2.vggface_preprocess(tx)
    gy = torch.randn(2, 4, 4, 3)
    r, = torch.autograd.grad(y, tx, gy.double())
    gx = x.to(dev).requires_grad_(True)
    out = ops.vgg_preprocess(gx, face=True)
    g, = torch.autograd.grad(out, gx, gy.to(dev))
    assert nerr(out, y) <= 1e-6 and nerr(g, r) <= 1e-6


# ------------------------------------------------------------------------------------------------ networks
def test_real_encoder_forward_and_grads(dev):
    """RealEncoder (ResNet50 + heads, real_encoder.py:23-34) at 256x256: outputs 1e-3, parameter gradients L2."""
    from confignet_b200 import networks
    p_cpu, grp = _encoder_pair(dev, 81)
    rng = np.random.RandomState(5)
    img = (rng.rand(2, 256, 256, 3).astype(np.float32) * 2 - 1)
    ti = torch.tensor(img).double().requires_grad_(True)
    emb_r, rot_r = O2.real_encoder_forward(p_cpu, ti)
    ge, gr = rng.randn(2, 145).astype(np.float32), rng.randn(2, 3).astype(np.float32)
    loss_r = (emb_r * torch.tensor(ge).double()).sum() + (rot_r * torch.tensor(gr).double()).sum()
    tr = OrderedDict((k, v) for k, v in p_cpu.items() if netspec.is_trainable(k))
    g_r = grads_cpu(loss_r, OrderedDict(list(tr.items()) + [("__img", ti)]))
    mult = torch.tensor(O2.rotation_range_multiplier(), dtype=torch.float32, device=dev)
    gi = torch.tensor(img, device=dev).requires_grad_(True)
    emb, rot = networks.real_encoder_forward(grp.params, gi, mult)
    loss = (emb * torch.tensor(ge, device=dev)).sum() + (rot * torch.tensor(gr, device=dev)).sum()
    gs = torch.autograd.grad(loss, grp.trainable_weights + [gi], allow_unused=True)
    assert nerr(emb, emb_r) <= 1e-3, nerr(emb, emb_r)
    assert nerr(rot, rot_r) <= 1e-3, nerr(rot, rot_r)
    names = list(tr.keys()) + ["__img"]
    g = OrderedDict((n, (torch.zeros_like(g_r[n]) if q is None else q)) for n, q in zip(names, gs))
    compare_grads(g, g_r, 3e-2, "real encoder", metric="l2")


def test_stage2_generator_step_losses_and_grads(dev):
    """generator_training_step of the second stage (confignet_second_stage.py:149-218): every loss term and the
    gradients of generator, latent regressor, synthetic encoder and real encoder (incl. through the rotation)."""
    from confignet_b200 import networks, ops
    ns, nr = 2, 2
    p_g, g_g = make_params(netspec.generator_spec(145, 256), 41, dev, dtype=torch.float64)
    p_lr, g_lr = make_params(netspec.latent_regressor_spec(145, 256), 42, dev, dtype=torch.float64)
    p_se, g_se = make_params(netspec.synthetic_encoder_spec(FM, 2), 43, dev, dtype=torch.float64)
    p_d, g_d = make_params(netspec.discriminator_spec(256), 44, dev, dtype=torch.float64)
    p_sd, g_sd = make_params(netspec.discriminator_spec(256), 45, dev, dtype=torch.float64)
    p_ld, g_ld = make_params(netspec.latent_discriminator_spec(145, 4), 46, dev, dtype=torch.float64)
    p_v, g_v = make_params(netspec.vgg19_spec(), 47, dev, perturb=0.0, vgg_like=True, dtype=torch.float64)
    p_e, g_e = _encoder_pair(dev, 48)
    rng = np.random.RandomState(4)
    fparams = [rng.rand(ns, d[0]).astype(np.float32) for d in FM.values()]
    simgs = (rng.rand(ns, 256, 256, 3).astype(np.float32) * 2 - 1)
    rimgs = (rng.rand(nr, 256, 256, 3).astype(np.float32) * 2 - 1)
    masks = (rng.rand(ns, 256, 256) < 0.01).astype(np.uint8)
    srot = _rot(ns, 5)
    W = dict(O.DEFAULT_LOSS_WEIGHTS); W["image_loss_weight"] = 5e-4
    batch = dict(facemodel_params=[torch.tensor(a).double() for a in fparams], synth_rotations=torch.tensor(srot).double(),
                 synth_imgs=torch.tensor(simgs).double(), eye_masks=masks, real_imgs=torch.tensor(rimgs).double())
    l_r = O2.stage2_generator_step_losses(p_g, p_lr, p_se, p_e, p_d, p_sd, p_ld, p_v, FM, batch, weights=W)
    allp = OrderedDict()
    for pre, p in (("g/", p_g), ("lr/", p_lr), ("se/", p_se), ("enc/", p_e)):
        for k, v in p.items():
            if v.requires_grad:
                allp[pre + k] = v
    g_r = grads_cpu(l_r["loss_sum"], allp)

    mult = torch.tensor(O2.rotation_range_multiplier(), dtype=torch.float32, device=dev)
    s_d, r_d = torch.tensor(simgs, device=dev), torch.tensor(rimgs, device=dev)
    synth_lat = networks.synthetic_encoder_forward(g_se.params, [torch.tensor(a, device=dev) for a in fparams], FM)
    o_s = networks.generator_forward(g_g.params, synth_lat, srot, 256)
    real_lat, real_rot = networks.real_encoder_forward(g_e.params, r_d, mult)
    o_r = networks.generator_forward(g_g.params, real_lat, real_rot, 256)
    l = OrderedDict()
    l["image_loss_synth"] = W["image_loss_weight"] * networks.perceptual_loss(g_v.params, s_d, o_s)
    l["image_loss_real"] = W["image_loss_weight"] * networks.perceptual_loss(g_v.params, r_d, o_r)
    l["eye_loss"] = W["eye_loss_weight"] * networks.eye_loss(s_d, o_s, masks)
    for i, o in enumerate(networks.discriminator_forward(g_sd.params, o_s).values()):
        l["GAN_loss_synth_%d" % i] = networks.gan_g_loss(o)
    for i, o in enumerate(networks.discriminator_forward(g_d.params, o_r).values()):
        l["GAN_loss_real_%d" % i] = networks.gan_g_loss(o)
    l["latent_GAN_loss"] = W["domain_adverserial_loss_weight"] * networks.gan_d_loss_mixed(
        networks.latent_discriminator_forward(g_ld.params, real_lat), networks.latent_discriminator_forward(g_ld.params, synth_lat))
    labels = torch.cat((torch.cat((synth_lat, real_lat), 0),
                        W["latent_regressor_rot_weight"] * torch.cat((torch.tensor(srot, device=dev), real_rot), 0)), -1)
    l["latent_regression_loss"] = networks.normalized_latent_regression_loss(
        g_lr.params, torch.cat((o_s, o_r), 0), labels, W["latent_regression_weight"])
    l["loss_sum"] = networks._sum(l.values())
    gs = torch.autograd.grad(l["loss_sum"], g_g.trainable_weights + g_lr.trainable_weights + g_se.trainable_weights +
                             g_e.trainable_weights, allow_unused=True)
    assert list(l.keys()) == list(l_r.keys())
    for k in l_r:
        assert abs(float(l[k]) - float(l_r[k])) <= 1e-3 * max(1.0, abs(float(l_r[k]))), (k, float(l[k]), float(l_r[k]))
    names = list(allp.keys())
    assert len(names) == len(gs)
    g = OrderedDict((n, (torch.zeros_like(allp[n]) if q is None else q)) for n, q in zip(names, gs))
    compare_grads(g, g_r, 3e-2, "stage-2 generator step", metric="l2")


def test_fine_tune_losses_and_grads(dev):
    """One iteration of fine_tune_on_img's loss (confignet_second_stage.py:360-390): VGG19 + VGGFace(VGG16) perceptual
    terms, GAN, domain-adversarial and normalised regression terms; gradients wrt generator, embeddings, rotations."""
    from confignet_b200 import networks
    n = 2
    p_g, g_g = make_params(netspec.generator_spec(145, 256), 51, dev, dtype=torch.float64)
    p_lr, g_lr = make_params(netspec.latent_regressor_spec(145, 256), 52, dev, dtype=torch.float64)
    p_d, g_d = make_params(netspec.discriminator_spec(256), 54, dev, dtype=torch.float64)
    p_ld, g_ld = make_params(netspec.latent_discriminator_spec(145, 4), 56, dev, dtype=torch.float64)
    p_v, g_v = make_params(netspec.vgg19_spec(), 57, dev, perturb=0.0, vgg_like=True, dtype=torch.float64)
    p_f, g_f = make_params(netspec.vgg16_spec(), 58, dev, perturb=0.0, vgg_like=True, dtype=torch.float64)
    rng = np.random.RandomState(9)
    imgs = (rng.rand(n, 256, 256, 3).astype(np.float32) * 2 - 1)
    lo, hi = 7, 37                                           # blendshape_values slice of the sorted latent layout
    pre, expr, post = rng.randn(1, lo).astype(np.float32), rng.randn(n, hi - lo).astype(np.float32), rng.randn(1, 145 - hi).astype(np.float32)
    rot = _rot(n, 3)
    W = dict(O.DEFAULT_LOSS_WEIGHTS); W["image_loss_weight"] = 5e-4
    tv = [torch.tensor(a).double().requires_grad_(True) for a in (pre, expr, post, rot)]
    l_r = O2.fine_tune_losses(p_g, p_lr, p_d, p_ld, p_v, p_f, torch.tensor(imgs).double(), *tv, weights=W)
    allp = OrderedDict(("g/" + k, v) for k, v in p_g.items())
    for k, v in zip(("pre", "expr", "post", "rot"), tv):
        allp[k] = v
    g_r = grads_cpu(l_r["loss_sum"], allp)

    gv = [torch.tensor(a, device=dev).requires_grad_(True) for a in (pre, expr, post, rot)]
    im = torch.tensor(imgs, device=dev)
    emb = torch.cat((gv[0].expand(n, -1), gv[1], gv[2].expand(n, -1)), dim=1)
    out = networks.generator_forward(g_g.params, emb, gv[3], 256)
    l = OrderedDict()
    l["image_loss_real"] = 0.5 * W["image_loss_weight"] * networks.perceptual_loss(g_v.params, im, out)
    l["face_reco_loss"] = 0.5 * W["image_loss_weight"] * networks.perceptual_loss(g_f.params, out, im, model_type="VGGFace")
    for i, o in enumerate(networks.discriminator_forward(g_d.params, out).values()):
        l["GAN_loss_real_%d" % i] = networks.gan_g_loss(o)
    l["latent_GAN_loss"] = W["domain_adverserial_loss_weight"] * networks.gan_d_loss(1, networks.latent_discriminator_forward(g_ld.params, emb))
    labels = torch.cat((emb, W["latent_regressor_rot_weight"] * gv[3]), dim=-1)
    l["latent_regression_loss"] = networks.normalized_latent_regression_loss(g_lr.params, out, labels, W["latent_regression_weight"])
    l["loss_sum"] = networks._sum(l.values())
    gs = torch.autograd.grad(l["loss_sum"], g_g.trainable_weights + gv, allow_unused=True)
    assert list(l.keys()) == list(l_r.keys())
    for k in l_r:
        assert abs(float(l[k]) - float(l_r[k])) <= 1e-3 * max(1.0, abs(float(l_r[k]))), (k, float(l[k]), float(l_r[k]))
    g = OrderedDict((nm, (torch.zeros_like(allp[nm]) if q is None else q)) for nm, q in zip(allp.keys(), gs))
    compare_grads(g, g_r, 3e-2, "fine-tune", metric="l2")


def test_latent_gan_losses_and_grads(dev):
    """LatentGAN steps (latent_gan.py:117-165) at the reference's sizes (latent 145, hidden 217, batch 32)."""
    from confignet_b200 import latent_gan as LG, networks
    p_g, g_g = make_params(netspec.latent_gan_mlp_spec(145), 61, dev, dtype=torch.float64)
    p_d, g_d = make_params(netspec.latent_gan_mlp_spec(145, num_out=1), 62, dev, dtype=torch.float64)
    rng = np.random.RandomState(6)
    real, z = rng.randn(32, 145).astype(np.float32), rng.randn(32, 145).astype(np.float32)
    l_r = O2.latent_gan_discriminator_losses(p_d, p_g, torch.tensor(real).double(), torch.tensor(z).double())
    with torch.no_grad():
        fake = LG._mlp_forward(g_g.params, torch.tensor(z, device=dev), 3)
    rd = torch.tensor(real, device=dev).requires_grad_(True)
    o_real = LG._mlp_forward(g_d.params, rd, 3, second_order=True)
    o_fake = LG._mlp_forward(g_d.params, fake, 3, second_order=True)
    l = OrderedDict()
    l["GAN_loss_real"] = networks.gan_d_loss(1, o_real)
    l["GAN_loss_fake"] = networks.gan_d_loss(0, o_fake)
    l["gp_loss"] = networks.gradient_regularization(o_real, rd)
    l["loss_sum"] = networks._sum(l.values())
    for k in l_r:
        assert abs(float(l[k]) - float(l_r[k])) <= 1e-4 * max(1.0, abs(float(l_r[k]))), k
    compare_grads(grads_gpu(l["loss_sum"], g_d), grads_cpu(l_r["loss_sum"], p_d), 5e-4, "latent GAN discriminator")
    lg_r = O2.latent_gan_generator_losses(p_d, p_g, torch.tensor(z).double())
    lg = networks.gan_g_loss(LG._mlp_forward(g_d.params, LG._mlp_forward(g_g.params, torch.tensor(z, device=dev), 3), 3))
    assert abs(float(lg) - float(lg_r["gan_loss"])) <= 1e-4
    compare_grads(grads_gpu(lg, g_g), grads_cpu(lg_r["loss_sum"], p_g), 5e-4, "latent GAN generator")


# ------------------------------------------------------------------------------------------------ class surface
def test_confignet_class_surface_end_to_end(dev, tmp_path):
    """ConfigNet (second stage): one full training iteration through the public API, encode_images / generate_images
    round trip shapes and dtypes, save/load, fine_tune_on_img, LatentGAN on the extracted embeddings."""
    import confignet_b200
    from confignet_b200 import ConfigNet, LatentGAN
    from confignet_b200.synthetic_data import SyntheticDataset
    np.random.seed(0)
    cfg = {"output_shape": (256, 256, 3), "batch_size": 4, "facemodel_inputs": netspec.default_facemodel_inputs(),
           "image_loss_weight": 5e-4}
    model = ConfigNet(cfg, device=dev)
    assert model.config["model_type"] == "ConfigNet" and model.config["latent_dim"] == 145
    real, synth = SyntheticDataset(8, 256, seed=1), SyntheticDataset(8, 256, seed=2)
    w0 = model.encoder.get_weights()
    model.train(real, synth, None, None, str(tmp_path), None, n_steps=1)
    for hist, keys in ((model.g_losses, ["image_loss_synth", "image_loss_real", "eye_loss", "latent_GAN_loss",
                                        "latent_regression_loss", "loss_sum"]),
                       (model.d_losses, ["GAN_loss_real_0", "gp_loss_5", "loss_sum"]),
                       (model.latent_d_losses, ["GAN_loss_real", "gp_loss", "loss_sum"])):
        for k in keys:
            assert k in hist and np.isfinite(hist[k][-1]), k
    w1 = model.encoder.get_weights()
    names = netspec.real_encoder_keras_order(145)                       # get_weights() lists the nested ResNet50 keras' way
    moved = [n for n, a, b in zip(names, w0, w1) if not np.array_equal(a, b)]
    assert moved and all(netspec.is_trainable(n) for n in moved)        # moving statistics are never updated
    emb, rot = model.encode_images(real.imgs[:3])
    assert emb.shape == (3, 145) and rot.shape == (3, 3) and emb.dtype == np.float32 and rot.dtype == np.float32
    lim = np.pi * np.array([30, 10, 0]) / 180
    assert np.all(np.abs(rot) <= lim + 1e-6)
    emb_f, _ = model.encode_images(real.imgs[:3].astype(np.float32) / 127.5 - 1.0)
    assert np.abs(emb_f - emb).max() <= 1e-4 * np.abs(emb).max()
    imgs = model.generate_images(emb, rot)
    assert imgs.shape == (3, 256, 256, 3) and imgs.dtype == np.uint8
    new = model.set_facemodel_param_in_latents(emb, "blendshape_values", np.zeros((1, 62), np.float32))
    idx = list(model.get_facemodel_param_idxs_in_latent("blendshape_values"))
    assert idx == list(range(7, 37)) and np.array_equal(np.delete(new, idx, 1), np.delete(emb, idx, 1))
    model.save(str(tmp_path), "m")
    again = confignet_b200.load_confignet(str(tmp_path / "m.json"), device=dev)
    assert isinstance(again, ConfigNet)
    imgs2 = again.generate_images(emb, rot)
    diff = np.abs(imgs2.astype(np.int32) - imgs.astype(np.int32))
    assert diff.max() == 0, ("save/load round trip changed the images", int(diff.max()), float((diff > 0).mean()))
    assert np.array_equal(model.generate_images(emb, rot), imgs)                                # bit-reproducible
    e2, r2 = model.fine_tune_on_img(real.imgs[:2], n_iters=2)
    assert e2.shape == (2, 145) and r2.shape == (2, 3) and model.generator_fine_tuned is not None
    assert np.array_equal(e2[0, :7], e2[1, :7]) and np.array_equal(e2[0, 37:], e2[1, 37:])      # shared pre/post embeddings
    assert len(model.fine_tune_losses) == 2 and all(np.isfinite(float(l["loss_sum"])) for l in model.fine_tune_losses)
    assert not np.array_equal(model.generate_images(emb, rot), imgs)                            # fine-tuned generator is used
    e3, _ = model.fine_tune_on_img(real.imgs[:2], n_iters=1, force_neutral_expression=True)
    neutral = model.synthetic_encoder.per_facemodel_input_mlps["blendshape_values"].predict(np.zeros((1, 62), np.float32))
    assert np.abs(e3[:, 7:37] - neutral).max() <= 1e-6                                          # expression stays frozen
    gan = LatentGAN({"latent_dim": 145, "batch_size": 8}, device=dev)
    gan.train(real, model, str(tmp_path), None, n_iters=2)
    lat = gan.generate_latents(5, truncation=0.7)
    assert lat.shape == (5, 145) and lat.dtype == np.float32 and np.isfinite(lat).all()
    gan.save(str(tmp_path), "gan")
    gan2 = LatentGAN.load(str(tmp_path / "gan.json"), device=dev)
    np.random.seed(3); a = gan.generate_latents(4)
    np.random.seed(3); b = gan2.generate_latents(4)
    assert np.array_equal(a, b)
