"""-m gpu: whole networks and training-step losses/gradients of the CUDA path against the CPU oracle (run in
fp64 as the gold value) on identical seeded weights and inputs (256x256, the reference's resolution; tiny
batches so the oracle finishes in seconds).

Tolerances.  Forward outputs and every loss term: 1e-3 (max|a-b|/max|b|), the north-star parity bar.
Parameter gradients of the whole generator / generator step: relative L2 error <= 3e-2 per tensor.  The
gradient through ~25 LeakyReLU/normalisation stages at random init is chaotic at that level for ANY fp32
implementation: the fp32 CPU oracle itself differs from the fp64 oracle by 3e-3..1e-2 on these tensors
(measured, profiles/r01_precision_study.md), so 1e-3 is not a meaningful bar there.  Tight (2e-4) gradient
checks live at operator and block level (test_ops_gpu.py, test_blocks_tight below)."""
from collections import OrderedDict
import numpy as np
import pytest
import torch

from confignet_b200 import netspec
from oracle import confignet_oracle as O
from parity_utils import nerr, make_params, grads_cpu, grads_gpu, compare_grads

pytestmark = pytest.mark.gpu
FM = netspec.default_facemodel_inputs()


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _impl(name):
    from confignet_b200 import _lib as L
    return {"ffma": L.IMPL_FFMA, "auto": L.IMPL_AUTO}[name]


def _rot(n, seed):
    rng = np.random.RandomState(seed)
    r = np.zeros((n, 3), np.float32)
    r[:, 0] = np.pi * rng.uniform(-30, 30, n) / 180
    r[:, 1] = np.pi * rng.uniform(-10, 10, n) / 180
    return r


@pytest.mark.parametrize("impl,tol", [("ffma", 5e-4), ("auto", 1e-3)])
def test_generator_forward_and_grads(dev, impl, tol):
    from confignet_b200 import networks, ops
    B = 2
    p_cpu, grp = make_params(netspec.generator_spec(145, 256), 11, dev, dtype=torch.float64)
    rng = np.random.RandomState(0)
    z = rng.randn(B, 145).astype(np.float32); rot = _rot(B, 1)
    out_r = O.generator_forward(p_cpu, torch.tensor(z).double(), torch.tensor(rot).double(), 256)
    go = rng.randn(*out_r.shape).astype(np.float32)
    g_r = grads_cpu((out_r * torch.tensor(go).double()).sum(), p_cpu)
    ops.IMPL[0] = _impl(impl)
    try:
        out = networks.generator_forward(grp.params, torch.tensor(z, device=dev), rot, 256)
        g = grads_gpu((out * torch.tensor(go, device=dev)).sum(), grp)
    finally:
        ops.IMPL[0] = _impl("auto")
    assert out.shape == (B, 256, 256, 3)
    assert nerr(out, out_r) <= tol, nerr(out, out_r)
    u8, u8r = ops.to_uint8(out.detach()).cpu().numpy().astype(np.int32), O.to_uint8_images(out_r.detach().numpy()).astype(np.int32)
    assert np.abs(u8 - u8r).max() <= (1 if impl == "auto" else 1)
    compare_grads(g, g_r, 3e-2, "generator", metric="l2")


@pytest.mark.parametrize("impl,tol", [("ffma", 5e-4), ("auto", 1e-3)])
def test_discriminator_loss_with_r1(dev, impl, tol):
    from confignet_b200 import networks, ops
    B = 2
    p_cpu, grp = make_params(netspec.discriminator_spec(256), 21, dev, dtype=torch.float64)
    rng = np.random.RandomState(2)
    real = (rng.rand(B, 256, 256, 3).astype(np.float32) * 2 - 1)
    fake = np.tanh(rng.randn(B, 256, 256, 3)).astype(np.float32)
    l_r = O.compute_discriminator_loss(p_cpu, torch.tensor(real).double(), torch.tensor(fake).double())
    g_r = grads_cpu(l_r["loss_sum"], p_cpu)
    ops.IMPL[0] = _impl(impl)
    try:
        l = networks.compute_discriminator_loss(grp.params, torch.tensor(real, device=dev), torch.tensor(fake, device=dev))
        g = grads_gpu(l["loss_sum"], grp)
    finally:
        ops.IMPL[0] = _impl("auto")
    assert list(l.keys()) == list(l_r.keys())
    for k in l_r:
        assert abs(float(l[k]) - float(l_r[k])) <= tol * max(1.0, abs(float(l_r[k]))), (k, float(l[k]), float(l_r[k]))
    # max-norm bar on whole-network R1 gradients: 1e-2.  The fp32 CPU oracle itself is 3e-3..1e-2 from the fp64 oracle on
    # these tensors (LeakyReLU sign flips of near-zero pre-activations under five InstanceNorm stages, see the module
    # docstring); a 1e-7 change in one kernel's rounding (the flat 1x1 fromRGB kernel of round 2) moved the CUDA-core
    # path's worst tensor from 2.3e-3 to 4.5e-3, bit-reproducibly.  Tight bounds live at operator / block level.
    compare_grads(g, g_r, 1e-2, "discriminator")


def test_latent_discriminator_and_synthetic_encoder(dev):
    from confignet_b200 import networks
    B = 6
    p_ld, g_ld = make_params(netspec.latent_discriminator_spec(145, 4), 31, dev, dtype=torch.float64)
    p_se, g_se = make_params(netspec.synthetic_encoder_spec(FM, 2), 32, dev, dtype=torch.float64)
    rng = np.random.RandomState(3)
    params = [rng.rand(B, d[0]).astype(np.float32) for d in FM.values()]
    z_r = O.synthetic_encoder_forward(p_se, [torch.tensor(a).double() for a in params], FM)
    z = networks.synthetic_encoder_forward(g_se.params, [torch.tensor(a, device=dev) for a in params], FM)
    assert z.shape == (B, 145) and nerr(z, z_r) <= 1e-4
    # one concatenated matrix is split by input dims (synthetic_encoder.py:41-46)
    z2 = networks.synthetic_encoder_forward(g_se.params, torch.tensor(np.concatenate(params, 1), device=dev), FM)
    assert torch.equal(z, z2)
    real = rng.randn(B, 145).astype(np.float32)
    l_r = O.compute_latent_discriminator_loss(p_ld, torch.tensor(real).double(), z_r.detach())
    l = networks.compute_latent_discriminator_loss(g_ld.params, torch.tensor(real, device=dev), z.detach())
    for k in l_r:
        assert abs(float(l[k]) - float(l_r[k])) <= 1e-4 * max(1.0, abs(float(l_r[k]))), k
    compare_grads(grads_gpu(l["loss_sum"], g_ld), grads_cpu(l_r["loss_sum"], p_ld), 5e-4, "latent discriminator")


@pytest.mark.parametrize("impl,tol", [("ffma", 5e-4), ("auto", 1e-3)])
def test_generator_step_losses_and_grads(dev, impl, tol):
    """generator_training_step (confignet_first_stage.py:506-558): all loss terms and the gradients of
    generator, latent regressor and synthetic encoder."""
    from confignet_b200 import networks, ops
    ns, nr = 1, 1
    p_g, g_g = make_params(netspec.generator_spec(145, 256), 41, dev, dtype=torch.float64)
    p_lr, g_lr = make_params(netspec.latent_regressor_spec(145, 256), 42, dev, dtype=torch.float64)
    p_se, g_se = make_params(netspec.synthetic_encoder_spec(FM, 2), 43, dev, dtype=torch.float64)
    p_d, g_d = make_params(netspec.discriminator_spec(256), 44, dev, dtype=torch.float64)
    p_sd, g_sd = make_params(netspec.discriminator_spec(256), 45, dev, dtype=torch.float64)
    p_ld, g_ld = make_params(netspec.latent_discriminator_spec(145, 4), 46, dev, dtype=torch.float64)
    p_v, g_v = make_params(netspec.vgg19_spec(), 47, dev, perturb=0.0, vgg_like=True, dtype=torch.float64)
    rng = np.random.RandomState(4)
    fparams = [rng.rand(ns, d[0]).astype(np.float32) for d in FM.values()]
    gt = (rng.rand(ns, 256, 256, 3).astype(np.float32) * 2 - 1)
    masks = (rng.rand(ns, 256, 256) < 0.01).astype(np.uint8)
    real_lat = rng.randn(nr, 145).astype(np.float32)
    srot, rrot = _rot(ns, 5), _rot(nr, 6)
    batch = dict(facemodel_params=[torch.tensor(a).double() for a in fparams], synth_rotations=torch.tensor(srot).double(),
                 gt_imgs=torch.tensor(gt).double(), eye_masks=masks, real_latents=torch.tensor(real_lat).double(),
                 real_rotations=torch.tensor(rrot).double())
    l_r = O.generator_step_losses(p_g, p_lr, p_se, p_d, p_sd, p_ld, p_v, FM, batch)
    allp = OrderedDict()
    for pre, p in (("g/", p_g), ("lr/", p_lr), ("se/", p_se)):
        for k, v in p.items():
            allp[pre + k] = v
    g_r = grads_cpu(l_r["loss_sum"], allp)

    ops.IMPL[0] = _impl(impl)
    try:
        W = O.DEFAULT_LOSS_WEIGHTS
        gt_d = torch.tensor(gt, device=dev)
        rl = torch.tensor(real_lat, device=dev)
        synth_lat = networks.synthetic_encoder_forward(g_se.params, [torch.tensor(a, device=dev) for a in fparams], FM)
        o_s = networks.generator_forward(g_g.params, synth_lat, srot, 256)
        o_r = networks.generator_forward(g_g.params, rl, rrot, 256)
        l = OrderedDict()
        l["image_loss"] = W["image_loss_weight"] * networks.perceptual_loss(g_v.params, gt_d, o_s)
        l["eye_loss"] = W["eye_loss_weight"] * networks.eye_loss(gt_d, o_s, masks)
        for i, o in enumerate(networks.discriminator_forward(g_sd.params, o_s).values()):
            l["GAN_loss_synth_%d" % i] = networks.gan_g_loss(o)
        for i, o in enumerate(networks.discriminator_forward(g_d.params, o_r).values()):
            l["GAN_loss_real_%d" % i] = networks.gan_g_loss(o)
        l["latent_GAN_loss"] = W["domain_adverserial_loss_weight"] * networks.gan_g_loss(
            networks.latent_discriminator_forward(g_ld.params, synth_lat))
        labels = torch.cat((torch.cat((synth_lat, rl), 0),
                            W["latent_regressor_rot_weight"] * torch.tensor(np.concatenate((srot, rrot)), device=dev)), -1)
        l["latent_regression_loss"] = W["latent_regression_weight"] * networks.latent_regression_loss(
            g_lr.params, torch.cat((o_s, o_r), 0), labels)
        l["loss_sum"] = networks._sum(l.values())
        gs = torch.autograd.grad(l["loss_sum"], g_g.trainable_weights + g_lr.trainable_weights + g_se.trainable_weights,
                                 allow_unused=True)
    finally:
        ops.IMPL[0] = _impl("auto")
    assert list(l.keys()) == list(l_r.keys())
    for k in l_r:
        assert abs(float(l[k]) - float(l_r[k])) <= tol * max(1.0, abs(float(l_r[k]))), (k, float(l[k]), float(l_r[k]))
    names = list(allp.keys())
    g = OrderedDict((n, (torch.zeros_like(allp[n]) if q is None else q)) for n, q in zip(names, gs))
    compare_grads(g, g_r, 3e-2, "generator step", metric="l2")


def test_blocks_tight(dev):
    """Fused generator block (upsample -> conv -> LeakyReLU -> AdaIN with its MLP) and discriminator block
    (strided conv -> style + LeakyReLU -> InstanceNorm) incl. the R1 second-order path, at 2e-4 vs fp64."""
    from confignet_b200 import networks, ops
    torch.manual_seed(7)
    spec = OrderedDict()
    netspec._conv(spec, "blk/conv", (3, 3, 3), 16, 32)
    netspec._mlp(spec, "blk/adain", 2, 20, 24, 64)
    p_cpu, grp = make_params(spec, 71, dev, dtype=torch.float64)
    x = torch.randn(2, 4, 4, 4, 16); z = torch.randn(2, 20)
    xr = x.double().requires_grad_(True)
    yr = O.conv_adain(O.upsample_nearest2(xr), z.double(), p_cpu, "blk")
    gy = torch.randn(*yr.shape)
    ref = torch.autograd.grad(yr, [xr] + list(p_cpu.values()), gy.double())
    xg = x.to(dev).requires_grad_(True)
    y = networks.conv_adain(xg, z.to(dev), grp.params, "blk", 2)
    got = torch.autograd.grad(y, [xg] + grp.trainable_weights, gy.to(dev))
    assert nerr(y, yr) <= 2e-4
    for a, b in zip(got, ref):
        assert nerr(a, b) <= 2e-4
    # discriminator block + style head, R1-style second order wrt the block weights
    spec = OrderedDict()
    netspec._conv(spec, "b/conv", (3, 3), 8, 16)
    spec["b/in/gamma"] = ((16,), "ones"); spec["b/in/beta"] = ((16,), "zeros")
    netspec._dense(spec, "head", 32, 1)
    p_cpu, grp = make_params(spec, 72, dev, dtype=torch.float64)
    img = torch.randn(3, 12, 12, 8)

    def run(params, im, block, dense, cuda):
        im = im.requires_grad_(True)
        yb, style = block(im, params, "b", True)
        out = dense(style, params["head/kernel"], params["head/bias"]) if cuda else O.dense(style, params, "head")
        g, = torch.autograd.grad(out, im, torch.ones_like(out), create_graph=True)
        loss = (g ** 2).sum() + (yb ** 2).mean() + out.sum()
        return loss, torch.autograd.grad(loss, list(params.values()), allow_unused=True)
    l_r, g_r = run(p_cpu, img.double(), O.discr_block, None, False)
    l_g, g_g = run(grp.params, img.to(dev), networks.discr_block, ops.dense, True)
    assert abs(float(l_g) - float(l_r)) <= 2e-4 * abs(float(l_r))
    for a, b in zip(g_g, g_r):
        assert nerr(a, b) <= 2e-4


def test_graph_replayed_steps_match_eager_steps(dev):
    """The stage-1 D / synth-D / G steps run as CUDA-graph replays after two eager calls (runtime.GraphedFn).  Two
    models with identical weights and the same NumPy stream - one replaying, one always eager - must report the same
    losses step after step (tolerance: the split-K atomics' summation order) and end with the same weights."""
    from confignet_b200.confignet_first_stage import ConfigNetFirstStage
    from confignet_b200.runtime import KerasAdam
    from confignet_b200.synthetic_data import SyntheticDataset
    from confignet_b200 import netspec

    def run(graphs):
        cfg = {"output_shape": (256, 256, 3), "batch_size": 2, "facemodel_inputs": netspec.default_facemodel_inputs(),
               "cuda_graphs": graphs}
        model = ConfigNetFirstStage(cfg, device=dev)
        real, synth = SyntheticDataset(6, 256, seed=1), SyntheticDataset(6, 256, seed=2)
        d_opt, g_opt = KerasAdam(**model.config["optimizer"]), KerasAdam(**model.config["optimizer"])
        np.random.seed(5)
        hist = []
        for _ in range(5):
            d = model.discriminator_training_step(real, d_opt)
            sd = model.synth_discriminator_training_step(synth, d_opt)
            g = model.generator_training_step(real, synth, g_opt)
            model.update_smoothed_weights()
            hist.append([float(v) for dct in (d, sd, g) for v in dct.values()])
        replayed = [k for k, (_, v) in model._graphs.items() if v.graph is not None]
        return np.array(hist), model.get_weights(), replayed, d_opt.iterations

    h_g, w_g, replayed, it_g = run(True)
    h_e, w_e, none, it_e = run(False)
    h_e2, w_e2, _, _ = run(False)
    assert len(replayed) == 3 and not none and it_g == it_e == 10
    assert np.isfinite(h_g).all()
    scale = np.abs(h_e).max(axis=1)
    d_graph = np.abs(h_g - h_e).max(axis=1) / scale          # per step
    d_eager = np.abs(h_e2 - h_e).max(axis=1) / scale         # two eager runs: the noise floor (split-K atomics + Adam, beta_1 = 0)
    print("graph vs eager per step:", d_graph, " eager vs eager:", d_eager)
    # steps 1-2 run eagerly in both; step 3 is the captured step (same weights going in); 4-5 are pure replays
    # (B = 2 on untrained networks is chaotic: two eager runs are already 3e-2..2e-1 apart from step 3 on, and which
    # step a pair happens to agree on varies from run to run - so later steps are held against the pair's worst step)
    # Measured on B200: eager pairs differ by 3e-7, 1e-3, 5e-2..6e-2, 3e-2..1e-1, 6e-2..2e-1 at steps 1..5, the replayed
    # model by 2e-7, 2e-3, 7e-2, 1.5e-1..1.9e-1, 1.9e-1..2.4e-1 - the same growth.  The bounds below leave room for
    # that chaos and still catch a replay that computes something else (stale inputs, a missing kernel, a wrong
    # learning rate: those show up as O(1) differences at the captured step).
    noise = max(0.5, 4 * float(d_eager[2:].max()))
    assert d_graph[0] <= 1e-3 and d_graph[1] <= 2e-2, (d_graph, d_eager)
    assert d_graph[2] <= 0.25 and np.all(d_graph[3:] <= noise), (d_graph, d_eager)
    for name in w_e:
        a = np.concatenate([x.ravel() for x in w_g[name]]); b = np.concatenate([x.ravel() for x in w_e[name]])
        c = np.concatenate([x.ravel() for x in w_e2[name]])
        assert np.linalg.norm(a - b) <= max(1e-1 * np.linalg.norm(b), 4 * np.linalg.norm(c - b)), name
