"""CPU-only tests (-m "not gpu") of the second-stage additions: the explicit-padding (ResNet50 stem) plan geometry,
the stage-2 oracle pieces against independent NumPy restatements, the class-surface host logic, and the
batch-statistics loss under data parallelism over gloo (world_size 2)."""
import ctypes
import os
import subprocess
import sys
from collections import OrderedDict

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from oracle import confignet_oracle as O
from oracle import confignet_oracle_stage2 as O2
from confignet_b200 import netspec


@pytest.mark.parametrize("cfg", [(1, (16, 16), 3, 4, 7, 2, 3), (2, (9, 11), 2, 3, 7, 2, 3), (1, (6, 6), 2, 2, 3, 1, 0),
                                 (1, (8, 8), 2, 3, 3, 2, 1)])
def test_plan_geometry_explicit_padding(cfg):
    """ZeroPadding2D(p) + VALID conv (keras-applications ResNet50 stem): forward, dgrad and wgrad plans evaluated on
    the host equal the oracle's conv / its autograd gradients."""
    from confignet_b200 import _lib as L
    lib = L.load_hooks()          # cn_debug_conv_host lives in the hooks build only
    B, dims, cin, cout, k, s, pad = cfg
    rng = np.random.RandomState(1)
    d = L.make_conv_desc(2, B, dims, cin, cout, [k, k], s, 1, pad)
    x = rng.randn(B, *dims, cin).astype(np.float32)
    w = rng.randn(k, k, cin, cout).astype(np.float32)
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    wt = torch.tensor(w, dtype=torch.float64, requires_grad=True)
    y = O2.conv_valid_padded(xt, wt, None, s, pad)
    od = (ctypes.c_int * 3)()
    assert lib.cn_conv_out_dims(ctypes.byref(d), od) == 0 and tuple(od[:2]) == tuple(y.shape[1:3])
    gy = rng.randn(*y.shape).astype(np.float32)
    gx, gw = torch.autograd.grad(y, (xt, wt), torch.tensor(gy, dtype=torch.float64))
    fp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    out = np.zeros(y.shape, np.float32)
    assert lib.cn_debug_conv_host(ctypes.byref(d), 0, fp(x), fp(w), fp(out)) == 0
    ogx = np.full(x.shape, 7.0, np.float32)
    assert lib.cn_debug_conv_host(ctypes.byref(d), 1, fp(gy), fp(w), fp(ogx)) == 0
    ogw = np.zeros(w.shape, np.float32)
    assert lib.cn_debug_conv_host(ctypes.byref(d), 2, fp(x), fp(gy), fp(ogw)) == 0
    assert np.abs(out - y.detach().numpy()).max() < 1e-4
    assert np.abs(ogx - gx.numpy()).max() < 1e-4
    assert np.abs(ogw - gw.numpy()).max() < 1e-4


def test_oracle_normalized_regression_against_numpy():
    """compute_normalized_latent_regression_loss (confignet_second_stage.py:93-107) restated line by line in NumPy."""
    rng = np.random.RandomState(0)
    out, lab = rng.randn(6, 10), rng.randn(6, 10) * 2 + 0.5
    den = np.sqrt(lab.var(axis=0, keepdims=True) + 1e-3)
    den = np.concatenate((den[:, :-3], np.ones((1, 3))), axis=1)
    o_n = out.mean(axis=0) + (out - out.mean(axis=0)) / den
    l_n = lab.mean(axis=0) + (lab - lab.mean(axis=0)) / den
    want = ((l_n - o_n) ** 2).mean(axis=-1).mean() * 10.0
    got = float(O2.normalized_regression(torch.tensor(out), torch.tensor(lab), 10.0))
    assert abs(got - want) < 1e-12


def test_oracle_resnet_pieces_against_numpy():
    rng = np.random.RandomState(1)
    x = rng.randn(1, 5, 6, 2)
    got = O2.maxpool_3x3_s2_pad1(torch.tensor(x)).numpy()
    xp = np.zeros((1, 7, 8, 2)); xp[:, 1:-1, 1:-1] = x
    oh, ow = (7 - 3) // 2 + 1, (8 - 3) // 2 + 1
    want = np.zeros((1, oh, ow, 2))
    for i in range(oh):
        for j in range(ow):
            want[0, i, j] = xp[0, 2 * i:2 * i + 3, 2 * j:2 * j + 3].reshape(9, 2).max(axis=0)
    assert got.shape == want.shape and np.abs(got - want).max() == 0
    # ResNet50 feature size and block structure on a small input (64x64 -> 2x2x2048 -> 2048)
    p = O.to_torch(netspec.init_real_encoder_params(145, 3), dtype=torch.float64)
    img = torch.tensor(rng.rand(1, 64, 64, 3) * 2 - 1)
    emb, rot = O2.real_encoder_forward(p, img)
    assert emb.shape == (1, 145) and rot.shape == (1, 3)
    lim = O2.rotation_range_multiplier()
    assert np.all(np.abs(rot.numpy()) <= lim + 1e-12) and float(rot[0, 2]) == 0.0
    spec = netspec.real_encoder_spec(145)
    n_conv = sum(1 for k in spec if k.endswith("_conv/kernel"))
    n_bn = sum(1 for k in spec if k.endswith("_bn/gamma"))
    assert n_conv == 53 and n_bn == 53            # ResNet50: 1 + 16*3 + 4 shortcut convs, one BN each
    assert sum(int(np.prod(s)) for s, _ in spec.values()) == 23587712 + 2048 * 3 + 3 + 2048 * 145 + 145


def test_latent_gan_oracle_shapes_and_r1():
    spec_g = netspec.latent_gan_mlp_spec(145)
    spec_d = netspec.latent_gan_mlp_spec(145, num_out=1)
    assert spec_g["mlp/dense0/kernel"][0] == (145, 217) and spec_d["mlp/dense2/kernel"][0] == (217, 1)
    p_g = O.to_torch(netspec.init_params(spec_g, 1), dtype=torch.float64, requires_grad=True)
    p_d = O.to_torch(netspec.init_params(spec_d, 2), dtype=torch.float64, requires_grad=True)
    rng = np.random.RandomState(0)
    l = O2.latent_gan_discriminator_losses(p_d, p_g, torch.tensor(rng.randn(4, 145)), torch.tensor(rng.randn(4, 145)))
    assert list(l.keys()) == ["GAN_loss_real", "GAN_loss_fake", "gp_loss", "loss_sum"]
    assert float(l["gp_loss"]) > 0
    g = O.grads_of(l["loss_sum"], p_d)
    assert all(torch.isfinite(x).all() for x in g)


def test_param_group_trainable_filter_and_keras_order():
    """keras lists BatchNorm moving statistics in get_weights() but not in trainable_weights (real_encoder.py:13)."""
    names = list(netspec.real_encoder_spec(145).keys())
    assert names[:6] == ["resnet/conv1_conv/kernel", "resnet/conv1_conv/bias", "resnet/conv1_bn/gamma", "resnet/conv1_bn/beta",
                         "resnet/conv1_bn/moving_mean", "resnet/conv1_bn/moving_variance"]
    assert names[-4:] == ["rotation_regressor/kernel", "rotation_regressor/bias",
                          "feature_to_latent_mlp/kernel", "feature_to_latent_mlp/bias"]
    from confignet_b200.runtime import ParamGroup
    arrays = OrderedDict([("a/gamma", np.ones(3, np.float32)), ("a/moving_mean", np.zeros(3, np.float32)),
                          ("b/kernel", np.ones((2, 2), np.float32))])
    g = ParamGroup(arrays, "cpu", trainable=netspec.is_trainable)
    assert len(g.trainable_weights) == 2 and len(g.get_weights()) == 3
    assert not g.params["a/moving_mean"].requires_grad and g.params["a/gamma"].requires_grad
    with pytest.raises(ValueError):
        g.pack_grads([None, None, None])


# ------------------------------------------------------------------------------------------------ data parallel (gloo, 2 ranks)
_DP_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from confignet_b200.runtime import gather_rows, shard_rows, world
from oracle import confignet_oracle_stage2 as O2
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, ws = world()
rng = np.random.RandomState(0)
out, lab = rng.randn(8, 12), rng.randn(8, 12)
lo, hi = shard_rows(8)
o = torch.tensor(out[lo:hi], requires_grad=True); l = torch.tensor(lab[lo:hi], requires_grad=True)
loss = O2.normalized_regression(gather_rows(o), gather_rows(l), 10.0)
go, gl = torch.autograd.grad(loss, (o, l))
fo = torch.tensor(out, requires_grad=True); fl = torch.tensor(lab, requires_grad=True)
full = O2.normalized_regression(fo, fl, 10.0)
wo, wl = torch.autograd.grad(full, (fo, fl))
# the local gradient is world x the global one on this rank's rows: allreduce_grads()'s 1/world undoes it
e = max(float((go / ws - wo[lo:hi]).abs().max()), float((gl / ws - wl[lo:hi]).abs().max()), abs(float(loss) - float(full)))
assert e < 1e-12, e
dist.destroy_process_group()
print("rank", rank, "ok", e)
'''


def test_batch_statistics_loss_under_data_parallelism(tmp_path):
    script = tmp_path / "dp_worker2.py"
    script.write_text(_DP_WORKER)
    port = str(31500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_stage2_step_host_halves_consume_the_reference_stream():
    """Host halves of the stage-2 steps on a CPU model with the device halves recorded instead of run.  ConfigNet's
    discriminator step must use ConfigNet's batch assembly (confignet_second_stage.py:119-130: image rows, flips, INPUT
    image rows for the encoder -> generator reconstructions), not stage 1's prior samples (an inherited stage-1 step
    would draw latents and rotations instead and leave the NumPy stream somewhere else)."""
    from confignet_b200.confignet_second_stage import ConfigNet
    from confignet_b200.confignet_first_stage import ConfigNetFirstStage
    from confignet_b200.runtime import KerasAdam
    from confignet_b200.synthetic_data import SyntheticDataset
    assert ConfigNet.discriminator_training_step is not ConfigNetFirstStage.discriminator_training_step
    B, res = 4, 16
    m = ConfigNet({"output_shape": (res, res, 3), "batch_size": B, "facemodel_inputs": netspec.default_facemodel_inputs()},
                  initialize=False, device="cpu")
    rec = []
    m._graphed = lambda name, opt, fn, nets, **kw: (lambda *t: (rec.append((name, [x.clone() for x in t], kw)), OrderedDict(loss_sum=torch.zeros(())))[1])
    real, synth = SyntheticDataset(9, res, seed=1), SyntheticDataset(7, res, seed=2)
    opt = KerasAdam()
    m.discriminator = m.latent_discriminator = m.generator = m.latent_regressor = m.synthetic_encoder = m.encoder = None

    np.random.seed(3)
    m.discriminator_training_step(real, opt)
    after = np.random.randint(0, 1 << 30)
    np.random.seed(3)
    idx = np.random.randint(0, 9, B); flips = np.random.randint(0, 2, size=B); in_idx = np.random.randint(0, 9, B)
    assert np.random.randint(0, 1 << 30) == after                      # nothing else was drawn
    name, t, kw = rec[-1]
    assert name == "d2" and opt.iterations == 1
    assert np.array_equal(t[0].numpy(), real.imgs[idx]) and np.array_equal(t[1].numpy(), flips.astype(bool))
    assert np.array_equal(t[2].numpy(), real.imgs[in_idx])

    np.random.seed(4)
    m.latent_discriminator_training_step(real, synth, opt)
    np.random.seed(4)
    idx = np.random.randint(0, 9, B); flips = np.random.randint(0, 2, size=B); sidx = np.random.randint(0, 7, B)
    name, t, kw = rec[-1]
    assert name == "latent_d2" and opt.iterations == 2
    assert np.array_equal(t[0].numpy(), real.imgs[idx]) and np.array_equal(t[1].numpy(), flips.astype(bool))
    for x, n in zip(t[2:], netspec.default_facemodel_inputs().keys()):
        assert np.array_equal(x.numpy(), synth.metadata_inputs[n][sidx])

    np.random.seed(5)
    m.generator_training_step(real, synth, opt)
    np.random.seed(5)
    sidx = np.random.randint(0, 7, B // 2); ridx = np.random.randint(0, 9, B - B // 2); rflips = np.random.randint(0, 2, size=B - B // 2)
    name, t, kw = rec[-1]
    assert name == "g2" and kw.get("dp_ok") is False                   # the all-gathered batch-statistics loss is not capturable under DP
    assert np.array_equal(t[0].numpy(), synth.imgs[sidx]) and np.array_equal(t[1].numpy(), synth.eye_masks[sidx].astype(np.float32))
    assert np.array_equal(t[2].numpy(), real.imgs[ridx]) and np.array_equal(t[3].numpy(), rflips.astype(bool))
    assert np.array_equal(t[4].numpy(), synth.metadata_inputs["rotations"][sidx])


def test_pretrained_standins_warn_and_npz_exports_load():
    """The VGG19 / VGGFace / ResNet50 weights the reference downloads (perceptual_loss.py:19-41, real_encoder.py:13) are
    seeded stand-ins here: the first training / fine-tuning call says so once, and an .npz export in Keras'
    get_weights() order (or keyed by variable name) replaces them - truncated-away trailing layers ignored, shapes checked."""
    import warnings
    import tempfile
    from confignet_b200 import pretrained
    from confignet_b200.runtime import ParamGroup, Network

    class M:
        PRETRAINED = (("perceptual_loss", "VGG19"), ("encoder", "ResNet50"))
        config = {}
    m = M()
    m.perceptual_loss = Network(ParamGroup(netspec.init_params(netspec.vgg19_spec(), 1, vgg_like=True), "cpu"), lambda *a: None)
    m.encoder = Network(ParamGroup(netspec.init_real_encoder_params(145, 2), "cpu", trainable=netspec.is_trainable), lambda *a: None)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        pretrained.warn_if_standin(m, M.PRETRAINED, "generator_training_step")
        pretrained.warn_if_standin(m, M.PRETRAINED, "generator_training_step")
    assert len(w) == 2 and "stand-in" in str(w[0].message)
    rng = np.random.RandomState(0)
    g = m.perceptual_loss.group
    want = [rng.randn(*s).astype(np.float32) for s in g.shapes]
    extra = [rng.randn(3, 3, 512, 512).astype(np.float32), rng.randn(512).astype(np.float32)]      # block4_conv3...: not in the spec
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "vgg19.npz")
        np.savez(path, *(want + extra))
        pretrained.load_vgg(g, path, "VGG19")
        for a, b in zip(g.get_weights(), want):
            assert np.array_equal(a, b)
        assert g.pretrained
        bad = list(want); bad[0] = bad[0][:2]
        np.savez(path, *bad)
        with pytest.raises(ValueError):
            pretrained.load_vgg(g, path, "VGG19")
    e = m.encoder.group
    res = {n[len("resnet/"):]: rng.randn(*s).astype(np.float32) for n, s in zip(e.names, e.shapes) if n.startswith("resnet/")}
    heads = {n: a for n, a in zip(e.names, e.get_weights()) if not n.startswith("resnet/")}
    pretrained.load_resnet50(e, res)
    got = dict(zip(e.names, e.get_weights()))
    assert all(np.array_equal(got["resnet/" + k], v) for k, v in res.items())
    assert all(np.array_equal(got[k], v) for k, v in heads.items())
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m2 = M(); m2.perceptual_loss, m2.encoder = m.perceptual_loss, m.encoder
        pretrained.warn_if_standin(m2, M.PRETRAINED, "fine_tune_on_img")
    assert not w
