"""CPU only: how far is the fp32 CPU oracle from the fp64 CPU oracle on the stage-2 generator-step parameter
gradients (the inputs of tests/test_stage2_gpu.py::test_stage2_generator_step_losses_and_grads)?  The parity
tolerance on whole-network gradients is set from this (profiles/r01_precision_study.md)."""
import os, sys
from collections import OrderedDict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from confignet_b200 import netspec
from oracle import confignet_oracle as O
from oracle import confignet_oracle_stage2 as O2
FM = netspec.default_facemodel_inputs()


def rot(n, seed):
    rng = np.random.RandomState(seed)
    r = np.zeros((n, 3), np.float32)
    r[:, 0] = np.pi * rng.uniform(-30, 30, n) / 180
    r[:, 1] = np.pi * rng.uniform(-10, 10, n) / 180
    return r


def run(dtype):
    def mk(spec, seed, perturb=0.05, vgg_like=False):
        arrays = netspec.init_params(spec, seed, vgg_like=vgg_like)
        if perturb:
            arrays = netspec.perturb_params(arrays, seed + 1000, perturb)
        return O.to_torch(arrays, dtype=dtype, requires_grad=True)
    p_g = mk(netspec.generator_spec(145, 256), 41); p_lr = mk(netspec.latent_regressor_spec(145, 256), 42)
    p_se = mk(netspec.synthetic_encoder_spec(FM, 2), 43); p_d = mk(netspec.discriminator_spec(256), 44)
    p_sd = mk(netspec.discriminator_spec(256), 45); p_ld = mk(netspec.latent_discriminator_spec(145, 4), 46)
    p_v = mk(netspec.vgg19_spec(), 47, perturb=0.0, vgg_like=True)
    p_e = O.to_torch(netspec.init_real_encoder_params(145, 48), dtype=dtype, requires_grad=True)
    for k, v in p_e.items():
        if not netspec.is_trainable(k):
            v.requires_grad_(False)
    rng = np.random.RandomState(4)
    ns = nr = 2
    fparams = [rng.rand(ns, d[0]).astype(np.float32) for d in FM.values()]
    simgs = (rng.rand(ns, 256, 256, 3).astype(np.float32) * 2 - 1)
    rimgs = (rng.rand(nr, 256, 256, 3).astype(np.float32) * 2 - 1)
    masks = (rng.rand(ns, 256, 256) < 0.01).astype(np.uint8)
    srot = rot(ns, 5)
    W = dict(O.DEFAULT_LOSS_WEIGHTS); W["image_loss_weight"] = 5e-4
    batch = dict(facemodel_params=[torch.tensor(a).to(dtype) for a in fparams], synth_rotations=torch.tensor(srot).to(dtype),
                 synth_imgs=torch.tensor(simgs).to(dtype), eye_masks=masks, real_imgs=torch.tensor(rimgs).to(dtype))
    l = O2.stage2_generator_step_losses(p_g, p_lr, p_se, p_e, p_d, p_sd, p_ld, p_v, FM, batch, weights=W)
    allp = OrderedDict()
    for pre, p in (("g/", p_g), ("lr/", p_lr), ("se/", p_se), ("enc/", p_e)):
        for k, v in p.items():
            if v.requires_grad:
                allp[pre + k] = v
    gs = torch.autograd.grad(l["loss_sum"], list(allp.values()), allow_unused=True)
    return OrderedDict((k, (torch.zeros_like(v) if g is None else g).double()) for (k, v), g in zip(allp.items(), gs))


torch.set_num_threads(os.cpu_count())
g64 = run(torch.float64)
g32 = run(torch.float32)
top = max(float(v.norm()) for v in g64.values())
rows = []
for k in g64:
    if float(g64[k].norm()) < 1e-3 * top:
        continue
    rows.append((float((g32[k] - g64[k]).norm() / (g64[k].norm() + 1e-30)), k))
rows.sort(reverse=True)
print("fp32 CPU oracle vs fp64 CPU oracle, relative L2 per parameter gradient (worst 12 of %d):" % len(rows))
for e, k in [r for r in rows if "rotation" in r[1] or "enc/" in r[1]][:6] + rows[:12]:
    print("  %.3e  %s" % (e, k))
