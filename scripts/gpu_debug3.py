import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from confignet_b200 import netspec, networks, ops, _lib as L
from oracle import confignet_oracle as O
from parity_utils import nerr, make_params, grads_cpu, grads_gpu
dev = torch.device("cuda:0")
def l2(a, b):
    a = a.double().cpu(); b = b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
def _rot(n, seed):
    rng = np.random.RandomState(seed)
    r = np.zeros((n, 3), np.float32)
    r[:, 0] = np.pi * rng.uniform(-30, 30, n) / 180
    r[:, 1] = np.pi * rng.uniform(-10, 10, n) / 180
    return r
B = 2
p_cpu, grp = make_params(netspec.generator_spec(145, 256), 11, dev, dtype=torch.float64)
rng = np.random.RandomState(0)
z = rng.randn(B, 145).astype(np.float32); rot = _rot(B, 1)
out_r = O.generator_forward(p_cpu, torch.tensor(z).double(), torch.tensor(rot).double(), 256)
go = rng.randn(*out_r.shape).astype(np.float32)
g_r = grads_cpu((out_r * torch.tensor(go).double()).sum(), p_cpu)
for rep in range(2):
  for impl, name in ((L.IMPL_FFMA, "ffma"), (L.IMPL_AUTO, "auto")):
    ops.IMPL[0] = impl
    out = networks.generator_forward(grp.params, torch.tensor(z, device=dev), rot, 256)
    g = grads_gpu((out * torch.tensor(go, device=dev)).sum(), grp)
    ops.IMPL[0] = L.IMPL_AUTO
    print("== generator", name, "rep", rep, "fwd max-err %.2e l2 %.2e" % (nerr(out, out_r), l2(out, out_r)))
    for k in g_r:
        e = nerr(g[k], g_r[k])
        if e > 1e-3:
            print("   %-34s max-err %.2e  l2 %.2e  max|g| %.2e" % (k, e, l2(g[k], g_r[k]), float(g_r[k].abs().max())))
