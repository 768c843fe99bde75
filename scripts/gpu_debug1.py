import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from confignet_b200 import netspec, networks, ops, _lib as L
from oracle import confignet_oracle as O
from parity_utils import nerr, make_params, grads_cpu, grads_gpu
dev = torch.device("cuda:0")
torch.set_num_threads(os.cpu_count())
# ---- 1. precision vs K of the tensor-core kernel
print("== precision vs K (fwd conv 2-D 3x3, B=2, 16x16, cout=64)")
for cin in (32, 128, 512, 2048):
    for positive in (False, True):
        torch.manual_seed(0)
        x = torch.randn(2, 16, 16, cin); w = torch.randn(3, 3, cin, 64) / np.sqrt(9 * cin)
        if positive: x, w = x.abs(), w.abs()
        yr = O.conv_same(x.double(), w.double())
        res = []
        for impl in (L.IMPL_FFMA, L.IMPL_TC):
            ops.IMPL[0] = impl
            y = ops.conv_act(x.to(dev), w.to(dev))
            res.append(nerr(y, yr))
        ops.IMPL[0] = L.IMPL_AUTO
        print("K=%6d positive=%d  ffma %.2e  tc %.2e" % (9 * cin, positive, res[0], res[1]))
# ---- 2. generator gradients per tensor
B = 2
p_cpu, grp = make_params(netspec.generator_spec(145, 256), 11, dev, dtype=torch.float64)
rng = np.random.RandomState(0)
z = rng.randn(B, 145).astype(np.float32); rot = np.array([[0.3, -0.1, 0], [-0.2, 0.05, 0]], np.float32)
out_r = O.generator_forward(p_cpu, torch.tensor(z).double(), torch.tensor(rot).double(), 256)
go = rng.randn(*out_r.shape).astype(np.float32)
g_r = grads_cpu((out_r * torch.tensor(go).double()).sum(), p_cpu)
# fp32 CPU oracle as a yardstick of what plain fp32 arithmetic gives
p32 = O.to_torch({k: v.detach().numpy() for k, v in p_cpu.items()}, dtype=torch.float32, requires_grad=True)
out32 = O.generator_forward(p32, torch.tensor(z), torch.tensor(rot), 256)
g32 = grads_cpu((out32 * torch.tensor(go)).sum(), p32)
for impl, name in ((L.IMPL_FFMA, "ffma"), (L.IMPL_AUTO, "auto")):
    ops.IMPL[0] = impl
    out = networks.generator_forward(grp.params, torch.tensor(z, device=dev), rot, 256)
    g = grads_gpu((out * torch.tensor(go, device=dev)).sum(), grp)
    ops.IMPL[0] = L.IMPL_AUTO
    print("== generator", name, "fwd err", nerr(out, out_r), " (cpu fp32 oracle vs fp64: %.2e)" % nerr(out32, out_r))
    for k in g_r:
        e = nerr(g[k], g_r[k]); e32 = nerr(g32[k], g_r[k])
        if e > 2e-4 or e32 > 2e-4:
            print("   %-34s err %.2e   cpu-fp32 err %.2e   max|g| %.2e" % (k, e, e32, float(g_r[k].abs().max())))
