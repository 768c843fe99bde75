import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from confignet_b200 import ops, _lib as L
from oracle import confignet_oracle as O
from parity_utils import nerr
dev = torch.device("cuda:0")
print("== precision vs K (fwd conv 2-D 3x3, B=2, 16x16, cout=64): normalised max err and rms err / rms ref")
def rerr(a, b):
    a = a.double().cpu(); b = b.double().cpu()
    return float(((a - b) ** 2).mean().sqrt() / (b ** 2).mean().sqrt())
for cin in (32, 128, 512, 2048):
    for positive in (False, True):
        torch.manual_seed(0)
        x = torch.randn(2, 16, 16, cin); w = torch.randn(3, 3, cin, 64) / np.sqrt(9 * cin)
        if positive: x, w = x.abs(), w.abs()
        yr = O.conv_same(x.double(), w.double())
        y32 = O.conv_same(x, w)
        res = []
        for impl in (L.IMPL_FFMA, L.IMPL_TC):
            ops.IMPL[0] = impl
            y = ops.conv_act(x.to(dev), w.to(dev))
            res += [nerr(y, yr), rerr(y, yr)]
        ops.IMPL[0] = L.IMPL_AUTO
        print("K=%6d positive=%d  ffma max %.2e rms %.2e | tc max %.2e rms %.2e | cpu-fp32 max %.2e rms %.2e" % (9 * cin, positive, res[0], res[1], res[2], res[3], nerr(y32, yr), rerr(y32, yr)))
