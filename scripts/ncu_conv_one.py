"""Runs the three tensor-core conv kernels once each on VGG block3-like and generator 3-D shapes (for ncu)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from confignet_b200 import _lib as L
lib = L.load()
dev = torch.device("cuda:0")
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr())
SHAPES = {"default": [(2, 16, (64, 64), 256, 256, 3, 1, 1), (3, 16, (8, 8, 8), 256, 128, 3, 1, 2)],
          # the discriminator's 48 / 96-channel stride-2 blocks and VGG block1_conv2: the layers furthest below their ceilings
          "dblocks": [(2, 32, (128, 128), 48, 96, 3, 2, 1), (2, 16, (256, 256), 64, 64, 3, 1, 1)]}
for (nd, B, dims, cin, cout, k, s, up) in SHAPES[os.environ.get("CN_NCU_SHAPES", "default")]:
    d = L.make_conv_desc(nd, B, dims, cin, cout, [k] * nd, s, up)
    od = (ctypes.c_int * 3)(); L.call("cn_conv_out_dims", ctypes.byref(d), od)
    x = torch.randn(B, *dims, cin, device=dev); w = torch.randn(*([k] * nd), cin, cout, device=dev) * 0.02
    y = torch.empty((B,) + tuple(od[:nd]) + (cout,), device=dev); gy = torch.randn_like(y)
    gx = torch.empty_like(x); gw = torch.empty_like(w)
    for it in range(int(os.environ.get('CN_NCU_REPS', '2'))):
        L.call("cn_conv_fwd", ctypes.byref(d), P(x), P(w), None, 0, 0.0, P(y), L.IMPL_TC, st())
        L.call("cn_conv_dgrad", ctypes.byref(d), P(gy), P(w), P(gx), L.IMPL_TC, st())
        L.call("cn_conv_wgrad", ctypes.byref(d), P(x), P(gy), P(gw), None, L.IMPL_TC, st())
    torch.cuda.synchronize()
print("done")
