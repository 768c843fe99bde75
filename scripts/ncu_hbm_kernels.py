"""Runs the HBM-bound kernels of the step once each at bench sizes (for `ncu --set full`): per-channel statistics,
coefficient and affine kernels of the InstanceNorm / AdaIN / style family (forward and backward), the skinny conv
layers, fused Adam."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from confignet_b200 import _lib as L, ops
lib = L.load()
dev = torch.device("cuda:0")
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
torch.manual_seed(0)
for it in range(2):
    # discriminator block 0 output: InstanceNorm(LeakyReLU(c)) forward + backward, layer style
    c = torch.randn(32, 128, 128, 48, device=dev, requires_grad=True)
    gam = torch.ones(48, device=dev, requires_grad=True); bet = torch.zeros(48, device=dev, requires_grad=True)
    y = ops.lrelu_instance_norm(c, gam, bet, 0.3)
    s = ops.layer_style(c)
    (y.sum() + s.sum()).backward()
    # generator AdaIN at 128x128x32
    a = torch.randn(16, 128, 128, 32, device=dev, requires_grad=True)
    sb = torch.randn(16, 64, device=dev, requires_grad=True)
    ops.adain(a, sb, mask_alpha=0.3).sum().backward()
    # skinny convs
    for (B, dims, cin, cout, k, s_, up, act) in [(32, (256, 256), 3, 48, 3, 2, 1, 0), (16, (128, 128), 32, 3, 4, 1, 2, 3)]:
        d = L.make_conv_desc(2, B, dims, cin, cout, [k, k], s_, up)
        od = (ctypes.c_int * 3)(); L.call("cn_conv_out_dims", ctypes.byref(d), od)
        x = torch.randn(B, *dims, cin, device=dev); w = torch.randn(k, k, cin, cout, device=dev) * 0.05
        b = torch.randn(cout, device=dev)
        yy = torch.empty(B, od[0], od[1], cout, device=dev); gy = torch.randn_like(yy)
        gx = torch.empty_like(x); gw = torch.empty_like(w)
        L.call("cn_conv_fwd", ctypes.byref(d), P(x), P(w), P(b), act, 0.0, P(yy), L.IMPL_AUTO, st())
        L.call("cn_conv_dgrad", ctypes.byref(d), P(gy), P(w), P(gx), L.IMPL_AUTO, st())
        L.call("cn_conv_wgrad", ctypes.byref(d), P(x), P(gy), P(gw), None, L.IMPL_AUTO, st())
    # fused Adam on a generator-sized flat buffer
    n = 8 * 1024 * 1024
    p_, g_, m_, v_ = [torch.randn(n, device=dev) for _ in range(4)]
    v_.abs_()
    ops.adam_ema_step(p_, g_, m_, v_, None, 4e-4, 0.0, 0.9, 1e-7, 0.0, 1.0)
    torch.cuda.synchronize()
print("done")
