"""Times the HBM-bound kernels of the step at the shapes the bench step launches them with: CUDA events on the launch
stream, an L2 flush (256 MB write) before every timed launch, median of `reps`.  Prints us and algorithmic GB/s
(4 B x elements x tensors read + written) against MEASURED_PEAKS.json's HBM rate.

    python scripts/gpu_hbm_time.py [reps]
"""
import sys, os, ctypes, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from confignet_b200 import _lib as L, ops

lib = L.load()
dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 7
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6546.9
flush_buf = torch.empty(64 * 1024 * 1024, device=dev)


def timed(fn):
    ts = []
    for _ in range(reps):
        flush_buf.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def line(name, shape, us, nbytes):
    gbs = nbytes / us / 1e3
    print("%-34s %-24s %9.1f us %8.0f GB/s  %.2f of HBM" % (name, shape, us, gbs, gbs / PEAK), flush=True)


# (n, pixels-per-sample dims, channels): D blocks 0..4 at batch 32, generator AdaIN sites at batch 16
SHAPES = [(32, (128, 128), 48), (32, (64, 64), 96), (32, (32, 32), 192), (32, (16, 16), 384), (32, (8, 8), 768),
          (16, (8, 8, 8), 256), (16, (16, 16, 16), 128), (16, (16, 16), 256), (16, (64, 64), 64), (16, (128, 128), 32),
          (16, (256, 256), 32)]
torch.manual_seed(0)
for (n, dims, ch) in SHAPES:
    a = torch.randn(n, *dims, ch, device=dev)
    b = torch.randn_like(a); c = torch.randn_like(a)
    el = a.numel() * 4
    sh = "%dx%sx%d" % (n, "x".join(map(str, dims)), ch)
    p = a.numel() // (n * ch)
    line("chan_sums(a)", sh, timed(lambda: ops._sums(a, flags=1, alpha=0.3)), el)
    line("chan_sums(a,b)", sh, timed(lambda: ops._sums(a, b, flags=1, alpha=0.3)), 2 * el)
    line("chan_sums(a,b,c)", sh, timed(lambda: ops._sums(a, b, c, flags=5, alpha=0.3)), 3 * el)
    s = ops._sums(a, b, c, flags=5, alpha=0.3)
    gam = torch.ones(ch, device=dev); sb = torch.randn(n, 2 * ch, device=dev)
    us = timed(lambda: ops._coef(ops.COEF_IN_BWD, s, gam, None, n, ch, p, 1e-3, 1, (ch,), (ch,)))
    print("%-34s %-24s %9.1f us   (%d slices)" % ("norm_coef IN_BWD", sh, us, s.shape[0]), flush=True)
    us = timed(lambda: ops._coef(ops.COEF_ADAIN_FWD, s, sb, None, n, ch, p, 1e-3))
    print("%-34s %-24s %9.1f us" % ("norm_coef ADAIN_FWD", sh, us), flush=True)
    (coef,), _, _ = ops._coef(ops.COEF_ADAIN_FWD, s, sb, None, n, ch, p, 1e-3)
    line("chan_affine(a)", sh, timed(lambda: ops._affine(a, None, None, coef, 1, 0.3)), 2 * el)
    line("chan_affine(a,b)", sh, timed(lambda: ops._affine(a, b, None, coef, 3, 0.3)), 3 * el)
    line("chan_affine(a,b,c)", sh, timed(lambda: ops._affine(a, b, c, coef, 7, 0.3)), 4 * el)
    y = torch.empty_like(a)
    line("act_bwd", sh, timed(lambda: L.call("cn_act_bwd", ops._p(b), ops._p(a), L.ACT_LRELU, 0.3, ops._p(y), a.numel(), ops._stream())), 3 * el)
    line("torch add (autograd accumulation)", sh, timed(lambda: torch.add(a, b, out=y)), 3 * el)
    line("reduce sqdiff", sh, timed(lambda: ops.reduce_sum(a, ops.RED_SQDIFF, y=b)), 2 * el)
    del a, b, c, y, s

# bias gradient / weight gradient of the skinny and folded layers: cn_conv_wgrad with a bias pointer
for (B, dims, cin, cout, k, s_, up) in [(32, (256, 256), 3, 48, 3, 2, 1), (16, (256, 256), 3, 64, 3, 1, 1), (16, (128, 128), 32, 3, 4, 1, 2),
                                        (16, (8, 8, 8), 256, 128, 3, 1, 2), (16, (64, 64), 32, 32, 4, 1, 2), (32, (256, 256), 3, 3, 1, 1, 1)]:
    nd = len(dims)
    d = L.make_conv_desc(nd, B, dims, cin, cout, [k] * nd, s_, up)
    od = (ctypes.c_int * 3)(); L.call("cn_conv_out_dims", ctypes.byref(d), od)
    x = torch.randn(B, *dims, cin, device=dev); w = torch.randn(*([k] * nd), cin, cout, device=dev) * 0.05
    bias = torch.randn(cout, device=dev)
    yy = torch.empty(B, *[od[i] for i in range(nd)], cout, device=dev); gy = torch.randn_like(yy)
    gx = torch.empty_like(x); gw = torch.empty_like(w); gb = torch.empty(cout, device=dev)
    sh = "%dx%sx%d->%d k%d s%d u%d" % (B, "x".join(map(str, dims)), cin, cout, k, s_, up)
    io = (x.numel() + yy.numel()) * 4
    st = ops._stream
    line("conv fwd", sh, timed(lambda: L.call("cn_conv_fwd", ctypes.byref(d), ops._p(x), ops._p(w), ops._p(bias), 0, 0.0, ops._p(yy), L.IMPL_AUTO, st())), io)
    line("conv dgrad", sh, timed(lambda: L.call("cn_conv_dgrad", ctypes.byref(d), ops._p(gy), ops._p(w), ops._p(gx), L.IMPL_AUTO, st())), io)
    line("conv wgrad", sh, timed(lambda: L.call("cn_conv_wgrad", ctypes.byref(d), ops._p(x), ops._p(gy), ops._p(gw), None, L.IMPL_AUTO, st())), io)
    line("conv wgrad + bias", sh, timed(lambda: L.call("cn_conv_wgrad", ctypes.byref(d), ops._p(x), ops._p(gy), ops._p(gw), ops._p(gb), L.IMPL_AUTO, st())), io)
print("done")
