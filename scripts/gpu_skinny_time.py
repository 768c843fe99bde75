"""Times the HBM-bound skinny conv layers (fwd / dgrad / wgrad) at bench sizes and prints achieved GB/s of
algorithmic traffic (wide tensor + narrow tensor, each moved once) against the measured HBM peak."""
import sys, os, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from confignet_b200 import _lib as L
lib = L.load()
dev = torch.device("cuda:0")
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
peak = 6546.9
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=10):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


for name, (B, dims, cin, cout, k, s, up, act) in {
        "D.block0 3->48 k3 s2 @256 B32": (32, (256, 256), 3, 48, 3, 2, 1, 0),
        "VGG b1c1 3->64 k3 s1 @256 B16": (16, (256, 256), 3, 64, 3, 1, 1, 2),
        "map_final up2 32->3 k4 @128 B16": (16, (128, 128), 32, 3, 4, 1, 2, 3)}.items():
    d = L.make_conv_desc(2, B, dims, cin, cout, [k, k], s, up)
    od = (ctypes.c_int * 3)(); L.call("cn_conv_out_dims", ctypes.byref(d), od)
    x = torch.randn(B, *dims, cin, device=dev); w = torch.randn(k, k, cin, cout, device=dev) * 0.05
    b = torch.randn(cout, device=dev)
    y = torch.empty(B, od[0], od[1], cout, device=dev); gy = torch.randn_like(y)
    gx = torch.empty_like(x); gw = torch.empty_like(w); gb = torch.empty_like(b)
    nbytes = 4 * (x.numel() + y.numel())
    for op, fn in (("fwd", lambda: L.call("cn_conv_fwd", ctypes.byref(d), P(x), P(w), P(b), act, 0.0, P(y), L.IMPL_AUTO, st())),
                   ("dgrad", lambda: L.call("cn_conv_dgrad", ctypes.byref(d), P(gy), P(w), P(gx), L.IMPL_AUTO, st())),
                   ("wgrad", lambda: L.call("cn_conv_wgrad", ctypes.byref(d), P(x), P(gy), P(gw), None, L.IMPL_AUTO, st()))):
        ms = timeit(fn)
        print("%-34s %-5s %8.1f us  %7.0f GB/s  %.2f of HBM peak" % (name, op, ms * 1e3, nbytes / ms / 1e6, nbytes / ms / 1e6 / peak))
