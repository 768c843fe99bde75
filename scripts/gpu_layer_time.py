"""Times single conv-family launches (CUDA events, median of reps) for the layer shapes of the bench step.
Usage: python scripts/gpu_layer_time.py [dbg-mask ...]   (dbg masks are the cn_debug_set test hook bits)
Not a bench number: used to find which part of a kernel (gather / STS / MMA) bounds it."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import os
os.environ.setdefault("CN_TEST_HOOKS", "1")      # the cn_debug_* hooks live in libconfignet_b200_hooks.so only
from confignet_b200 import _lib as L

lib = L.load()
dev = torch.device("cuda:0")
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None

LAYERS = [  # (nd, batch, dims, cin, cout, k, stride, up)
    (2, 16, (64, 64), 256, 256, 3, 1, 1),
    (2, 16, (256, 256), 64, 64, 3, 1, 1),
    (2, 32, (128, 128), 48, 96, 3, 2, 1),
    (2, 32, (64, 64), 96, 192, 3, 2, 1),
    (3, 16, (8, 8, 8), 256, 128, 3, 1, 2),
    (3, 16, (16, 16, 16), 128, 64, 3, 1, 1),
    (2, 16, (64, 64), 32, 32, 4, 1, 2),
    (2, 16, (32, 32), 512, 512, 3, 1, 1),
]


def time_it(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    masks = [int(a) for a in sys.argv[1:]] or [0]
    if os.environ.get("CLUSTER"):
        lib.cn_debug_set_cluster(int(os.environ["CLUSTER"]))
    only = os.environ.get("OPS", "fwd,dgrad,wgrad").split(",")
    print("%-50s %-6s " % ("layer", "op") + " ".join("dbg=%-2d ms (TF/s)   " % m for m in masks))
    for (nd, B, dims, cin, cout, k, s, up) in LAYERS:
        d = L.make_conv_desc(nd, B, dims, cin, cout, [k] * nd, s, up)
        x = torch.randn(B, *dims, cin, device=dev)
        w = torch.randn(*([k] * nd), cin, cout, device=dev) * 0.05
        od = (ctypes.c_int * 3)(); L.call("cn_conv_out_dims", ctypes.byref(d), od)
        oshape = (B,) + tuple(od[:nd]) + (cout,)
        gy = torch.randn(*oshape, device=dev)
        y = torch.empty(oshape, device=dev); gx = torch.empty_like(x); gw = torch.empty_like(w)
        flops = 2.0 * np.prod(oshape) * cin * k ** nd
        fns = {"fwd": lambda: L.call("cn_conv_fwd", ctypes.byref(d), P(x), P(w), None, 0, 0.0, P(y), 0, st()),
               "dgrad": lambda: L.call("cn_conv_dgrad", ctypes.byref(d), P(gy), P(w), P(gx), 0, st()),
               "wgrad": lambda: L.call("cn_conv_wgrad", ctypes.byref(d), P(x), P(gy), P(gw), None, 0, st())}
        for op in only:
            res = []
            for m in masks:
                lib.cn_debug_set(m)
                t = time_it(fns[op])
                res.append("%7.3f (%6.1f)      " % (t, flops / t / 1e9))
            lib.cn_debug_set(0)
            print("%-50s %-6s " % (str((nd, B, dims, cin, cout, k, s, up)), op) + " ".join(res), flush=True)


if __name__ == "__main__":
    main()
