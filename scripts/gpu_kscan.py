"""Fixed vs per-k-block cost of the tcgen05 pixel kernel: forward 3x3 conv on 16x64x64 pixels, Cout=128, Cin swept."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from confignet_b200 import _lib as L
lib = L.load(); dev = torch.device("cuda:0")
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
if os.environ.get("CLUSTER"): lib.cn_debug_set_cluster(int(os.environ["CLUSTER"]))
masks = [int(a) for a in sys.argv[1:]] or [0]
cout = int(os.environ.get("COUT", "128"))
for cin in (32, 64, 128, 256, 512, 1024):
    B, dims, k = 16, (64, 64), 3
    d = L.make_conv_desc(2, B, dims, cin, cout, [k, k], 1, 1)
    x = torch.randn(B, *dims, cin, device=dev); w = torch.randn(k, k, cin, cout, device=dev) * 0.05
    y = torch.empty(B, *dims, cout, device=dev)
    fn = lambda: L.call("cn_conv_fwd", ctypes.byref(d), P(x), P(w), None, 0, 0.0, P(y), 0, st())
    out = []
    for m in masks:
        lib.cn_debug_set(m); fn(); torch.cuda.synchronize(); ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        out.append("dbg=%d %.3f ms" % (m, float(np.median(ts))))
    lib.cn_debug_set(0)
    print("cin=%4d kb=%4d  " % (cin, cin * 9 // 32) + "  ".join(out), flush=True)
