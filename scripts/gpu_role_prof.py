"""In-kernel role timings (clock64) of CTA 0 of the tcgen05 pixel kernel for one layer (test hook cn_debug_set_prof)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from confignet_b200 import _lib as L
lib = L.load(); dev = torch.device("cuda:0")
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
lib.cn_debug_set_cluster(int(os.environ.get("CLUSTER", "1")))
lib.cn_debug_set_prof.argtypes = [ctypes.c_void_p]
buf = torch.zeros(16, dtype=torch.int64, device=dev)
OP = os.environ.get("OP", "fwd")
for (B, dims, cin, cout, k) in [(16, (64, 64), 256, 256, 3), (16, (64, 64), 256, 64, 3), (16, (256, 256), 64, 64, 3)]:
    d = L.make_conv_desc(2, B, dims, cin, cout, [k, k], 1, 1)
    x = torch.randn(B, *dims, cin, device=dev); w = torch.randn(k, k, cin, cout, device=dev) * 0.05
    y = torch.empty(B, *dims, cout, device=dev); gy = torch.randn(B, *dims, cout, device=dev); gw = torch.empty_like(w)
    def run():
        if OP == "fwd":
            L.call("cn_conv_fwd", ctypes.byref(d), P(x), P(w), None, 0, 0.0, P(y), 0, st())
        else:
            L.call("cn_conv_wgrad", ctypes.byref(d), P(x), P(gy), P(gw), None, 0, st())
    for dbg in [int(a) for a in sys.argv[1:]] or [0]:
        lib.cn_debug_set(dbg)
        for _ in range(2):
            run()
        lib.cn_debug_set_prof(ctypes.c_void_p(buf.data_ptr()))
        run()
        torch.cuda.synchronize()
        lib.cn_debug_set_prof(None)
        v = buf.cpu().numpy().astype(float); nkb = max(v[9], 1)
        print("cin=%d cout=%d dbg=%d  kb=%d  [clk per k-block]" % (cin, cout, dbg, nkb))
        print("   gather warp0 (per own k-block = x2): wait_empty %.0f  put %.0f  load %.0f  | role total %.0f" % (tuple(2 * v[i] / nkb for i in (0, 1, 2)) + (v[3],)))
        print("   mma: wait_acc %.0f  wait_A %.0f  wait_B %.0f  issue %.0f | role total %.0f" % (v[4] / nkb, v[5] / nkb, v[6] / nkb, v[7] / nkb, v[8]))
        print("   B producer: wait_empty %.0f | role total %.0f   CTA total %.0f clk" % (v[10] / nkb, v[11], v[12]))
    lib.cn_debug_set(0)
