"""In-kernel role timings (clock64) of CTA 0 of the tcgen05 pixel kernel for a few layers (test hook cn_debug_set_prof).
Prints, per k-block: gather warp 0 (waiting for a free A stage / splitting + tcgen05.st incl. waiting for its loads /
issuing loads), the MMA warp (waiting for accumulator, A, B / issuing), the B producer, and the CTA's total clocks."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import os
os.environ.setdefault("CN_TEST_HOOKS", "1")      # the cn_debug_* hooks live in libconfignet_b200_hooks.so only
from confignet_b200 import _lib as L
lib = L.load(); dev = torch.device("cuda:0")
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
lib.cn_debug_set_prof.argtypes = [ctypes.c_void_p]
buf = torch.zeros(16, dtype=torch.int64, device=dev)
SHAPES = [(2, 16, (64, 64), 256, 256, 3, 1, 1, "fwd"), (2, 16, (64, 64), 256, 256, 3, 1, 1, "dgrad"),
          (2, 16, (256, 256), 64, 64, 3, 1, 1, "fwd"), (2, 32, (128, 128), 48, 96, 3, 2, 1, "dgrad"),
          (2, 32, (128, 128), 48, 96, 3, 2, 1, "wgrad"), (2, 16, (128, 128), 128, 128, 3, 1, 1, "fwd")]
if len(sys.argv) > 1:
    SHAPES = SHAPES[:int(sys.argv[1])]
for (nd, B, dims, cin, cout, k, s, up, op) in SHAPES:
    d = L.make_conv_desc(nd, B, dims, cin, cout, [k] * nd, s, up)
    od = (ctypes.c_int * 3)(); L.call("cn_conv_out_dims", ctypes.byref(d), od)
    x = torch.randn(B, *dims, cin, device=dev); w = torch.randn(*([k] * nd), cin, cout, device=dev) * 0.05
    oshape = (B,) + tuple(od[:nd]) + (cout,)
    y = torch.empty(oshape, device=dev); gy = torch.randn(oshape, device=dev); gw = torch.empty_like(w); gx = torch.empty_like(x)
    bias = torch.randn(cout, device=dev)

    def run():
        if op == "fwd":
            L.call("cn_conv_fwd", ctypes.byref(d), P(x), P(w), P(bias), 0, 0.0, P(y), 0, st())
        elif op == "dgrad":
            L.call("cn_conv_dgrad", ctypes.byref(d), P(gy), P(w), P(gx), 0, st())
        else:
            L.call("cn_conv_wgrad", ctypes.byref(d), P(x), P(gy), P(gw), None, 0, st())
    for cl, dbgmask in ((1, 0), (1, 16), (1, 4), (1, 1)):
        lib.cn_debug_set_cluster(cl)
        lib.cn_debug_set(dbgmask)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3
        lib.cn_debug_set_prof(ctypes.c_void_p(buf.data_ptr()))
        run()
        torch.cuda.synchronize()
        lib.cn_debug_set_prof(None)
        v = buf.cpu().numpy().astype(float); nkb = max(v[9], 1)
        print("%s %s cluster=%d dbg=%d (16 = no epilogue stores, 4 = no MMA issue, 1 = no A loads): %.1f us/call, CTA0: %d k-blocks, total %.0f clk (%.0f clk per k-block)" %
              (op, (nd, B, dims, cin, cout, k, s, up), cl, dbgmask, us, nkb, v[12], v[12] / nkb))
        print("   gather warp0 per own k-block: wait_free_stage %.0f  split+st(+load wait) %.0f  issue_loads %.0f | role total %.0f" %
              (tuple(2 * v[i] / nkb for i in (0, 1, 2)) + (v[3],)))
        print("   mma per k-block: wait_acc %.0f  wait_A %.0f  wait_B %.0f  issue %.0f | role total %.0f" %
              (v[4] / nkb, v[5] / nkb, v[6] / nkb, v[7] / nkb, v[8]))
        print("   B producer: wait_free_stage %.0f per k-block | role total %.0f" % (v[10] / nkb, v[11]), flush=True)
    lib.cn_debug_set(0)
