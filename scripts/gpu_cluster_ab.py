"""Per-layer A/B of the tensor-core conv kernels: cluster size 1 vs 2 (weight-stream multicast) for every
tcgen05 (op, layer) of the bench step (keys read from a bench.py --breakdown table).  Not a bench number."""
import sys, os, re, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import os
os.environ.setdefault("CN_TEST_HOOKS", "1")      # the cn_debug_* hooks live in libconfignet_b200_hooks.so only
from confignet_b200 import _lib as L
lib = L.load(); dev = torch.device("cuda:0")
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
rows = []
for l in open(sys.argv[1]).read().splitlines()[2:]:
    m = re.match(r"(\w+)\s+(\(.*\))\s+(\d)\s+(\d+)\s+([\d.]+)\s+([\d.]+)", l)
    if m and int(m.group(3)) == 2:
        rows.append((m.group(1), eval(m.group(2)), int(m.group(4)), float(m.group(5))))


def time_it(fn, reps=7):
    fn(); fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


tot = {1: 0.0, 2: 0.0, "best": 0.0}
print("%-6s %-62s calls   c1 us    c2 us   c2/c1" % ("op", "layer"))
for op, key, calls, _ in rows:
    nd, B, dims, cin, cout, k, s, up = key
    dims, k = tuple(dims[:nd]), list(k[:nd])
    d = L.make_conv_desc(nd, B, dims, cin, cout, k, s, up)
    od = (ctypes.c_int * 3)(); L.call("cn_conv_out_dims", ctypes.byref(d), od)
    x = torch.randn(B, *dims, cin, device=dev); w = torch.randn(*k, cin, cout, device=dev) * 0.05
    oshape = (B,) + tuple(od[:nd]) + (cout,)
    gy = torch.randn(*oshape, device=dev); y = torch.empty(oshape, device=dev); gx = torch.empty_like(x); gw = torch.empty_like(w)
    fns = {"fwd": lambda: L.call("cn_conv_fwd", ctypes.byref(d), P(x), P(w), None, 0, 0.0, P(y), 0, st()),
           "dgrad": lambda: L.call("cn_conv_dgrad", ctypes.byref(d), P(gy), P(w), P(gx), 0, st()),
           "wgrad": lambda: L.call("cn_conv_wgrad", ctypes.byref(d), P(x), P(gy), P(gw), None, 0, st())}
    t = {}
    for c in (1, 2):
        lib.cn_debug_set_cluster(c)
        t[c] = time_it(fns[op]) * 1e3
        tot[c] += t[c] * calls
    tot["best"] += min(t[1], t[2]) * calls
    print("%-6s %-62s %4d %8.1f %8.1f %6.2f" % (op, str(key), calls, t[1], t[2], t[2] / t[1]), flush=True)
print("per-step totals (ms): cluster1 %.2f  cluster2 %.2f  per-layer best %.2f" % (tot[1] / 1e3, tot[2] / 1e3, tot["best"] / 1e3))
