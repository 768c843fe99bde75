"""GPU check of the conv kernel family: FFMA vs torch-CPU fp64 (small), tcgen05 vs FFMA/CPU (larger).
Run on the GPU box:  python scripts/gpu_conv_check.py [quick]
Prints one line per (config, op, impl): normalised max error = max|a-b| / max|b|."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes
import torch.nn.functional as F
from confignet_b200 import _lib as L

lib = L.load()
dev = torch.device("cuda:0")
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None


def ref_conv(x, w, s, up):
    nd = x.dim() - 2
    if up == 2:
        for d in range(1, nd + 1):
            x = x.repeat_interleave(2, dim=d)
    if nd == 0:
        return x @ w
    perm_in = (0, nd + 1) + tuple(range(1, nd + 1))
    xc = x.permute(*perm_in)
    wc = w.permute(nd + 1, nd, *range(nd))
    pads = []
    for d in reversed(range(nd)):
        i, k = x.shape[1 + d], w.shape[d]
        o = -(-i // s); tot = max((o - 1) * s + k - i, 0)
        pads += [tot // 2, tot - tot // 2]
    xc = F.pad(xc, pads)
    y = (F.conv2d if nd == 2 else F.conv3d)(xc, wc, None, stride=s)
    return y.permute(0, *range(2, nd + 2), 1)


def nerr(a, b):
    a = a.double().cpu(); b = b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def run(nd, B, dims, cin, cout, k, s, up, impls, refmode="cpu"):
    torch.manual_seed(0)
    d = L.make_conv_desc(nd, B, dims, cin, cout, [k] * nd, s, up)
    x = torch.randn(B, *dims, cin, device=dev)
    w = torch.randn(*([k] * nd), cin, cout, device=dev) * (1.0 / np.sqrt(cin * k ** nd))
    bias = torch.randn(cout, device=dev)
    od = (ctypes.c_int * 3)(); L.call("cn_conv_out_dims", ctypes.byref(d), od)
    oshape = (B,) + tuple(od[:nd]) + (cout,)
    gy = torch.randn(*oshape, device=dev)
    if refmode == "cpu":
        xr = x.double().cpu().requires_grad_(True); wr = w.double().cpu().requires_grad_(True)
        yr = ref_conv(xr, wr, s, up) + bias.double().cpu()
        gxr, gwr = torch.autograd.grad(yr, (xr, wr), gy.double().cpu())
        gbr = gy.double().cpu().reshape(-1, cout).sum(0)
        yr = yr.detach()
    else:   # FFMA kernels as reference (already validated against the CPU)
        yr = torch.empty(oshape, device=dev); gxr = torch.empty_like(x); gwr = torch.empty_like(w); gbr = torch.empty(cout, device=dev)
        L.call("cn_conv_fwd", ctypes.byref(d), P(x), P(w), P(bias), 0, 0.0, P(yr), L.IMPL_FFMA, st())
        L.call("cn_conv_dgrad", ctypes.byref(d), P(gy), P(w), P(gxr), L.IMPL_FFMA, st())
        L.call("cn_conv_wgrad", ctypes.byref(d), P(x), P(gy), P(gwr), P(gbr), L.IMPL_FFMA, st())
        torch.cuda.synchronize()
    for impl in impls:
        name = {0: "auto", 1: "ffma", 2: "tc"}[impl]
        res = []
        for op in ("fwd", "dgrad", "wgrad"):
            try:
                t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
                if op == "fwd":
                    y = torch.full(oshape, 3.0, device=dev)
                    t0.record(); L.call("cn_conv_fwd", ctypes.byref(d), P(x), P(w), P(bias), 0, 0.0, P(y), impl, st()); t1.record()
                    torch.cuda.synchronize(); e = nerr(y, yr)
                elif op == "dgrad":
                    gx = torch.full_like(x, 3.0)
                    t0.record(); L.call("cn_conv_dgrad", ctypes.byref(d), P(gy), P(w), P(gx), impl, st()); t1.record()
                    torch.cuda.synchronize(); e = nerr(gx, gxr)
                else:
                    gw = torch.full_like(w, 3.0); gb = torch.full((cout,), 3.0, device=dev)
                    t0.record(); L.call("cn_conv_wgrad", ctypes.byref(d), P(x), P(gy), P(gw), P(gb), impl, st()); t1.record()
                    torch.cuda.synchronize(); e = max(nerr(gw, gwr), nerr(gb, gbr))
                res.append("%s %.2e (%.3f ms)" % (op, e, t0.elapsed_time(t1)))
            except L.CnError as ex:
                res.append("%s ERR(%s)" % (op, str(ex)[:60]))
        print("nd=%d B=%d dims=%s cin=%d cout=%d k=%d s=%d up=%d [%s vs %s]  " % (nd, B, dims, cin, cout, k, s, up, name, refmode) + " | ".join(res), flush=True)


if __name__ == "__main__":
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    small = [
        (2, 2, (8, 8), 3, 5, 4, 1, 1), (2, 2, (8, 8), 3, 5, 3, 2, 1), (2, 1, (7, 9), 2, 3, 3, 2, 1),
        (2, 2, (4, 6), 3, 4, 4, 1, 2), (3, 1, (4, 4, 4), 2, 3, 3, 1, 2), (0, 5, (), 7, 3, 1, 1, 1),
        (2, 2, (16, 16), 3, 48, 3, 2, 1), (2, 2, (16, 16), 32, 3, 4, 1, 2), (0, 8, (), 2048, 148, 1, 1, 1),
    ]
    for c in small:
        run(*c, impls=[L.IMPL_FFMA])
    tc = [
        (2, 2, (16, 16), 64, 64, 3, 1, 1), (2, 2, (16, 16), 32, 32, 4, 1, 1), (2, 2, (16, 16), 48, 96, 3, 2, 1),
        (2, 2, (8, 8), 64, 32, 4, 1, 2), (3, 2, (4, 4, 4), 64, 32, 3, 1, 2), (3, 1, (8, 8, 8), 32, 64, 3, 1, 1),
        (2, 1, (16, 16), 128, 256, 1, 1, 1), (2, 2, (32, 32), 96, 192, 3, 2, 1), (2, 4, (16, 16), 512, 256, 4, 1, 1),
    ]
    for c in tc:
        run(*c, impls=[L.IMPL_TC], refmode="cpu")
    if not quick:
        big = [
            (2, 8, (256, 256), 64, 64, 3, 1, 1), (3, 8, (8, 8, 8), 256, 128, 3, 1, 2), (3, 8, (4, 4, 4), 512, 256, 3, 1, 2),
            (2, 8, (128, 128), 48, 96, 3, 2, 1), (2, 8, (64, 64), 32, 32, 4, 1, 2), (2, 8, (32, 32), 256, 256, 3, 1, 1),
        ]
        for c in big:
            run(*c, impls=[L.IMPL_TC, L.IMPL_FFMA], refmode="ffma")
