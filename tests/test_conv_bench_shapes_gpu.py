"""-m gpu: every conv / Dense geometry the timed bench step launches (the rows of profiles/r01_conv_breakdown_final.txt:
BASELINE.json configs[1] at batch 32 per GPU, i.e. persistent-CTA launches with more work items than SMs, split-K at the
bench's M, phased / folded plans at full size), forward + input gradient + weight / bias gradient through the C ABI
with the kernel selection the product uses (IMPL_AUTO).

Checker: the oracle's conv restatement (oracle.confignet_oracle.conv_same / upsample_nearest2 -> F.conv, each citing its
reference line) evaluated in fp64.  At these sizes (up to 77 GFLOP per call) it runs on the device for speed - the
same oracle code, torch's fp64 kernels as the arithmetic; the small-shape tests in test_ops_gpu.py run it on the CPU.
Tolerance 2e-4 (max|a-b| / max|b|): the 3xTF32 tensor-core bar of test_ops_gpu.py."""
import ast
import os
import numpy as np
import pytest
import torch

from oracle import confignet_oracle as O
from parity_utils import nerr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bench_geometries():
    """unique (nd, batch, in_dims, cin, cout, ksize, stride, upsample) keys of the committed per-layer table"""
    keys = []
    with open(os.path.join(ROOT, "profiles", "r01_conv_breakdown_final.txt")) as fp:
        for line in fp.readlines()[2:]:
            a, b = line.find("("), line.rfind(")")
            if a < 0:
                continue
            k = ast.literal_eval(line[a:b + 1])
            if k not in keys:
                keys.append(k)
    return keys


GEOMS = bench_geometries()


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def test_table_covers_the_bench_step():
    assert len(GEOMS) >= 60
    assert (2, 16, (64, 64, 1), 256, 256, (3, 3, 1), 1, 1) in GEOMS and (3, 16, (8, 8, 8), 256, 128, (3, 3, 3), 1, 2) in GEOMS


@pytest.mark.parametrize("key", GEOMS, ids=[str(k).replace(" ", "") for k in GEOMS])
def test_bench_layer_matches_oracle(dev, key):
    from confignet_b200 import ops
    nd, B, in_dims, cin, cout, ksize, stride, up = key[:8]
    pad = key[8] if len(key) > 8 else -1
    if pad >= 0:
        pytest.skip("explicit padding (ResNet50 stem) is covered by test_stage2_gpu.py::test_stem_conv_explicit_padding")
    g = torch.Generator(device="cpu").manual_seed(hash(key) % (1 << 31))
    dims = tuple(in_dims[:nd])
    x = torch.randn((B,) + dims + (cin,), generator=g)
    w = torch.randn(tuple(ksize[:nd]) + (cin, cout), generator=g) / float(np.sqrt(cin * np.prod(ksize[:nd])))
    b = torch.randn(cout, generator=g)
    xg, wg, bg = [t.to(dev).requires_grad_(True) for t in (x, w, b)]
    y = ops.conv_act(xg, wg, bg, stride=stride, upsample=up)
    gy = torch.randn(y.shape, generator=g).to(dev)
    gx, gw, gb = torch.autograd.grad(y, (xg, wg, bg), gy)
    # oracle in fp64 on the device
    xr, wr, br = [t.to(dev).double().requires_grad_(True) for t in (x, w, b)]
    xu = O.upsample_nearest2(xr) if up == 2 else xr
    yr = (xr @ wr + br) if nd == 0 else O.conv_same(xu, wr, br, stride)
    gxr, gwr, gbr = torch.autograd.grad(yr, (xr, wr, br), gy.double())
    errs = dict(y=nerr(y, yr), gx=nerr(gx, gxr), gw=nerr(gw, gwr), gb=nerr(gb, gbr))
    assert max(errs.values()) <= 2e-4, errs
    # the launch is deterministic: a second call returns the same bits (split-K slabs, no atomics)
    y2 = ops.conv_act(xg, wg, bg, stride=stride, upsample=up)
    gx2, gw2, gb2 = torch.autograd.grad(y2, (xg, wg, bg), gy)
    assert torch.equal(y, y2) and torch.equal(gx, gx2) and torch.equal(gw, gw2) and torch.equal(gb, gb2)
