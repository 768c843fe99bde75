"""Shared helpers of the parity tests: identical seeded parameters for the CPU oracle and the CUDA path."""
from collections import OrderedDict
import numpy as np
import torch
from confignet_b200 import netspec
from oracle import confignet_oracle as O


def nerr(a, b):
    """max |a-b| / max |b| (the 'relative fp32 tolerance' of the parity gate is measured on this)."""
    a = torch.as_tensor(np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a)).double()
    b = torch.as_tensor(np.asarray(b.detach().cpu() if isinstance(b, torch.Tensor) else b)).double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def make_params(spec, seed, device, perturb=0.05, vgg_like=False, dtype=torch.float32):
    """-> (oracle params on CPU requiring grad, CUDA ParamGroup)"""
    from confignet_b200.runtime import ParamGroup
    arrays = netspec.init_params(spec, seed, vgg_like=vgg_like)
    if perturb:
        arrays = netspec.perturb_params(arrays, seed + 1000, perturb)
    p_cpu = O.to_torch(arrays, dtype=dtype, requires_grad=True)
    group = ParamGroup(arrays, device)
    return p_cpu, group


def grads_cpu(loss, p_cpu):
    gs = torch.autograd.grad(loss, list(p_cpu.values()), allow_unused=True)
    return OrderedDict((k, (torch.zeros_like(v) if g is None else g)) for (k, v), g in zip(p_cpu.items(), gs))


def grads_gpu(loss, group):
    gs = torch.autograd.grad(loss, group.trainable_weights, allow_unused=True)
    return OrderedDict((k, (torch.zeros_like(v) if g is None else g)) for (k, v), g in zip(group.params.items(), gs))


def l2err(a, b):
    a = torch.as_tensor(np.asarray(a.detach().cpu())).double()
    b = torch.as_tensor(np.asarray(b.detach().cpu())).double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def compare_grads(g_gpu, g_cpu, tol, what="", metric="max"):
    """metric 'max': max|a-b|/max|b| per tensor.  metric 'l2': ||a-b||/||b|| per tensor, skipping tensors whose
    reference norm is below 1e-3 of the largest one (conv biases in front of an InstanceNorm have a true
    gradient of ~0 that is pure cancellation noise in ANY fp32 implementation)."""
    worst = (0.0, None)
    top = max(float(v.double().norm()) for v in g_cpu.values())
    for k in g_cpu:
        if metric == "l2":
            if float(g_cpu[k].double().norm()) < 1e-3 * top:
                continue
            e = l2err(g_gpu[k], g_cpu[k])
        else:
            e = nerr(g_gpu[k], g_cpu[k]) if float(g_cpu[k].abs().max()) > 0 else float(g_gpu[k].abs().max())
        if e > worst[0]:
            worst = (e, k)
    assert worst[0] <= tol, "%s gradient mismatch (%s): %s err %.3e > %.1e" % (what, metric, worst[1], worst[0], tol)
    return worst
