"""Kernel timeline of generate_images at batch 1 (BASELINE config 1; torch.profiler / CUPTI over replayed graphs): which
kernels the 0.83 ms are made of, in launch order, and the idle time between them.  Not a bench number."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity

import bench
from confignet_b200.confignet_first_stage import ConfigNetFirstStage

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda:0")
model = ConfigNetFirstStage({"output_shape": (256, 256, 3), "batch_size": 4, "facemodel_inputs": bench.facemodel_cfg()}, device=dev)
np.random.seed(0)
lat, rot = model.sample_latent_vector(B), model.sample_rotations(B)
for _ in range(4):
    model.generate_images(lat, rot)
torch.cuda.synchronize()
N = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        model.generate_images(lat, rot)
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
per = len(evs) // N
last = evs[-per:]
span = last[-1].time_range.end - last[0].time_range.start
busy = sum(e.time_range.end - e.time_range.start for e in last)
print("generate_images batch %d: %d device events per call, span %.1f us, kernel-busy %.1f us, idle %.1f us" % (B, per, span, busy, span - busy))
agg = {}
for e in last:
    k = e.name[:70]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += e.time_range.end - e.time_range.start
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    print("%-72s %4d  %8.1f us  %5.1f%%" % (k, n, t, 100 * t / span))
print("the ten longest single launches, in order of appearance:")
top = sorted(last, key=lambda e: -(e.time_range.end - e.time_range.start))[:10]
for e in sorted(top, key=lambda e: e.time_range.start):
    print("  %8.1f us  at +%7.1f us  %s" % (e.time_range.end - e.time_range.start, e.time_range.start - last[0].time_range.start, e.name[:80]))
