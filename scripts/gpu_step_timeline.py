"""GPU-busy vs wall time of the bench step (torch.profiler / CUPTI; no nsys in the image).

Prints total kernel time, idle gaps, and the top kernels by GPU time for `--steps` G+D steps so that
host-bound stretches (launch gaps) can be told from kernel-bound ones.  Not a bench number.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import bench
    from confignet_b200.confignet_first_stage import ConfigNetFirstStage
    from confignet_b200.runtime import KerasAdam
    from confignet_b200.synthetic_data import SyntheticDataset
    dev = torch.device("cuda:0")
    cfg = {"output_shape": (256, 256, 3), "batch_size": 32, "facemodel_inputs": bench.facemodel_cfg()}
    model = ConfigNetFirstStage(cfg, device=dev)
    real = SyntheticDataset(96, 256, seed=1).to_device(dev)
    synth = SyntheticDataset(96, 256, seed=2).to_device(dev)
    d_opt, g_opt = KerasAdam(**model.config["optimizer"]), KerasAdam(**model.config["optimizer"])
    np.random.seed(0)

    def step():
        model.discriminator_training_step(real, d_opt)
        model.generator_training_step(real, synth, g_opt)
        model.update_smoothed_weights()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    busy = sum(e.time_range.end - e.time_range.start for e in evs)
    span = evs[-1].time_range.end - evs[0].time_range.start
    gaps = []
    cur_end = evs[0].time_range.end
    for e in evs[1:]:
        if e.time_range.start > cur_end:
            gaps.append(e.time_range.start - cur_end)
        cur_end = max(cur_end, e.time_range.end)
    lines = []
    lines.append("steps %d: span %.2f ms/step, kernel-busy %.2f ms/step, idle %.2f ms/step in %d gaps (median gap %.1f us), %d device events/step"
                 % (args.steps, span / 1e3 / args.steps, busy / 1e3 / args.steps, sum(gaps) / 1e3 / args.steps,
                    len(gaps) // args.steps, float(np.median(gaps)) if gaps else 0.0, len(evs) // args.steps))
    agg = {}
    for e in evs:
        a = agg.setdefault(e.name[:90], [0, 0.0])
        a[0] += 1
        a[1] += e.time_range.end - e.time_range.start
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        lines.append("%-92s %6d %9.3f ms/step %5.1f%%" % (name, n // args.steps, t / 1e3 / args.steps, 100 * t / busy))
    text = "\n".join(lines)
    print(text)
    if args.out:
        with open(args.out, "w") as fp:
            fp.write(text + "\n")


if __name__ == "__main__":
    main()
