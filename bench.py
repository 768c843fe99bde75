"""bench.py - G+D training-step throughput of the ConfigNet hot path on B200 (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference (oracle/)

A "step" = discriminator_training_step + generator_training_step (confignet_first_stage.py:466-476,
506-560) on one synthetic 256x256 batch of 32 images per GPU.  Prints ONE JSON line (rank 0).

  value     images/s with the image / mask stores already resident in HBM (CUDA events, max over ranks)
  e2e       same metric through the public class API with HOST datasets: pinned-memory H2D of every batch
            and a D2H read of all loss terms inside the timed region
  roofline  the tcgen05 implicit-GEMM conv kernels: algorithmic FLOPs / CUDA-event time of their launches
  cpu_baseline  the oracle (CPU restatement; TensorFlow 2.1 is not installable) on this box's host cores

No work is skipped inside the timed region: both steps run forward, backward (incl. the R1 double
backward), gradient packing, [all-reduce], and the Adam updates.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("CN_GRAPHS_DP", "1")      # CUDA-graph replay of the steps under data parallelism too (see finish())

import numpy as np
import torch

PER_GPU_BATCH = 32
RES = 256


def facemodel_cfg():
    from confignet_b200 import netspec
    return {k: tuple(v) for k, v in netspec.default_facemodel_inputs().items()}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fp:
            d = json.load(fp)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)), "measured"
    return 6650.0, 1590.0, "fallback"


def bench_config(world):
    return {"workload": "generator+discriminator train_step, synthetic 256x256 batch=32 per GPU (BASELINE.json configs[1])",
            "per_gpu_batch": PER_GPU_BATCH, "global_batch": PER_GPU_BATCH * world, "resolution": RES,
            "parallelism": "dp%d" % world,
            "l2": "inputs larger than L2: >1 GB of activations per step, no flush needed",
            "resident": "image and mask stores in HBM; per-step RNG draws (<100 KB) made by the step",
            "arithmetic": "fp32 in / fp32 out; tensor-core convs as 3xTF32 (tf32 big/small operand split, 3 MMAs per product, "
                          "fp32 accumulation with chunked promotion)"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None, "reasons": reasons}


# ------------------------------------------------------------------------------------------------ reference arm
def oracle_step_inputs(rng, fm, b):
    ns = b // 2
    nr = b - ns
    real_u8 = rng.randint(0, 256, (b, RES, RES, 3), dtype=np.uint8)
    lat = rng.standard_normal((b, 145)).astype(np.float32)
    rot = np.zeros((b, 3), np.float32)
    rot[:, 0] = np.pi * rng.uniform(-30, 30, b) / 180
    rot[:, 1] = np.pi * rng.uniform(-10, 10, b) / 180
    fparams = [rng.uniform(0, 1, (ns, d[0])).astype(np.float32) for d in fm.values()]
    gt_u8 = rng.randint(0, 256, (ns, RES, RES, 3), dtype=np.uint8)
    masks = (rng.rand(ns, RES, RES) < 0.01).astype(np.uint8)
    return dict(real_u8=real_u8, lat=lat, rot=rot, fparams=fparams, gt_u8=gt_u8, masks=masks,
                real_lat=lat[:nr], real_rot=rot[:nr], synth_rot=rot[:ns])


def oracle_one_step(tr, inp):
    d = tr.discriminator_step(inp["real_u8"], inp["lat"], inp["rot"])
    g = tr.generator_step(inp["fparams"], inp["synth_rot"], inp["gt_u8"], inp["masks"], inp["real_lat"], inp["real_rot"])
    return float(d["loss_sum"].detach()) + float(g["loss_sum"].detach())


def time_oracle(steps, warmup, sample_batch):
    from oracle import confignet_oracle as O
    from confignet_b200 import netspec
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fm = netspec.default_facemodel_inputs()
    tr = O.OracleFirstStage(fm, RES)
    rng = np.random.RandomState(0)
    for _ in range(warmup):
        oracle_one_step(tr, oracle_step_inputs(rng, fm, sample_batch))
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_one_step(tr, oracle_step_inputs(rng, fm, sample_batch))
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return sample_batch / dt, dt, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    b = 4
    val, dt, cores = time_oracle(args.steps, min(args.warmup, 1), b)
    sample = "1 D step + 1 G step at batch %d per step (1/%d of the per-GPU batch), torch-CPU fp32 oracle" % (b, PER_GPU_BATCH // b)
    line = {"impl": "reference", "metric": "256x256 face images/sec (G+D step)", "value": val, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(bench_config(args.gpus), reference_sample_batch=b,
                           note="CPU restatement of the reference (TensorFlow 2.1 not installable), bounded sample of the workload"),
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def losses_to_host(*loss_dicts):
    """One D2H read of every loss term (the reference does ~40 float() syncs per iteration)."""
    vals = [v.detach().reshape(1) for d in loss_dicts for v in d.values()]
    return torch.cat(vals).cpu().numpy()


def run_b200(args):
    import torch.distributed as dist
    from confignet_b200 import ops, _lib as L
    from confignet_b200.confignet_first_stage import ConfigNetFirstStage
    from confignet_b200.runtime import KerasAdam
    from confignet_b200.synthetic_data import SyntheticDataset

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda:%d" % local)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created: park stdout on stderr meanwhile
        # (rank 0's stdout carries ONE JSON line)
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    lib = L.load()

    cfg = {"output_shape": (RES, RES, 3), "batch_size": PER_GPU_BATCH * world, "facemodel_inputs": facemodel_cfg()}
    model = ConfigNetFirstStage(cfg, device=dev)
    n_store = 96
    host_real, host_synth = SyntheticDataset(n_store, RES, seed=1), SyntheticDataset(n_store, RES, seed=2)
    dev_real = SyntheticDataset(n_store, RES, seed=1).to_device(dev)
    dev_synth = SyntheticDataset(n_store, RES, seed=2).to_device(dev)
    d_opt, g_opt = KerasAdam(**model.config["optimizer"]), KerasAdam(**model.config["optimizer"])
    np.random.seed(0)

    def step(real_set, synth_set):
        d = model.discriminator_training_step(real_set, d_opt)
        g = model.generator_training_step(real_set, synth_set, g_opt)
        model.update_smoothed_weights()
        return d, g

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step(dev_real, dev_synth)
    barrier()

    # ---- value: stores resident in HBM, CUDA events around the K steps
    sampler = ClockSampler(local)
    if rank == 0:                  # one nvidia-smi poller per job (8 of them would compete with the ranks' host halves)
        sampler.start()
    # roofline: a CUDA-event pair around every conv-family launch INSIDE the timed region (default), or - with
    # --roofline-pass separate - over extra steps right after it (to measure what the ~1200 event records cost)
    inline = args.roofline_pass == "inline"
    prof = []
    if inline:
        ops.PROFILE[0] = prof
    from confignet_b200.runtime import GraphedFn
    lib.cn_launch_count(1)
    replayed0 = GraphedFn.REPLAYED_LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(dev_real, dev_synth)
    e1.record()
    barrier()
    launches = int(lib.cn_launch_count(0)) + (GraphedFn.REPLAYED_LAUNCHES - replayed0)     # eager launches + launches replayed from CUDA graphs
    ops.PROFILE[0] = None
    ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    clocks = sampler.summary()
    value = PER_GPU_BATCH * world / (ms * 1e-3)
    prof_steps = args.steps
    if not inline:
        ops.PROFILE[0] = prof
        prof_steps = max(1, min(args.steps, 4))
        for _ in range(prof_steps):
            step(dev_real, dev_synth)
        barrier()
        ops.PROFILE[0] = None

    # ---- roofline of the tcgen05 conv kernels (events recorded around each launch in the timed region)
    tc_exec = sum(p[6] for p in prof if p[4] == 2)
    tc_flops = sum(p[1] for p in prof if p[4] == 2)
    tc_ms = sum(p[2].elapsed_time(p[3]) for p in prof if p[4] == 2)
    cc_ms = sum(p[2].elapsed_time(p[3]) for p in prof if p[4] != 2)
    all_flops = sum(p[1] for p in prof)
    hbm, tensor_peak, how = measured_peaks()
    achieved = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "tc_dram_traffic.json")      # written by scripts/summarize_launches.py from ncu
    if os.path.exists(tpath):
        with open(tpath) as fp:
            t = json.load(fp)
        traffic, traffic_src = t.get("dram_bytes_per_launch"), t.get("source")
    n_tc = sum(1 for p in prof if p[4] == 2)
    roofline = {"bound": "tensor", "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
                "frac": achieved / tensor_peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": how + " cuBLAS bf16 dense (sustained); kind::tf32 runs at half of it and every fp32 product "
                               "takes 3 tf32 MMAs (3xTF32), so frac <= 1/6 on executed FLOPs",
                "kernel": "igemm_tc_pixel_kernel<B_MN,WG> (tcgen05 kind::tf32 fwd/dgrad/wgrad, 3 MMAs per product)",
                "algorithmic_gflop_per_launch": tc_flops / max(n_tc, 1) / 1e9,
                "avg_launch_us": tc_ms * 1e3 / max(n_tc, 1),
                "executed_tflops": tc_exec / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0,
                "executed_note": "FLOPs executed after sub-pixel folding of UpSampling(2)+conv; achieved counts the "
                                 "reference formulation (convs on the upsampled grid)",
                "launches": sum(1 for p in prof if p[4] == 2) // prof_steps,
                "tc_ms_per_step": tc_ms / prof_steps, "cuda_core_conv_ms_per_step": cc_ms / prof_steps,
                "algorithmic_conv_tflop_per_step": all_flops / prof_steps / 1e12,
                "step_algorithmic_tflops": all_flops / prof_steps / 1e12 / (ms * 1e-3),
                "measured": ("CUDA events around each conv launch inside the timed region" if inline else
                             "CUDA events around each conv launch over %d extra steps right after the timed region" % prof_steps)}

    if args.breakdown and rank == 0:
        agg = {}
        for (op, f, a, b, impl, key, _fx) in prof:
            k = (op, key, impl)
            t = agg.setdefault(k, [0, 0.0, 0.0])
            t[0] += 1; t[1] += a.elapsed_time(b); t[2] += f
        rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
        with open(args.breakdown, "w") as fp:
            fp.write("per-step conv/dense launches, %d steps averaged; impl 2 = tcgen05, 1 = CUDA-core\n" % prof_steps)
            fp.write("%-6s %-62s impl calls   ms/step  TFLOP/s\n" % ("op", "(nd,batch,in_dims,cin,cout,ksize,stride,upsample)"))
            for (op, key, impl), (n, t, f) in rows:
                fp.write("%-6s %-62s %4d %5d %9.3f %8.2f\n" % (op, str(key), impl, n // prof_steps, t / prof_steps,
                                                             f / (t * 1e-3) / 1e12 if t > 0 else 0))
    def finish():
        if world > 1:
            # The captured step graphs hold NCCL work: tearing the communicator down under them hangs (measured: the
            # 2-GPU run sat in destroy_process_group until its timeout).  Everything is measured and printed - drop
            # the graphs, meet the other ranks once more and leave without the teardown.
            sys.stdout.flush()
            model._graphs.clear()
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)

    if args.no_e2e:
        if rank == 0:
            print(json.dumps({"value": value, "ms_per_step": ms, "note": "profiling run (no e2e leg)"}), flush=True)
        finish()
        return
    # ---- e2e: public API, host datasets, H2D + D2H inside the timed region
    for _ in range(2):
        d, g = step(host_real, host_synth)
        losses_to_host(d, g)
    d2h = 0
    # Every step's losses are read back inside the timed region, but without draining the GPU: the D2H copy into a
    # pinned buffer is enqueued right behind the step, and the host looks at step k-1's buffer while step k runs.
    n_loss = sum(len(x) for x in step(host_real, host_synth))
    pinned = [torch.zeros(n_loss, dtype=torch.float32).pin_memory() for _ in range(2)]

    def enqueue_readback(i, *loss_dicts):
        vals = torch.cat([v.detach().reshape(1) for dct in loss_dicts for v in dct.values()])
        pinned[i % 2].copy_(vals, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return pinned[i % 2], ev

    ConfigNetFirstStage.h2d_bytes = 0
    barrier()
    t0 = time.perf_counter()
    pending = None
    for i in range(args.steps):
        cur = step(host_real, host_synth)          # host halves (sampling, pinned uploads) + asynchronous device halves
        rb = enqueue_readback(i, *cur)
        if pending is not None:
            pending[1].synchronize()
            d2h += pending[0].numpy().copy().nbytes
        pending = rb
    pending[1].synchronize()
    d2h += pending[0].numpy().copy().nbytes
    torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0) / args.steps
    e2e = {"value": PER_GPU_BATCH * world / dt, "unit": "images/s",
           "h2d_bytes_per_step": ConfigNetFirstStage.h2d_bytes // args.steps, "d2h_bytes_per_step": d2h // args.steps}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        b = 16
        val, cdt, cores = time_oracle(3, 1, b)
        cpu_baseline = {"value": val, "unit": "images/s", "cores": cores, "kind": "port",
                        "sample": "3 x (1 D step + 1 G step) at batch %d (1/%d of the workload batch) after 1 warm-up step, "
                                  "torch-CPU fp32 oracle (CPU restatement of the reference; TensorFlow 2.1 is not "
                                  "installable), %.1f s per step" % (b, PER_GPU_BATCH // b, cdt)}
    if rank == 0:
        line = {"metric": "256x256 face images/sec (G+D step)", "value": value, "unit": "images/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": bench_config(world),
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
                "cpu_baseline": cpu_baseline}
        print(json.dumps(line), flush=True)
    finish()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the end-to-end leg")
    ap.add_argument("--roofline-pass", default="separate", choices=["inline", "separate"],
                    help="separate (default): the timed region runs the product path as is (CUDA-graph replays), the "
                         "per-launch event pairs of the roofline run over extra eager steps right after it; inline: event "
                         "pairs inside the timed region (forces eager execution)")
    ap.add_argument("--breakdown", default=None, help="write a per-layer conv time table to this file")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a GPU (the CUDA path has no CPU fallback)")
        run_b200(args)


if __name__ == "__main__":
    main()
