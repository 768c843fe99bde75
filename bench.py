"""bench.py - throughput of the ConfigNet hot path on B200 for the five BASELINE.json configurations.

    python bench.py [--config C] --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py [--config C] --impl reference --steps K --warmup W    # CPU restatement of the reference (oracle/)

--config (default 2, the configuration BASELINE.json's metric is quoted on), all through the class API
(confignet_b200.ConfigNetFirstStage / ConfigNet / LatentGAN) on synthetic 256x256 data, sizes PER GPU (weak scaling):
  1  generate_images, batch 1 (confignet_first_stage.py:633-639)                                   unit: images/s (and ms latency)
  2  discriminator_training_step + generator_training_step + EMA, batch 32 (:466-476,506-560)      unit: images/s
  3  full first-stage iteration: D, synth-D, latent-D, G steps + EMA, batch 64 (:604-626)          unit: images/s
  4  full second-stage iteration (adds the ResNet50 encoder), batch 32 per GPU - 128 over 4 GPUs (confignet_second_stage.py:277-299)
  5  LatentGAN D+G step at batch 256 global (latent_gan.py:234-247) + fine_tune_on_img on 32 images per GPU, n_iters = 10
     (confignet_second_stage.py:321-403)                                                          unit: image-iterations/s

Prints ONE JSON line (rank 0):
  value         units/s with the image / mask / embedding stores already resident in HBM (CUDA events, max over ranks)
  e2e           the same metric through the public API with HOST (NumPy) inputs: pinned-memory H2D of every batch and a D2H
                read of every result (all loss terms / the uint8 images / the fine-tuned embeddings) inside the timed region
  roofline      the tcgen05 implicit-GEMM conv kernels: algorithmic FLOPs / CUDA-event time of their launches
  cpu_baseline  the oracle (CPU restatement; TensorFlow 2.1 is not installable) on this box's host cores at the workload batch

No work is skipped inside the timed region: every step runs forward, backward (incl. the R1 double backward), gradient
packing, [all-reduce], and the Adam updates.
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

import numpy as np
import torch

RES = 256
METRIC = "256x256 face images/sec (G+D step)"

CONFIGS = {
    1: dict(per_gpu=1, unit="images/s",
            workload="single 256x256 generator forward on random latent, batch=1 (generate_images; BASELINE.json configs[0])"),
    2: dict(per_gpu=32, unit="images/s",
            workload="generator+discriminator train_step, synthetic 256x256 batch=32 per GPU (BASELINE.json configs[1])"),
    3: dict(per_gpu=64, unit="images/s",
            workload="full ConfigNet first-stage iteration (D + synth-D + latent-D + G steps, perceptual loss, EMA), batch=64 per GPU "
                     "(BASELINE.json configs[2])"),
    4: dict(per_gpu=32, unit="images/s",
            workload="full ConfigNet second-stage iteration (real + synthetic encoders + generator), batch=32 per GPU = 128 over 4 GPUs, "
                     "NCCL gradient all-reduce (BASELINE.json configs[3])"),
    5: dict(per_gpu=32, unit="image-iterations/s",
            workload="LatentGAN D+G step (batch 256 global) + one-shot fine_tune_on_img generator loop, 32 images per GPU x n_iters=10 "
                     "(BASELINE.json configs[4])"),
}
FT_ITERS = 10
LGAN_BATCH = 256


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


def bench_config(cfg_id, world):
    c = CONFIGS[cfg_id]
    return {"workload": c["workload"], "config_id": cfg_id, "per_gpu_batch": c["per_gpu"], "global_batch": c["per_gpu"] * world,
            "resolution": RES, "parallelism": "dp%d" % world,
            "l2": "inputs larger than L2: >1 GB of activations per step, no flush needed" if cfg_id != 1 else
                  "batch-1 latency path: the generator's 32 MB of kernels and its activations stay in the 126 MB L2 by design (the "
                  "demo loop's steady state); no flush",
            "resident": "image / mask / embedding stores in HBM; per-step RNG draws (<100 KB) made by the step",
            "arithmetic": "fp32 in / fp32 out; tensor-core convs as 3xTF32 (tf32 big/small operand split, 3 MMAs per product, "
                          "fp32 accumulation with chunked promotion)"}


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region, read through NVML in this process (the library nvidia-smi is
    a front end of): forking an nvidia-smi child out of a multi-GB Python process every 200 ms stalled the host half of
    the first timed step by up to 30 ms (measured, profiles/r02_first_step_outlier.txt).  Falls back to nvidia-smi when
    the NVML binding is missing.  The thread starts before the warm-up; only samples between begin() and summary() count."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.active = index, [], False, False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = index
            if vis and all(x.strip().isdigit() for x in vis.split(",")):
                phys = int(vis.split(",")[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def sample(self):
        try:
            if self.nvml is not None:
                n = self.nvml
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
                r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                bit = lambda name: "Active" if r & getattr(n, name, 0) else "Not Active"
                f = [str(sm), str(mx), bit("nvmlClocksThrottleReasonHwSlowdown"), bit("nvmlClocksThrottleReasonHwThermalSlowdown"),
                     bit("nvmlClocksThrottleReasonSwThermalSlowdown"), bit("nvmlClocksThrottleReasonSwPowerCap")]
            else:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
            if len(f) >= 6 and self.active:
                self.samples.append(f)
        except Exception:
            pass

    def run(self):
        while not self.stop_flag:
            self.sample()
            time.sleep(0.25)

    def begin(self):
        self.samples, self.active = [], True

    def summary(self):
        if not self.samples:
            self.sample()
        self.stop_flag, self.active = True, False
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None, "reasons": reasons,
                "samples": len(self.samples), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------ CPU arm (oracle/)
class OracleWorkload:
    """One step of configuration `cfg_id` on the CPU restatement of the reference (oracle/), at batch `b` - the same
    step functions the parity tests check the CUDA path against.  -> seconds per step."""

    def __init__(self, cfg_id, b):
        from oracle import confignet_oracle as O
        from oracle import confignet_oracle_stage2 as O2
        from confignet_b200 import netspec
        self.cfg_id, self.b = cfg_id, b
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.fm = netspec.default_facemodel_inputs()
        self.rng = np.random.RandomState(0)
        if cfg_id in (1, 2, 3):
            self.tr = O.OracleFirstStage(self.fm, RES)
        else:
            self.tr = O2.OracleSecondStage(self.fm, RES)
        if cfg_id == 5:
            self.lgan = O2.OracleLatentGAN()
            self.emb = self.rng.standard_normal((10000, 145)).astype(np.float32)

    def _u8(self, n):
        return self.rng.randint(0, 256, (n, RES, RES, 3), dtype=np.uint8)

    def _rot(self, n):
        r = np.zeros((n, 3), np.float32)
        r[:, 0] = np.pi * self.rng.uniform(-30, 30, n) / 180
        r[:, 1] = np.pi * self.rng.uniform(-10, 10, n) / 180
        return r

    def _fm(self, n):
        return [self.rng.uniform(0, 1, (n, d[0])).astype(np.float32) for d in self.fm.values()]

    def _lat(self, n):
        return self.rng.standard_normal((n, 145)).astype(np.float32)

    def step(self, ft_iters=FT_ITERS):
        b, ns = self.b, self.b // 2
        nr = b - ns
        tr = self.tr
        t0 = time.perf_counter()
        if self.cfg_id == 1:
            tr.generate_images(self._lat(b), self._rot(b))
        elif self.cfg_id in (2, 3):
            tr.discriminator_step(self._u8(b), self._lat(b), self._rot(b))
            if self.cfg_id == 3:
                tr.synth_discriminator_step(self._u8(b), self._fm(b), self._rot(b))
                tr.latent_discriminator_step(self._lat(b), self._fm(b))
            masks = (self.rng.rand(ns, RES, RES) < 0.01).astype(np.uint8)
            tr.generator_step(self._fm(ns), self._rot(ns), self._u8(ns), masks, self._lat(nr), self._rot(nr))
        elif self.cfg_id == 4:
            tr.discriminator_step(self._u8(b), self._u8(b))
            tr.synth_discriminator_step(self._u8(b), self._fm(b), self._rot(b))
            tr.latent_discriminator_step(self._u8(b), self._fm(b))
            masks = (self.rng.rand(ns, RES, RES) < 0.01).astype(np.uint8)
            tr.generator_step(self._fm(ns), self._rot(ns), self._u8(ns), masks, self._u8(nr))
        else:
            idx = self.rng.randint(0, self.emb.shape[0], LGAN_BATCH)
            self.lgan.step(self.emb[idx], self._lat(LGAN_BATCH), self._lat(LGAN_BATCH))
            tr.fine_tune(self._u8(b), ft_iters)
        return time.perf_counter() - t0


def units_per_step(cfg_id, batch, ft_iters=FT_ITERS):
    return batch * ft_iters if cfg_id == 5 else batch


def cpu_baseline_leg(cfg_id):
    """bounded sample (10-30 s of CPU work) at the workload batch"""
    b = CONFIGS[cfg_id]["per_gpu"]
    w = OracleWorkload(cfg_id, b)
    if cfg_id == 1:
        w.step()
        ts = sorted(w.step() for _ in range(5))
        dt, sample = ts[2], "median of 5 generate_images calls at batch 1 after 1 warm-up"
        units = 1
    elif cfg_id == 5:
        dt = w.step(ft_iters=2)
        units = units_per_step(5, b, 2)
        sample = "1 LatentGAN D+G step at batch %d + fine_tune on %d images for 2 of the %d iterations (no warm-up)" % (LGAN_BATCH, b, FT_ITERS)
    else:
        n = 2 if cfg_id == 2 else 1
        w.step()
        dt = sum(w.step() for _ in range(n)) / n
        units = b
        sample = "%d step(s) at the workload batch %d after 1 warm-up step, %.1f s per step" % (n, b, dt)
    return {"value": units / dt, "unit": CONFIGS[cfg_id]["unit"], "cores": w.cores, "kind": "port",
            "sample": sample + "; torch-CPU fp32 oracle (CPU restatement of the reference; TensorFlow 2.1 is not installable)"}


def run_reference(args):
    """The reference arm: the oracle port on all host cores, at the workload's per-GPU batch; W warm-up and K timed steps
    as asked, shortened only if they would not end within ~4 minutes (the line then says how many ran)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg_id = args.config
    b = CONFIGS[cfg_id]["per_gpu"]
    w = OracleWorkload(cfg_id, b)
    ft = FT_ITERS if cfg_id != 5 else 2
    budget = 240.0
    t_start = time.perf_counter()
    first = w.step(ft) if cfg_id == 5 else w.step()
    warm_done = 1
    while warm_done < args.warmup and (time.perf_counter() - t_start) + first * 2 < budget * 0.3:
        w.step(ft) if cfg_id == 5 else w.step()
        warm_done += 1
    times = []
    while len(times) < args.steps and (not times or (time.perf_counter() - t_start) + first < budget):
        times.append(w.step(ft) if cfg_id == 5 else w.step())
    dt = sum(times) / len(times)
    units = units_per_step(cfg_id, b, ft)
    val = units / dt
    sample = "%d of %d requested steps (+%d warm-up) at the workload batch %d per step" % (len(times), args.steps, warm_done, b)
    if cfg_id == 5:
        sample += ", fine_tune for %d of the %d iterations per step" % (ft, FT_ITERS)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": CONFIGS[cfg_id]["unit"],
            "n_gpus": args.gpus, "steps": len(times), "warmup": warm_done, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(bench_config(cfg_id, 1), note="CPU restatement of the reference (TensorFlow 2.1 not installable) on the host "
                           "cores, one process, at the per-GPU batch of the workload"),
            "cpu_baseline": {"value": val, "unit": CONFIGS[cfg_id]["unit"], "cores": w.cores, "kind": "port",
                             "sample": sample + "; torch-CPU fp32 oracle"},
            "e2e": {"value": val, "unit": CONFIGS[cfg_id]["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
class Workload:
    """Configuration `cfg_id` through the class API.  step(resident) runs one step on the device-resident (True) or the
    host (False) stores and returns what the caller must read back (device tensors) or has already received (NumPy)."""

    def __init__(self, cfg_id, dev, world, rank):
        from confignet_b200.confignet_first_stage import ConfigNetFirstStage
        from confignet_b200.confignet_second_stage import ConfigNet
        from confignet_b200.latent_gan import LatentGAN
        from confignet_b200.runtime import KerasAdam
        from confignet_b200.synthetic_data import SyntheticDataset
        self.cfg_id, self.dev, self.world = cfg_id, dev, world
        b = CONFIGS[cfg_id]["per_gpu"]
        self.b = b
        cfg = {"output_shape": (RES, RES, 3), "batch_size": b * world, "facemodel_inputs": facemodel_cfg()}
        if cfg_id in (4, 5):
            cfg["image_loss_weight"] = 5e-4
        self.model = (ConfigNet if cfg_id in (4, 5) else ConfigNetFirstStage)(cfg, device=dev)
        m = self.model
        n_store = max(96, 3 * b)
        if cfg_id in (2, 3, 4):
            self.host = (SyntheticDataset(n_store, RES, seed=1), SyntheticDataset(n_store, RES, seed=2))
            self.devs = (SyntheticDataset(n_store, RES, seed=1).to_device(dev), SyntheticDataset(n_store, RES, seed=2).to_device(dev))
            self.d_opt, self.g_opt = KerasAdam(**m.config["optimizer"]), KerasAdam(**m.config["optimizer"])
        rng = np.random.RandomState(3 + rank)
        if cfg_id == 1:
            self.lat = rng.standard_normal((b, m.config["latent_dim"])).astype(np.float32)
            self.rot = np.zeros((b, 3), np.float32)
            self.rot[:, 0] = np.pi * rng.uniform(-30, 30, b) / 180
            self.rot[:, 1] = np.pi * rng.uniform(-10, 10, b) / 180
            self.lat_d, self.rot_d = torch.from_numpy(self.lat).to(dev), torch.from_numpy(self.rot).to(dev)
        if cfg_id == 5:
            self.gan = LatentGAN({"latent_dim": m.config["latent_dim"], "batch_size": LGAN_BATCH}, device=dev)
            self.gan_opt = KerasAdam(**self.gan.config["optimizer"])
            self.emb = np.random.RandomState(5).standard_normal((10000, m.config["latent_dim"])).astype(np.float32)
            self.emb_d = torch.from_numpy(self.emb).to(dev)
            imgs = np.random.RandomState(7).randint(0, 256, (b * world, RES, RES, 3), dtype=np.uint8)      # the GLOBAL image set
            self.imgs = imgs
            self.imgs_d = torch.from_numpy(imgs).to(dev)
        np.random.seed(0)

    def step(self, resident):
        m, c = self.model, self.cfg_id
        if c == 1:
            if resident:
                return [m.generate_images_device(self.lat_d, self.rot_d)]
            return [m.generate_images(self.lat, self.rot)]
        if c == 5:
            d = self.gan.discriminator_training_step(self.emb_d if resident else self.emb, self.gan_opt)
            g = self.gan.generator_training_step(self.gan_opt)
            self.gan.update_smoothed_weights()
            emb, rot = m.fine_tune_on_img(self.imgs_d if resident else self.imgs, n_iters=FT_ITERS)
            return [d, g, emb, rot]
        real, synth = self.devs if resident else self.host
        out = [m.discriminator_training_step(real, self.d_opt)]
        if c in (3, 4):
            out.append(m.synth_discriminator_training_step(synth, self.d_opt))
            out.append(m.latent_discriminator_training_step(real, synth, self.d_opt) if c == 4 else
                       m.latent_discriminator_training_step(synth, self.d_opt))
        out.append(m.generator_training_step(real, synth, self.g_opt))
        m.update_smoothed_weights()
        return out

    def units(self):
        return units_per_step(self.cfg_id, self.b * self.world)

    def close(self):
        self.model.close()
        if self.cfg_id == 5:
            self.gan.close()


def device_results(results):
    """the device tensors among a step's results, flattened (loss scalars, uint8 images)"""
    vals = []
    for r in results:
        if isinstance(r, dict):
            vals += [v.detach().reshape(-1).float() for v in r.values()]
        elif isinstance(r, torch.Tensor):
            vals.append(r.detach().reshape(-1).float() if r.dtype != torch.uint8 else r.detach().reshape(-1))
    return vals


def host_bytes(results):
    return sum(r.nbytes for r in results if isinstance(r, np.ndarray))


def run_b200(args):
    import torch.distributed as dist
    from confignet_b200 import ops, _lib as L
    from confignet_b200 import runtime
    from confignet_b200.runtime import GraphedFn

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
    cfg_id = args.config
    wl = Workload(cfg_id, dev, world, rank)

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

    sampler = ClockSampler(local)
    if rank == 0:                  # one poller per job
        sampler.start()
    for _ in range(args.warmup):
        wl.step(True)
    barrier()

    # ---- value: stores resident in HBM, CUDA events around the K steps
    # roofline: a CUDA-event pair around every conv-family launch over extra EAGER steps right after the timed region
    # (default; a graph replay cannot carry them), or - with --roofline-pass inline - inside it (forces eager execution)
    inline = args.roofline_pass == "inline"
    prof = []
    if inline:
        ops.PROFILE[0] = prof
    lib.cn_launch_count(1)
    replayed0 = GraphedFn.REPLAYED_LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.begin()
    e0.record()
    marks, host_ms = [], []
    for _ in range(args.steps):
        th = time.perf_counter()
        wl.step(True)
        host_ms.append((time.perf_counter() - th) * 1e3)
        if args.debug_steps:
            ev = torch.cuda.Event(enable_timing=True); ev.record(); marks.append(ev)
    e1.record()
    barrier()
    if args.debug_steps and rank == 0:
        prev, per = e0, []
        for ev in marks:
            per.append(prev.elapsed_time(ev)); prev = ev
        print("value loop: device ms per step %s | host ms per step %s" % (["%.1f" % x for x in per], ["%.1f" % x for x in host_ms]),
              file=sys.stderr, flush=True)
    launches = int(lib.cn_launch_count(0)) + (GraphedFn.REPLAYED_LAUNCHES - replayed0)     # eager launches + launches replayed from CUDA graphs
    ops.PROFILE[0] = None
    ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    clocks = sampler.summary()
    value = wl.units() / (ms * 1e-3)
    latency = None
    if cfg_id == 1:                 # the latency of ONE call, each timed on its own (median of 5)
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record(); wl.step(True); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        latency = sorted(ts)[2]
    prof_steps = args.steps
    if not inline:
        ops.PROFILE[0] = prof
        prof_steps = max(1, min(args.steps, 4 if cfg_id != 5 else 1))
        for _ in range(prof_steps):
            wl.step(True)
        barrier()
        ops.PROFILE[0] = None

    # ---- roofline of the tcgen05 conv kernels (events recorded around each launch)
    tc_exec = sum(p[6] for p in prof if p[4] == 2)
    tc_flops = sum(p[1] for p in prof if p[4] == 2)
    tc_ms = sum(p[2].elapsed_time(p[3]) for p in prof if p[4] == 2)
    cc_ms = sum(p[2].elapsed_time(p[3]) for p in prof if p[4] != 2)
    all_flops = sum(p[1] for p in prof)
    hbm, tensor_peak, how = measured_peaks()
    achieved = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "tc_dram_traffic.json")      # written by scripts/summarize_launches.py from ncu (config 2)
    if cfg_id == 2 and os.path.exists(tpath):
        with open(tpath) as fp:
            t = json.load(fp)
        traffic, traffic_src = t.get("dram_bytes_per_launch"), t.get("source")
    n_tc = sum(1 for p in prof if p[4] == 2)

    def conv_bytes(key):
        """algorithmic bytes of one conv launch: input + output + kernel, each moved once (fp32)"""
        nd, batch, in_dims, cin, cout, ksize, stride, up = key[:8]      # a ninth entry = explicit padding (ResNet stem)
        n_in = n_out = batch
        for d_ in in_dims:
            n_in *= d_
            n_out *= -(-(d_ * up) // stride)
        kvol = 1
        for k_ in ksize:
            kvol *= k_
        return 4.0 * (n_in * cin + n_out * cout + kvol * cin * cout)
    tc_bytes = sum(conv_bytes(p[5]) for p in prof if p[4] == 2)
    tf32 = None
    ppath = os.path.join(ROOT, "profiles", "tf32_mma_peak.json")        # measured by scripts/gpu_probe_round2.py (MMA-only loop)
    if os.path.exists(ppath):
        with open(ppath) as fp:
            tf32 = json.load(fp)
    roofline = {"bound": "tensor", "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
                "frac": achieved / tensor_peak, "traffic": traffic, "traffic_source": traffic_src,
                "traffic_algorithmic": tc_bytes / max(n_tc, 1),
                "traffic_note": "algorithmic bytes per launch = fp32 input + output + kernel of the layer, each once; the ncu "
                                "figure is below it because the tail of every output is still in the 126 MB L2 when its kernel ends",
                "peak_source": how + " cuBLAS bf16 dense (sustained) - the contract's denominator.  The measured kind::tf32 "
                               "issue peak is in tf32_mma_peak: 1116 TFLOP/s with 256-column tiles, 934 with the 128-column tiles "
                               "of this kernel (75 clk per MMA for any N <= 128); every fp32 product takes 3 tf32 MMAs (3xTF32), "
                               "so the ceiling on executed FLOPs is 311 TFLOP/s = 0.23 of this peak",
                "tf32_mma_peak": tf32,
                "kernel": "igemm_tc_pixel_kernel<B_MN,WG> (tcgen05 kind::tf32 fwd/dgrad/wgrad, 3 MMAs per product)",
                "algorithmic_gflop_per_launch": tc_flops / max(n_tc, 1) / 1e9,
                "avg_launch_us": tc_ms * 1e3 / max(n_tc, 1),
                "executed_tflops": tc_exec / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0,
                "executed_note": "FLOPs executed after sub-pixel folding of UpSampling(2)+conv; achieved counts the "
                                 "reference formulation (convs on the upsampled grid)",
                "launches": n_tc // prof_steps,
                "tc_ms_per_step": tc_ms / prof_steps, "cuda_core_conv_ms_per_step": cc_ms / prof_steps,
                "algorithmic_conv_tflop_per_step": all_flops / prof_steps / 1e12,
                "step_algorithmic_tflops": all_flops / prof_steps / 1e12 / (ms * 1e-3),
                "measured": ("CUDA events around each conv launch inside the timed region" if inline else
                             "CUDA events around each conv launch over %d extra eager step(s) right after the timed region" % prof_steps)}

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
        # the captured step graphs hold no NCCL work (the all-reduce runs between two graphs): release them, then a
        # normal teardown of the process group
        wl.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()

    if args.no_e2e:
        if rank == 0:
            print(json.dumps({"value": value, "ms_per_step": ms, "note": "profiling run (no e2e leg)"}), flush=True)
        finish()
        return
    # ---- e2e: public API, host (NumPy) inputs, H2D + D2H inside the timed region
    for _ in range(2):
        r = wl.step(False)
        [v.cpu() for v in device_results(r)]
    d2h = 0
    # Every step's results are read back inside the timed region, but without draining the GPU: the D2H copy into a
    # pinned buffer is enqueued right behind the step, and the host looks at step k-1's buffer while step k runs.
    # (Results the API itself returns as NumPy - generate_images, fine_tune_on_img - have already crossed.)
    probe = device_results(wl.step(False))
    pinned = None
    if probe:
        n_res = sum(v.numel() for v in probe)
        dt_ = probe[0].dtype if all(v.dtype == probe[0].dtype for v in probe) else torch.float32
        pinned = [torch.zeros(n_res, dtype=dt_).pin_memory() for _ in range(2)]

    def enqueue_readback(i, results):
        vals = device_results(results)
        if not vals:
            return None
        flat = torch.cat([v.to(pinned[0].dtype) for v in vals])
        pinned[i % 2].copy_(flat, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return pinned[i % 2], ev

    runtime.H2D_BYTES[0] = 0
    barrier()
    t0 = time.perf_counter()
    pending = None
    ev0 = torch.cuda.Event(enable_timing=True); ev0.record()
    marks, host_ms = [], []
    for i in range(args.steps):
        th = time.perf_counter()
        cur = wl.step(False)              # host halves (sampling, pinned uploads) + asynchronous device halves
        host_ms.append((time.perf_counter() - th) * 1e3)
        if args.debug_steps:
            ev = torch.cuda.Event(enable_timing=True); ev.record(); marks.append(ev)
        d2h += host_bytes(cur)
        rb = enqueue_readback(i, cur)
        if pending is not None:
            pending[1].synchronize()
            d2h += pending[0].numpy().copy().nbytes
        pending = rb
    if pending is not None:
        pending[1].synchronize()
        d2h += pending[0].numpy().copy().nbytes
    torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0) / args.steps
    if args.debug_steps and rank == 0:
        prev, per = ev0, []
        for ev in marks:
            per.append(prev.elapsed_time(ev)); prev = ev
        print("e2e loop: device ms per step %s | host ms per step %s" % (["%.1f" % x for x in per], ["%.1f" % x for x in host_ms]),
              file=sys.stderr, flush=True)
    e2e = {"value": wl.units() / dt, "unit": CONFIGS[cfg_id]["unit"],
           "h2d_bytes_per_step": runtime.H2D_BYTES[0] // args.steps, "d2h_bytes_per_step": d2h // args.steps}
    if cfg_id == 1:
        ts = []
        for _ in range(5):
            t1 = time.perf_counter(); wl.step(False); ts.append((time.perf_counter() - t1) * 1e3)
        e2e["latency_ms_median_of_5"] = sorted(ts)[2]

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_baseline_leg(cfg_id)
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": CONFIGS[cfg_id]["unit"], "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": bench_config(cfg_id, world),
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
                "cpu_baseline": cpu_baseline}
        if latency is not None:
            line["latency_ms_median_of_5"] = latency
        print(json.dumps(line), flush=True)
    finish()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json configuration (1-based), default 2")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the end-to-end leg")
    ap.add_argument("--roofline-pass", default="separate", choices=["inline", "separate"],
                    help="separate (default): the timed region runs the product path as is (CUDA-graph replays), the "
                         "per-launch event pairs of the roofline run over extra eager steps right after it; inline: event "
                         "pairs inside the timed region (forces eager execution)")
    ap.add_argument("--breakdown", default=None, help="write a per-layer conv time table to this file")
    ap.add_argument("--debug-steps", action="store_true", help="per-step device and host times of both timed loops on stderr")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a GPU (the CUDA path has no CPU fallback)")
        run_b200(args)


if __name__ == "__main__":
    main()
