"""Measures the kind::tf32 tcgen05.mma issue peak on this B200 with an MMA-only loop (no loads; probe 5 of
confignet_b200/csrc/experiments/round2_probes.cu) and writes profiles/tf32_mma_peak.json - the denominator VERDICT r01
asked for next to the cuBLAS bf16 anchor of MEASURED_PEAKS.json.

    gpurun --timeout 300 -- 'python scripts/gpu_tf32_peak.py'

One CTA per SM (148), one issuing thread each; per (N, mode): dense TFLOP/s = grid * iters * mmas * 2*128*N*8 / launch time,
and the median clocks per MMA seen by the issuing thread.  modes: ts = A from tensor memory, ss = A from shared memory,
3xTF32 mixes as issued by the production kernel (all TS) and by the TMA-fed candidate (1 TS + 2 SS)."""
import ctypes
import json
import os
import subprocess
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "confignet_b200", "lib", "libcn_probes.so"))
lib.probe_mma_peak.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_void_p]


def smi():
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip().split(",")
        return [float(x) for x in out]
    except Exception:
        return None


def main():
    dev = torch.device("cuda:0")
    torch.zeros(1, device=dev)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    clk = torch.zeros(sms, dtype=torch.int64, device=dev)
    rows = []
    print("%-4s %-22s %10s %12s %14s" % ("N", "mode", "ms", "TFLOP/s", "clk per MMA"))
    for n in (64, 96, 128, 256):
        for mode, name, per in ((0, "ts", 1), (1, "ss", 1), (2, "3x: 3 TS (production)", 3), (3, "3x: 1 TS + 2 SS (TMA)", 3)):
            iters = 60000 // per if n <= 128 else 30000 // per
            ms = ctypes.c_float(0)
            rc = lib.probe_mma_peak(n, iters, mode, sms, ctypes.byref(ms), ctypes.c_void_p(clk.data_ptr()))
            if rc != 0:
                print("probe_mma_peak(%d, mode %d) -> %d" % (n, mode, rc)); continue
            mmas = iters * 4 * per
            tf = sms * mmas * 2.0 * 128 * n * 8 / (ms.value * 1e-3) / 1e12
            cpm = float(np.median(clk.cpu().numpy())) / mmas
            rows.append(dict(n=n, mode=name, ms=ms.value, dense_tf32_tflops=tf, clocks_per_mma=cpm, smi=smi()))
            print("%-4d %-22s %10.3f %12.1f %14.1f" % (n, name, ms.value, tf, cpm), flush=True)
    best = max(r["dense_tf32_tflops"] for r in rows)
    peaks = {}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fp:
            peaks = json.load(fp)
    out = {"how": "MMA-only loop, tcgen05.mma.cta_group::1.kind::tf32 M=128, one issuing thread per CTA, one CTA per SM "
                  "(scripts/gpu_tf32_peak.py, csrc/experiments/round2_probes.cu probe 5); burst figure (tens of ms per launch)",
           "dense_tf32_tflops_best": best, "three_x_tf32_ceiling_tflops": best / 3.0,
           "bf16_tflops_measured_cublas": peaks.get("bf16_tflops"), "bf16_tflops_sustained_measured_cublas": peaks.get("bf16_tflops_sustained"),
           "rows": rows}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "tf32_mma_peak.json"), "w") as fp:
        json.dump(out, fp, indent=1)
    print("best dense tf32 %.1f TFLOP/s -> 3xTF32 ceiling %.1f TFLOP/s" % (best, best / 3))


if __name__ == "__main__":
    main()
