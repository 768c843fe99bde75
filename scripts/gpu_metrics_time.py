"""Times the two metric networks on the B200 (CUDA events, inputs resident in HBM, warm): InceptionV3 features and the
MobileNetV2 attribute classifier, images per second and the conv kernel family each layer ran on.
    python scripts/gpu_metrics_time.py [batch]"""
import os
import sys
import warnings
from collections import Counter

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from confignet_b200 import ops, _lib as L                                   # noqa: E402
from confignet_b200.metrics import InceptionFeatureExtractor, CelebaAttributeClassifier, nets   # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
warnings.simplefilter("ignore")
dev = torch.device("cuda:0")
imgs = torch.randint(0, 256, (B, 256, 256, 3), dtype=torch.uint8, device=dev)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def breakdown(fn):
    ops.PROFILE[0] = []
    fn()
    torch.cuda.synchronize()
    prof, ops.PROFILE[0] = ops.PROFILE[0], None
    impl = Counter()
    ms = Counter()
    flops = 0.0
    for op, fl, e0, e1, im, key, ex in prof:
        impl[im] += 1
        ms[im] += e0.elapsed_time(e1)
        flops += fl
    return dict(impl), {k: round(v, 3) for k, v in ms.items()}, flops


ex = InceptionFeatureExtractor((256, 256, 3), device=dev)
t = timed(lambda: ex.features_device(imgs))
impl, ms, fl = breakdown(lambda: ex.features_device(imgs))
print("InceptionV3 features: batch %d, %.2f ms, %.0f images/s, %.1f GFLOP/image, %.1f TFLOP/s; conv launches by family (2 = tcgen05, 1 = CUDA-core) %s, conv ms by family %s"
      % (B, t, B / t * 1e3, fl / B / 1e9, fl / t / 1e9, impl, ms))
# the CPU restatement beside it (oracle, fp32, all host cores, a bounded sample of the same workload)
import time                                                                    # noqa: E402
from oracle import metrics_oracle as MO, confignet_oracle as O                  # noqa: E402
ncpu = 8
xc = torch.tensor(imgs[:ncpu].cpu().numpy().astype(np.float32) / np.float32(127.5) - np.float32(1))
pc = O.to_torch(ex.raw_weights, dtype=torch.float32)
with torch.no_grad():
    MO.inception_v3_features(pc, xc[:2])
    t0 = time.perf_counter()
    MO.inception_v3_features(pc, xc)
    tc = time.perf_counter() - t0
print("InceptionV3 features, CPU oracle (torch fp32, %d threads): %d images in %.2f s = %.1f images/s" % (torch.get_num_threads(), ncpu, tc, ncpu / tc))
clf = CelebaAttributeClassifier({"input_shape": [128, 128, 3], "predicted_attributes": ["a%d" % i for i in range(40)]}, device=dev)
t = timed(lambda: clf.predict_attributes(imgs))
x = ops.from_uint8(nets.resize_images(imgs, 128, 128))
impl, ms, fl = breakdown(lambda: clf.predict_device(x))
print("MobileNetV2 attribute classifier (256 -> 128 resize on the device): batch %d, %.2f ms, %.0f images/s; conv launches by family %s, conv ms by family %s"
      % (B, t, B / t * 1e3, impl, ms))

# ---- the non-conv kernels of csrc/metrics.cu against the HBM roofline: CUDA events, an L2 flush (256 MB write) before every launch
PEAK = 6547.0
try:
    import json
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as fp:
        mp = json.load(fp)
    PEAK = float(mp.get("hbm_gbs", PEAK))
except Exception:
    pass
flush = torch.empty(64 * 1024 * 1024, device=dev, dtype=torch.float32)


def hbm(name, fn, nbytes, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print("%-64s %8.1f us  %7.1f MB algorithmic  %6.0f GB/s  %.2f of %.0f GB/s" % (name, t * 1e3, nbytes / 1e6, nbytes / t / 1e6, nbytes / t / 1e6 / PEAK, PEAK))


print("\nmetric-network kernels against the HBM roofline (algorithmic bytes = fp32 input + output; L2 flushed before every launch; median of 5)")
x = torch.randn(B, 35, 35, 288, device=dev)
hbm("cn_pool2d_fwd avg 3x3/s1 SAME  %dx35x35x288 (mixed1/2 pool branch)" % B, lambda: nets.pool2d(x, 3, 1, True, L.POOL_AVG_VALID), 2 * x.numel() * 4)
x2 = torch.randn(B, 61, 61, 192, device=dev)
hbm("cn_pool2d_fwd max 3x3/s2 VALID %dx61x61x192 (stem)" % B, lambda: nets.pool2d(x2, 3, 2, False, L.POOL_MAX), (x2.numel() + B * 30 * 30 * 192) * 4)
x3 = torch.randn(B, 64, 64, 96, device=dev)
wk, bb = torch.randn(3, 3, 96, device=dev), torch.randn(96, device=dev)
hbm("cn_dwconv3x3_fwd s2 + bias + ReLU6 %dx64x64x96 (block_1)" % B, lambda: nets.dwconv3x3(x3, wk, bb, 2, L.ACT_RELU6), (x3.numel() + B * 32 * 32 * 96) * 4)
x4 = torch.randn(B, 32, 32, 144, device=dev)
wk4, bb4 = torch.randn(3, 3, 144, device=dev), torch.randn(144, device=dev)
hbm("cn_dwconv3x3_fwd s1 + bias + ReLU6 %dx32x32x144 (block_2)" % B, lambda: nets.dwconv3x3(x4, wk4, bb4, 1, L.ACT_RELU6), 2 * x4.numel() * 4)
hbm("cn_act_ext ReLU6 clamp in place %dx64x64x96" % B, lambda: nets.act_ext_(x3, L.ACT_RELU6), 2 * x3.numel() * 4)
hbm("cn_resize_bilinear uint8 256 -> 128, %d images" % B, lambda: nets.resize_images(imgs, 128, 128), imgs.numel() + B * 128 * 128 * 3)
hbm("cn_from_uint8 (preprocess_input) %d x 256 x 256 x 3" % B, lambda: ops.from_uint8(imgs), imgs.numel() * 5)
