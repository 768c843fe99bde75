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
