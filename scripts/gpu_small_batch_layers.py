"""Forward time of the generator's conv layers at generate_images batch sizes (CUDA events, warm, 20 calls each), the kernel
family that ran and the launches per call (2 = pack or split-K reduce beside the GEMM).   python scripts/gpu_small_batch_layers.py [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from confignet_b200 import ops, _lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda:0")
LAYERS = [("map_3d_0", (4, 4, 4), 512, 256, 3, 2), ("map_3d_1", (8, 8, 8), 256, 128, 3, 2), ("conv3d 128->64", (16, 16, 16), 128, 64, 3, 1),
          ("conv3d 64->64", (16, 16, 16), 64, 64, 3, 1), ("1x1 1024->512", (16, 16), 1024, 512, 1, 1), ("map_2d_0", (16, 16), 512, 256, 4, 1),
          ("map_2d_1", (16, 16), 256, 64, 4, 2), ("map_2d_2", (32, 32), 64, 32, 4, 2), ("map_2d_2b", (64, 64), 32, 32, 4, 2)]
lib = L.load()
print("batch %d" % B)
for name, dims, cin, cout, k, up in LAYERS:
    x = torch.randn(B, *dims, cin, device=dev)
    w = torch.randn(*([k] * len(dims)), cin, cout, device=dev) * 0.05
    b = torch.randn(cout, device=dev)
    with torch.no_grad():
        for _ in range(3):
            ops.conv_act(x, w, b, upsample=up)
        torch.cuda.synchronize()
        n0 = int(lib.cn_launch_count(0))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.conv_act(x, w, b, upsample=up)
        e1.record()
        torch.cuda.synchronize()
        n1 = int(lib.cn_launch_count(0))
    print("%-16s in %-13s %4d -> %3d k%d up%d: %7.1f us per call, kernel family %d, %.1f launches per call"
          % (name, dims, cin, cout, k, up, e0.elapsed_time(e1) / 20 * 1e3, lib.cn_last_conv_impl(), (n1 - n0) / 20))
