"""Does the per-row cost of the tcgen05 1x1 conv depend on how far apart the channel planes are?  Same pixel count, K and N, but planes of
2.9 MB (4 images of 640x1120) versus planes of 16 KB (700 images of 64x64, the K rows of a tile within one 2 MB page).  Dev tool, GPU only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fdn_tip2025_b200 import ops, packing

dev = "cuda"
for k, n in ((86, 32), (172, 64), (32, 152)):
    for b, h, w in ((4, 640, 1120), (700, 64, 64), (175, 128, 128)):
        x = torch.randn(b, k, h, w, device=dev)
        wgt = torch.randn(n, k, device=dev) / k ** 0.5
        packed = packing.pack_weight(wgt)
        out = torch.empty(b, n, h, w, device=dev)
        res = torch.randn(b, n, h, w, device=dev)
        fn = lambda: ops.pw_mma([x], packed, out, res=res, passes=3)
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        tiles = b * ((h * w + 127) // 128)
        cyc_per_row = ms * 1e-3 * 1.965e9 / (tiles / 148.0) / k
        print("K=%3d N=%3d  B=%3d planes of %7.1f KB: %.3f ms  %.0f GB/s  %.0f cycles per loaded row" % (
            k, n, b, h * w * 4 / 1024, ms, (k + 2 * n) * b * h * w * 4 / ms / 1e6, cyc_per_row), flush=True)
