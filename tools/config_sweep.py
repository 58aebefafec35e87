"""Throughput of every BASELINE.json configuration on one B200 (inputs resident in HBM, CUDA events).  GPU only.

    python tools/config_sweep.py [out.json]

configs[1] FDN_lolv1 416x608 batch 8; configs[2] FDN 640x1120 (8 images per GPU); configs[3] I_predict_net 640x1120 batch 8;
configs[4] FDN 2176x3840 single frame.  configs[0] is the CPU case (tests/test_oracle_golden.py).
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fdn_tip2025_b200 import archs, synth


def timed(fn, n):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


rows = []
lp = archs.I_predict_net()
lp.load_state_dict(synth.lpnet_state_dict(seed=3), strict=True)
lp = lp.cuda().eval()
for name, cls, dim, b, h, w, reps in (("configs[1] FDN_lolv1 416x608 b8", archs.FDN_lolv1, 24, 8, 416, 608, 3),
                                      ("configs[2] FDN 640x1120 b8", archs.FDN, 32, 8, 640, 1120, 2),
                                      ("configs[4] FDN 2176x3840 b1", archs.FDN, 32, 1, 2176, 3840, 1)):
    net = cls()
    net.load_state_dict(synth.fdn_state_dict(dim=dim, seed=0, damp=0.03), strict=True)
    net = net.cuda().eval()
    x = synth.low_light_images(b, h, w).cuda()
    ratio = lp(x)
    ms = timed(lambda: net(x, ratio_i=ratio), reps)
    out = net(x, ratio_i=ratio)[0]
    rows.append({"config": name, "ms_per_step": ms, "images_per_s": b / ms * 1e3, "finite": bool(torch.isfinite(out).all()),
                 "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9})
    print(json.dumps(rows[-1]), flush=True)
    del net, x, out
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
x = synth.low_light_images(8, 640, 1120).cuda()
ms = timed(lambda: lp(x), 10)
rows.append({"config": "configs[3] I_predict_net 640x1120 b8", "ms_per_step": ms, "images_per_s": 8 / ms * 1e3})
print(json.dumps(rows[-1]), flush=True)
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
