"""Frames/s of the device-resident inference loop with and without CUDA-graph replay, per frame size.  GPU only."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fdn_tip2025_b200 import archs, pipeline, synth

net = archs.FDN()
net.load_state_dict(synth.fdn_state_dict(dim=32, seed=0, damp=0.03), strict=True)
net = net.cuda().eval()
lp = archs.I_predict_net()
lp.load_state_dict(synth.lpnet_state_dict(seed=3), strict=True)
lp = lp.cuda().eval()
for h, w in ((256, 256), (400, 600), (640, 1120)):
    frames = (synth.low_light_images(1, h, w) * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous().cuda()
    res = {}
    for name, graphs in (("eager", False), ("graph", True)):
        pipe = pipeline.InferencePipeline(net, lp, "lolblur", use_graphs=graphs)
        fn = pipe.run_device_graphed if graphs else pipe.run_device
        for _ in range(3):
            fn(frames)
        torch.cuda.synchronize()
        n = 10
        t0 = time.perf_counter()
        for _ in range(n):
            fn(frames)
        torch.cuda.synchronize()
        res[name] = n / (time.perf_counter() - t0)
    print("%dx%d: eager %.1f frames/s, CUDA graph %.1f frames/s (x%.2f)" % (w, h, res["eager"], res["graph"], res["graph"] / res["eager"]), flush=True)
