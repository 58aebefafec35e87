"""The oracle (torch restatement of the reference forward) run with PyTorch eager kernels on the GPU: the like-for-like incumbent the
reference itself would be on a B200 (cuDNN / cuFFT / ATen elementwise).  Checker-side diagnostic, GPU only.
    python tests/tools/eager_port_on_gpu.py [H W]"""
import os
import sys
import time

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT)
import torch
from fdn_tip2025_b200 import archs, synth
from oracle import fdn_oracle as O

H = int(sys.argv[1]) if len(sys.argv) > 1 else 640
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1120
dev = torch.device("cuda")
sd = synth.fdn_state_dict(dim=32, seed=0, damp=0.03)
sdg = {k: v.to(dev) for k, v in sd.items()}
x = synth.low_light_images(1, H, W).to(dev)
ratio = torch.full((1, 1), 0.35, device=dev)
with torch.no_grad():
    for _ in range(2):
        ref = O.fdn(x, ratio, sdg, "lolblur")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 3
    for _ in range(n):
        ref = O.fdn(x, ratio, sdg, "lolblur")
    torch.cuda.synchronize()
    t_eager = (time.perf_counter() - t0) / n
net = archs.FDN()
net.load_state_dict(sd, strict=True)
net = net.to(dev).eval()
for _ in range(2):
    out = net(x, ratio_i=ratio)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(n):
    out = net(x, ratio_i=ratio)
torch.cuda.synchronize()
t_ours = (time.perf_counter() - t0) / n
d = (out[0] - ref[0]).abs()
print("FDN %dx%d, 1 image: PyTorch eager port %.1f ms (%.2f img/s, peak mem %.1f GB) | libfdn_b200 %.1f ms (%.2f img/s) | speed-up %.1fx | max-abs diff %.2e, PSNR %.1f dB"
      % (W, H, t_eager * 1e3, 1 / t_eager, torch.cuda.max_memory_allocated() / 1e9, t_ours * 1e3, 1 / t_ours, t_eager / t_ours, d.max().item(),
         O.psnr(out[0].cpu(), ref[0].cpu())))
