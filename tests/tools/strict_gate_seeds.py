"""End-to-end strict gate (max-abs <= 1e-3 vs the fp64 oracle) over several weight seeds, FFMA GEMMs.  Shows how often isolated
chaotic FDSA events push an fp32 evaluation order over 1e-3, for the fast and the generic FFT kernels.
    FDN_FFT_FAST=0|1 python tests/tools/strict_gate_seeds.py"""
import os
import sys

os.environ["FDN_B200_GEMM"] = "ffma"
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _ROOT)
sys.path.insert(0, os.path.join(_ROOT, "tests"))
import torch
import parity_cases as P

dev = torch.device("cuda")
# FFT accuracy on the pyramid of the failing case
from fdn_tip2025_b200 import ops
for h, w in ((128, 160), (64, 80), (32, 40), (640, 1120)):
    x = P.rnd(3, h, w, seed=h + w)
    x[0] += 3.0
    wf = w // 2 + 1
    spec = torch.empty(3, h, wf, 2, device=dev)
    ops.fft_rows_r2c(x.float().to(dev), spec)
    ops.fft_cols(spec, h * wf, wf, spec, h * wf, wf, 3, h, wf, w, ops.COLS_FWD)
    ref = torch.fft.rfft2(x.float().double())
    got = torch.view_as_complex(spec.cpu().contiguous()).to(torch.complex128)
    print("rfft2 %dx%d rel-L2 %.3e max/max %.3e" % (h, w, ((got - ref).abs().pow(2).sum() / ref.abs().pow(2).sum()).sqrt().item(),
                                                     (got - ref).abs().max().item() / ref.abs().max().item()), flush=True)
for seed in (7, 8, 9, 10, 11, 12):
    rep = []
    try:
        P.case_fdn(dev, "FDN", 128, 160, b=2, report=rep, seed=seed, strict=False, damp=float(os.environ.get("DAMP", "0.03")))
    except AssertionError as e:
        print("seed", seed, "loose gate failed:", e)
    print("FAST=%s seed %d: max-abs %.3e PSNR %.1f dB frac>1e-3 %.2e" % (os.environ.get("FDN_FFT_FAST", "1"), seed, rep[0][1], rep[0][2], rep[0][3]), flush=True)
