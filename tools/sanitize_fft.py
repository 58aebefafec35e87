"""Small FFT-stage run for compute-sanitizer (memcheck / racecheck / synccheck): the TMA-staged persistent row kernels, the column
kernels and the packed depthwise kernels on sizes that take seconds under the tool.  Dev tool, GPU only.

    compute-sanitizer --tool racecheck python tools/sanitize_fft.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_cases as P

dev = "cuda"
for h, w in ((64, 64), (160, 280), (128, 128)):
    P.case_rfft2_irfft2(dev, h, w, planes=2)
    P.case_fcaffn_fft_stage(dev, h, w, b=1, c=2)
    P.case_irfft2_nonhermitian(dev, h, w)
P.case_rfft2_irfft2(dev, 46, 94, planes=1)          # prime radices first
P.case_tblock(dev, 32, 16, 24, True, True)          # depthwise gate pair / GELU / FiLM kernels on packed fp32x2
print("sanitize_fft: ok")
