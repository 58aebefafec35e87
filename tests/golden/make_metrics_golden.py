"""Generate tests/golden/metrics_golden.pt from the reference's own basicsr/metrics/psnr_ssim.py (build container only).

psnr_ssim.py is loaded by path; its imports that are absent here (skimage, the basicsr package chain) are stubbed with the
reference's own metric_util / matlab_functions loaded by path, and Tensor.cuda / Module.cuda are made identities because
_ssim_3d moves its operands to a GPU (psnr_ssim.py:178-182).  Stores seeded image pairs' metric values for every mode.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("FDN_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference_metrics():
    for pkg in ("basicsr", "basicsr.utils", "basicsr.metrics", "skimage"):
        sys.modules.setdefault(pkg, types.ModuleType(pkg))
    sk = types.ModuleType("skimage.metrics")
    sk.structural_similarity = None
    sys.modules["skimage.metrics"] = sk
    _load("basicsr.utils.matlab_functions", os.path.join(REF, "basicsr/utils/matlab_functions.py"))
    _load("basicsr.metrics.metric_util", os.path.join(REF, "basicsr/metrics/metric_util.py"))
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    return _load("basicsr.metrics.psnr_ssim", os.path.join(REF, "basicsr/metrics/psnr_ssim.py"))


def image_pair(seed, h, w, scale):
    """A smooth frame and a degraded copy, HWC float64 in [0, scale]."""
    g = torch.Generator().manual_seed(seed)
    base = torch.nn.functional.interpolate(torch.rand(1, 3, h // 8 + 1, w // 8 + 1, generator=g), size=(h, w), mode="bicubic", align_corners=False)
    a = base.clamp(0, 1)
    b = (a + 0.05 * (torch.rand(1, 3, h, w, generator=g) - 0.5) + 0.02).clamp(0, 1)
    a, b = a[0].permute(1, 2, 0).double().numpy(), b[0].permute(1, 2, 0).double().numpy()
    if scale == 255:
        a, b = np.round(a * 255.), np.round(b * 255.)
    return a, b


CASES = [(11, 48, 64, 255, 0), (12, 40, 56, 1, 0), (13, 64, 48, 255, 4), (14, 33, 47, 1, 2)]


def main():
    M = load_reference_metrics()
    out = []
    for seed, h, w, scale, crop in CASES:
        a, b = image_pair(seed, h, w, scale)
        rec = {"seed": seed, "h": h, "w": w, "scale": scale, "crop": crop,
               "psnr": float(M.calculate_psnr(a, b, crop)), "psnr_y": float(M.calculate_psnr(a, b, crop, test_y_channel=True)),
               "ssim3d": float(M.calculate_ssim(a, b, crop)), "ssim2d": float(M.calculate_ssim(a, b, crop, ssim3d=False)),
               "ssim_y": float(M.calculate_ssim(a, b, crop, test_y_channel=True))}
        print(rec)
        out.append(rec)
    torch.save(out, os.path.join(HERE, "metrics_golden.pt"))


if __name__ == "__main__":
    main()
