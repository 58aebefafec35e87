"""Generate tests/golden/losses_golden.pt from the reference's own basicsr/models/losses/losses.py (build container only).

losses.py and loss_util.py are loaded by path (the basicsr package itself does not import here).
MARLoss is called with a stand-in vgg_loss that returns zero, so the stored value is mse + 0.01 * amplitude mse.
"""
import importlib.util
import os
import sys
import types

import torch

REF = os.environ.get("FDN_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference_losses():
    for pkg in ("basicsr", "basicsr.models", "basicsr.models.losses"):
        sys.modules.setdefault(pkg, types.ModuleType(pkg))
    _load("basicsr.models.losses.loss_util", os.path.join(REF, "basicsr/models/losses/loss_util.py"))
    return _load("basicsr.models.losses.losses", os.path.join(REF, "basicsr/models/losses/losses.py"))


def pair(seed, b, c, h, w):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(b, c, h, w, generator=g), torch.rand(b, c, h, w, generator=g)


FFT_CASES = [(1, 2, 3, 32, 48), (2, 1, 3, 128, 160), (3, 1, 3, 104, 152)]
MAR_CASES = [(4, 2, 3, 128, 192), (5, 1, 3, 256, 256)]


def main():
    L = load_reference_losses()
    out = {"fft": [], "mar": []}
    for seed, b, c, h, w in FFT_CASES:
        p, t = pair(seed, b, c, h, w)
        rec = {"seed": seed, "shape": (b, c, h, w), "mean": float(L.FFTLoss()(p, t)), "sum_w": float(L.FFTLoss(0.1, "sum")(p, t))}
        print(rec)
        out["fft"].append(rec)
    zero_vgg = lambda a, b: (torch.zeros(()), None)
    for seed, b, c, h, w in MAR_CASES:
        y, _ = pair(seed, b, c, h, w)
        x, _ = pair(seed + 100, b, c, h // 8, w // 8)
        rec = {"seed": seed, "shape": (b, c, h, w), "loss": float(L.MARLoss()(x, y, zero_vgg))}
        print(rec)
        out["mar"].append(rec)
    torch.save(out, os.path.join(HERE, "losses_golden.pt"))


if __name__ == "__main__":
    main()
