"""Validation metrics: the numpy oracle against the reference's own numbers (CPU), the emulated kernel
against the oracle (CPU), and the CUDA kernels against both (GPU)."""
import os
import sys

import pytest
import torch

import parity_cases as P

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics_golden.pt")


def test_metrics_oracle_matches_reference_outputs():
    sys.path.insert(0, os.path.dirname(GOLDEN))
    import make_metrics_golden as G
    from oracle import metrics_oracle as MO
    for rec in torch.load(GOLDEN):
        a, b = G.image_pair(rec["seed"], rec["h"], rec["w"], rec["scale"])
        c = rec["crop"]
        assert MO.psnr(a, b, c) == pytest.approx(rec["psnr"], abs=1e-12)
        assert MO.psnr(a, b, c, True) == pytest.approx(rec["psnr_y"], abs=1e-12)
        assert MO.ssim(a, b, c, ssim3d=False) == pytest.approx(rec["ssim2d"], abs=1e-12)
        assert MO.ssim(a, b, c, test_y_channel=True) == pytest.approx(rec["ssim_y"], abs=1e-12)
        # the reference evaluates its 3-D Gaussian in float32 on the GPU path (psnr_ssim.py:178-182); the oracle keeps float64
        assert MO.ssim(a, b, c) == pytest.approx(rec["ssim3d"], abs=5e-6)
    assert MO.psnr(a, a, 0) == float("inf") and MO.ssim(a, a, 0) == pytest.approx(1.0, abs=1e-12)


def test_metrics_kernels_on_the_emulator():
    from emu import harness
    harness.enable()
    try:
        P.case_metrics("cpu", sizes=((24, 40),), golden=torch.load(GOLDEN)[:1])
    finally:
        from fdn_tip2025_b200 import _lib, ops
        import importlib
        _lib._handle = None
        importlib.reload(ops)
        harness._enabled = False


@pytest.mark.gpu
def test_metrics_kernels(cuda_dev):
    P.case_metrics(cuda_dev, sizes=((48, 64), (33, 47), (640, 1120)), golden=torch.load(GOLDEN))


@pytest.mark.gpu
def test_metrics_identical_frames(cuda_dev):
    from fdn_tip2025_b200 import ops, synth
    x = synth.low_light_images(2, 64, 96).to(cuda_dev)
    assert torch.isinf(ops.psnr(x, x)).all()
    assert torch.allclose(ops.ssim(x, x), torch.ones(2, dtype=torch.float64, device=cuda_dev), atol=1e-12)
