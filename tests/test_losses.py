"""Spectral losses: oracle vs the reference's own outputs (CPU); CUDA forward vs oracle and golden (GPU)."""
import os
import sys

import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "losses_golden.pt")
sys.path.insert(0, os.path.dirname(GOLDEN))


def _cases():
    import make_losses_golden as G
    return G, torch.load(GOLDEN)


def test_losses_oracle_matches_reference_outputs():
    from oracle import losses_oracle as LO
    G, gold = _cases()
    for rec in gold["fft"]:
        p, t = G.pair(rec["seed"], *rec["shape"])
        assert float(LO.fft_loss(p, t)) == rec["mean"]
        assert float(LO.fft_loss(p, t, 0.1, "sum")) == rec["sum_w"]
    for rec in gold["mar"]:
        b, c, h, w = rec["shape"]
        y, _ = G.pair(rec["seed"], b, c, h, w)
        x, _ = G.pair(rec["seed"] + 100, b, c, h // 8, w // 8)
        m, a = LO.mar_loss_terms(x, y)
        assert float(m + 0.01 * a) == rec["loss"]


@pytest.mark.gpu
def test_fft_loss_forward(cuda_dev):
    from fdn_tip2025_b200 import losses
    from oracle import losses_oracle as LO
    G, gold = _cases()
    for rec in gold["fft"] + [{"seed": 9, "shape": (1, 3, 640, 1120)}]:
        p, t = G.pair(rec["seed"], *rec["shape"])
        ref = float(LO.fft_loss(p.double(), t.double()))
        got = float(losses.FFTLoss()(p.to(cuda_dev), t.to(cuda_dev)))
        assert abs(got - ref) <= 2e-6 * ref, (rec, got, ref)
        got_s = float(losses.FFTLoss(0.1, "sum")(p.to(cuda_dev), t.to(cuda_dev)))
        assert abs(got_s - 0.1 * ref * (p.numel() // p.shape[-1] * (p.shape[-1] // 2 + 1) * 2)) <= 2e-6 * abs(got_s)
        if "mean" in rec:       # the reference's own fp32 value
            assert abs(got - rec["mean"]) <= 5e-6 * rec["mean"]
    p, t = G.pair(1, 2, 3, 32, 48)
    none = losses.FFTLoss(reduction="none")(p.to(cuda_dev), t.to(cuda_dev)).cpu().double()
    ref = LO.fft_loss(p.double(), t.double(), reduction="none")
    assert none.shape == ref.shape and (none - ref).abs().max() <= 2e-6 * ref.abs().max()


@pytest.mark.gpu
def test_mar_loss_forward(cuda_dev):
    from fdn_tip2025_b200 import losses
    from oracle import losses_oracle as LO
    G, gold = _cases()
    for rec in gold["mar"]:
        b, c, h, w = rec["shape"]
        y, _ = G.pair(rec["seed"], b, c, h, w)
        x, _ = G.pair(rec["seed"] + 100, b, c, h // 8, w // 8)
        m, a = LO.mar_loss_terms(x.double(), y.double())
        gm, ga, yd = losses.MARLoss().terms(x.to(cuda_dev), y.to(cuda_dev))
        assert abs(float(gm) - float(m)) <= 2e-6 * float(m) and abs(float(ga) - float(a)) <= 5e-6 * float(a)
        got = float(losses.MARLoss()(x.to(cuda_dev), y.to(cuda_dev)))
        assert abs(got - rec["loss"]) <= 5e-6 * rec["loss"]
        with_vgg = float(losses.MARLoss()(x.to(cuda_dev), y.to(cuda_dev), lambda u, v: (torch.tensor(0.5, device=cuda_dev), None)))
        assert abs(with_vgg - (got + 5.0)) <= 1e-9 * with_vgg
