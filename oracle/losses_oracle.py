"""CPU restatement (torch, any dtype) of the reference's spectral losses - TEST INFRASTRUCTURE ONLY.

Follows basicsr/models/losses/losses.py: FFTLoss.forward :98-115, MARLoss.forward :769-774 (without its VGG term, which needs
weights that are not part of the reference repository).  Pinned to the reference's own outputs by tests/golden/losses_golden.pt
(generator: tests/golden/make_losses_golden.py, which imports losses.py itself in the build container).
"""
import torch
import torch.nn.functional as F


def fft_loss(pred, target, loss_weight=1.0, reduction="mean"):
    """losses.py:98-115 (the reference calls .float() before the transform; the oracle keeps the caller's dtype)."""
    pf = torch.fft.rfft2(pred, norm="backward")
    pf = torch.stack([pf.real, pf.imag], dim=-1)
    tf = torch.fft.rfft2(target, norm="backward")
    tf = torch.stack([tf.real, tf.imag], dim=-1)
    return loss_weight * F.l1_loss(pf, tf, reduction=reduction)


def mar_loss_terms(x, y):
    """losses.py:769-774: (mse(x, y_d), mse(|rfft2 x|, |rfft2 y_d|)); MARLoss = first + 10 * vgg + 0.01 * second."""
    y_d = F.interpolate(y, scale_factor=1 / 8, mode="bilinear", align_corners=False)
    x_a = torch.abs(torch.fft.rfft2(x, norm="backward"))
    y_a = torch.abs(torch.fft.rfft2(y_d, norm="backward"))
    return F.mse_loss(x, y_d), F.mse_loss(x_a, y_a)
