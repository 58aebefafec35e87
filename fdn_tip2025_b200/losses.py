"""Forward passes of the reference's spectral losses on the B200 global-FFT kernels.

Same class names and constructor arguments as basicsr/models/losses/losses.py (FFTLoss :83-115, MARLoss :764-774).
Forward only - these modules run under no_grad and return float64 scalars on the device; a training step that needs
gradients keeps using autograd on the reference path (the backward of the FFT kernels is the row after this one).
"""
import torch
import torch.nn as nn

from . import ops

__all__ = ["FFTLoss", "MARLoss"]


def _planes(t):
    ops._device_ok(t)
    if t.dim() < 2:
        raise RuntimeError("expected a (..., H, W) tensor")
    return t.detach().float().contiguous()


def _rfft2(x, mode):
    """x [..., H, W] -> forward 2-D real FFT: mode COLS_FWD -> interleaved spectrum [..., H, W/2+1, 2]; COLS_FWD_ABS -> |X| map."""
    h, w = x.shape[-2:]
    wf = w // 2 + 1
    planes = x.numel() // (h * w)
    spec = torch.empty(*x.shape[:-1], wf, 2, dtype=torch.float32, device=x.device)
    ops.fft_rows_r2c(x, spec)
    if mode == ops.COLS_FWD:
        ops.fft_cols(spec, h * wf, wf, spec, h * wf, wf, planes, h, wf, w, ops.COLS_FWD)
        return spec
    out = torch.empty(*x.shape[:-1], wf, dtype=torch.float32, device=x.device)
    ops.fft_cols(spec, h * wf, wf, out, h * wf, wf, planes, h, wf, w, ops.COLS_FWD_ABS)
    return out


class FFTLoss(nn.Module):
    """losses.py:83-115: loss_weight * l1_loss(stack(re, im)(rfft2(pred)), stack(re, im)(rfft2(target))).

    The transform is linear, so rfft2(pred) - rfft2(target) is evaluated as rfft2(pred - target): one transform instead of two
    (same value up to fp32 rounding of the two spectra).  reduction 'mean' | 'sum' (no element-wise weight, like every call site
    of the reference: image_restoration_model.py passes none)."""

    def __init__(self, loss_weight=1.0, reduction="mean"):
        super().__init__()
        if reduction not in ("none", "mean", "sum"):
            raise ValueError("Unsupported reduction mode: %s. Supported ones are: ['none', 'mean', 'sum']" % reduction)
        self.loss_weight = loss_weight
        self.reduction = reduction

    @torch.no_grad()
    def forward(self, pred, target, weight=None, **kwargs):
        if weight is not None:
            raise NotImplementedError("element-wise weights are not used by the reference's FFT loss call sites")
        p, t = _planes(pred), _planes(target)
        if p.shape != t.shape:
            raise RuntimeError("pred and target must have the same shape")
        with torch.cuda.device(p.device):
            d = torch.empty_like(p)
            ops.diff(p, t, d)
            spec = _rfft2(d, ops.COLS_FWD)
            if self.reduction == "none":
                return self.loss_weight * spec.abs()
            total = ops.reduce_f64(spec, None, 0)
        if self.reduction == "mean":
            total = total / spec.numel()
        return (self.loss_weight * total).reshape(())


class MARLoss(nn.Module):
    """losses.py:764-774.  forward(x, y, vgg_loss): mse(x, y_d) + 10 * vgg_loss(x, y_d)[0] + 0.01 * mse(|rfft2(x)|, |rfft2(y_d)|)
    with y_d the 1/8 bilinear resample of y.  ``vgg_loss`` is the caller's perceptual-loss module (it needs VGG weights and is
    outside this library); pass None to get the two terms computed here."""

    def __init__(self, scale=1 / 8):
        super().__init__()
        if scale != 1 / 8:
            raise NotImplementedError("the reference instantiates MARLoss with its default scale 1/8 only")
        self.scale = scale

    @torch.no_grad()
    def terms(self, x, y):
        """(mse(x, y_d), mse(|rfft2 x|, |rfft2 y_d|)) as float64 scalars on the device."""
        x, y = _planes(x), _planes(y)
        with torch.cuda.device(x.device):
            yd = torch.empty(*y.shape[:-2], y.shape[-2] // 8, y.shape[-1] // 8, dtype=torch.float32, device=y.device)
            ops.down8_bilinear(y, yd)
            if yd.shape != x.shape:
                raise RuntimeError("x must have 1/8 of y's spatial size")
            mse = ops.reduce_f64(x, yd, 1) / x.numel()
            xa, ya = _rfft2(x, ops.COLS_FWD_ABS), _rfft2(yd, ops.COLS_FWD_ABS)
            mse_a = ops.reduce_f64(xa, ya, 1) / xa.numel()
        return mse.reshape(()), mse_a.reshape(()), yd

    @torch.no_grad()
    def forward(self, x, y, vgg_loss=None):
        mse, mse_a, yd = self.terms(x, y)
        loss = mse + 0.01 * mse_a
        if vgg_loss is not None:
            loss = loss + 10.0 * vgg_loss(x, yd)[0]
        return loss
