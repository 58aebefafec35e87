"""CPU restatement (numpy float64) of the reference's validation metrics - TEST INFRASTRUCTURE ONLY.

Follows basicsr/metrics/psnr_ssim.py of the reference line by line; nothing in the product package imports this module.
Pinned to the reference's own outputs by tests/golden/metrics_golden.pt (generator: tests/golden/make_metrics_golden.py, which
imports psnr_ssim.py itself in the build container).  Images are HWC arrays like the reference's.
"""
import numpy as np


def gaussian_kernel(ksize=11, sigma=1.5):
    """cv2.getGaussianKernel(11, 1.5) (psnr_ssim.py:100,144,224): exp(-(i - (n-1)/2)^2 / (2 sigma^2)), normalised, float64."""
    i = np.arange(ksize, dtype=np.float64) - (ksize - 1) / 2.0
    g = np.exp(-(i * i) / (2.0 * sigma * sigma))
    return g / g.sum()


def to_y_channel(img):
    """metric_util.py:34-47 + matlab_functions.bgr2ycbcr(y_only=True) :207-238: first stored channel is weighted as B."""
    img = img.astype(np.float32) / 255.
    if img.ndim == 3 and img.shape[2] == 3:
        y = (np.dot(img, [24.966, 128.553, 65.481]) + 16.0) / 255.
        img = y.astype(np.float32)[..., None]
    return img * 255.


def _crop(img, crop_border):
    return img[crop_border:-crop_border, crop_border:-crop_border, ...] if crop_border else img


def psnr(img1, img2, crop_border=0, test_y_channel=False):
    """calculate_psnr, psnr_ssim.py:8-70 (HWC inputs)."""
    img1 = _crop(np.asarray(img1, dtype=np.float64), crop_border)
    img2 = _crop(np.asarray(img2, dtype=np.float64), crop_border)
    if test_y_channel:
        img1, img2 = to_y_channel(img1), to_y_channel(img2)
    mse = np.mean((img1 - img2) ** 2)
    if mse == 0:
        return float("inf")
    max_value = 1. if img1.max() <= 1 else 255.
    return 20. * np.log10(max_value / np.sqrt(mse))


def _filter_sep(img, g, axis, mode):
    """Correlation with the symmetric 1-D kernel g along `axis`; mode 'edge' = replicate border, 'valid' = no padding."""
    r = len(g) // 2
    if mode == "edge":
        pad = [(0, 0)] * img.ndim
        pad[axis] = (r, r)
        img = np.pad(img, pad, mode="edge")
    n = img.shape[axis] - 2 * r
    out = np.zeros(tuple(n if a == axis else s for a, s in enumerate(img.shape)), dtype=np.float64)
    for d in range(len(g)):
        sl = [slice(None)] * img.ndim
        sl[axis] = slice(d, d + n)
        out += g[d] * img[tuple(sl)]
    return out


def _ssim_map(f, img1, img2, C1, C2):
    mu1, mu2 = f(img1), f(img2)
    s1, s2, s12 = f(img1 ** 2) - mu1 ** 2, f(img2 ** 2) - mu2 ** 2, f(img1 * img2) - mu1 * mu2
    return ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 ** 2 + mu2 ** 2 + C1) * (s1 + s2 + C2))


def ssim(img1, img2, crop_border=0, test_y_channel=False, ssim3d=True):
    """calculate_ssim, psnr_ssim.py:243-329.  The 2-D windows are outer products, the 3-D kernel window * k (psnr_ssim.py:143-151),
    so every filter is applied separably here (same sums, different association)."""
    img1 = _crop(np.asarray(img1, dtype=np.float64), crop_border)
    img2 = _crop(np.asarray(img2, dtype=np.float64), crop_border)
    g = gaussian_kernel()
    if test_y_channel:                                   # _ssim_cly :202-240: BORDER_REPLICATE, constants for range 255
        a, b = to_y_channel(img1)[..., 0].astype(np.float64), to_y_channel(img2)[..., 0].astype(np.float64)
        f = lambda x: _filter_sep(_filter_sep(x, g, 0, "edge"), g, 1, "edge")
        return _ssim_map(f, a, b, (0.01 * 255) ** 2, (0.03 * 255) ** 2).mean()
    max_value = 1 if img1.max() <= 1 else 255
    C1, C2 = (0.01 * max_value) ** 2, (0.03 * max_value) ** 2
    if ssim3d:                                           # _ssim_3d :163-200: Conv3d over (H, W, C), padding_mode='replicate'
        f = lambda x: _filter_sep(_filter_sep(_filter_sep(x, g, 0, "edge"), g, 1, "edge"), g, 2, "edge")
    else:                                                # _ssim :84-117: filter2D per channel, [5:-5, 5:-5]
        f = lambda x: _filter_sep(_filter_sep(x, g, 0, "valid"), g, 1, "valid")
    return float(_ssim_map(f, img1, img2, C1, C2).mean())
