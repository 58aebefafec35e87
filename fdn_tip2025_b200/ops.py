"""Tensor-level wrappers over the C ABI (one Python function per entry point of include/fdn_b200.h).

PyTorch is used for device memory and streams only; every function here launches kernels from
libfdn_b200.so on ``torch.cuda.current_stream()`` and returns immediately.
"""
import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _device_ok(t):
    if not t.is_cuda:
        raise RuntimeError("fdn_tip2025_b200 runs on CUDA tensors only (got %s); there is no CPU path" % t.device)


# bench.py's per-launch accounting: bytes of every tensor operand of the next C-ABI call (each counted once)
count_bytes = False
pending_bytes = 0


def _p(t):
    global pending_bytes
    if t is None:
        return None
    _device_ok(t)
    assert t.dtype == torch.float32 and t.is_contiguous(), "tensors crossing the C ABI are contiguous fp32"
    if count_bytes:
        pending_bytes += t.numel() * 4
    return t.data_ptr()


def empty(shape, like):
    return torch.empty(shape, dtype=torch.float32, device=like.device)


# ------------------------------------------------------------------------------------------------ FFT
def fft_prepare(h, w):
    _lib.call("fdn_fft_prepare", h, w)


def fft_rows_r2c(x, spec):
    """x [..., H, W] -> spec [..., H, W/2+1, 2]"""
    h, w = x.shape[-2:]
    planes = x.numel() // (h * w)
    _lib.call("fdn_fft_rows_r2c", _p(x), _p(spec), planes, h, w, _stream())


def fft_rows_c2r(spec, y, inv_norm, res=None, res_coef=0.0, img_scale=None, planes_per_image=1):
    h, w = y.shape[-2:]
    planes = y.numel() // (h * w)
    _lib.call("fdn_fft_rows_c2r", _p(spec), _p(y), planes, h, w, float(inv_norm), _p(res), float(res_coef), _p(img_scale),
              planes_per_image, _stream())


COLS_FWD, COLS_INV, COLS_FWD_MOD_INV, COLS_FWD_ANGLE, COLS_FWD_ABS = range(5)


def fft_cols(src, src_ps, src_rs, dst, dst_ps, dst_rs, planes, h, ncols, w_real, mode, c=0, amp=None, pha=None, w_xa=None,
             w_xp=None):
    _lib.call("fdn_fft_cols", _p(src), src_ps, src_rs, _p(dst), dst_ps, dst_rs, planes, h, ncols, w_real, mode, c, _p(amp),
              _p(pha), _p(w_xa), _p(w_xp), _stream())


def spec_mlp(spec, plane_stride, nbins, b, nc, wpack):
    _lib.call("fdn_spec_mlp", _p(spec), plane_stride, nbins, b, nc, _p(wpack), _stream())


# ------------------------------------------------------------------------------------------------ patch spectral
def fdffn_patch(x, add, wspec, out):
    b, c, h, w = x.shape
    _lib.call("fdn_fdffn_patch", _p(x), _p(add), _p(wspec), _p(out), b, c, h, w, _stream())


def fdffn_patch_dw(h, s1, wb, wspec, out):
    b, c, hh, w = h.shape
    _lib.call("fdn_fdffn_patch_dw", _p(h), _p(s1), _p(wb), _p(wspec), _p(out), b, c, hh, w, _stream())


def fdffn_spatial(h, wa, wb, wspec, out):
    b, c, hh, w = h.shape
    _lib.call("fdn_fdffn_spatial", _p(h), _p(wa), _p(wb), _p(wspec), _p(out), b, c, hh, w, _stream())


def fdsa_patch(hid, wfft, out):
    b, c4, h, w = hid.shape
    _lib.call("fdn_fdsa_patch", _p(hid), _p(wfft), _p(out), b, c4 // 4, h, w, _stream())


def fdsa_patch_dw(hid, wdw, wfft, out, vv):
    b, c4, h, w = hid.shape
    _lib.call("fdn_fdsa_patch_dw", _p(hid), _p(wdw), _p(wfft), _p(out), _p(vv), b, c4 // 4, h, w, _stream())


# ------------------------------------------------------------------------------------------------ per pixel
def pw_conv(srcs, wt, out, bias=None, ln=None, act=0, film=None, res=None, res_coef=1.0, img_scale=None, out_view=None):
    """srcs: list of (tensor [B,C,Hs,Ws], shift).  wt [K][N].  out [B,N,H,W] (or out_view=(bs, ps, rs) strides into `out`)."""
    b = srcs[0][0].shape[0]
    n = wt.shape[1]
    if out_view is None:
        h, w = out.shape[-2:]
        bs, ps, rs = n * h * w, h * w, w
    else:
        h, w, bs, ps, rs = out_view
    args = []
    for i in range(3):
        if i < len(srcs):
            t, sh = srcs[i]
            args += [_p(t), t.shape[1], sh]
        else:
            args += [None, 0, 0]
    assert sum(t.shape[1] for t, _ in srcs) == wt.shape[0], "channel count does not match the weight"
    _lib.call("fdn_pw_conv", *args, _p(wt), _p(bias), _p(ln[0]) if ln else None, _p(ln[1]) if ln else None, act,
              _p(film[0]) if film else None, _p(film[1]) if film else None, _p(res), float(res_coef), _p(img_scale), _p(out),
              bs, ps, rs, b, n, h, w, _stream())


def has_tcgen05():
    return _lib.load().fdn_has_tcgen05() == 1


def pw_mma_supported(k, n, prologue=0, stats_in_kernel=False):
    """Whether the tensor-core 1x1 kernel can run a layer with k inputs and n outputs (shared-memory plan, K limit)."""
    from . import packing
    nc, _ = packing.chunking(n)
    return _lib.load().fdn_pw_mma_supported(int(k), int(nc), int(prologue), 1 if stats_in_kernel else 0) == 1


def pw_mma(srcs, packed, out, prologue=0, ln=None, aux=None, aux_bs=0, stats=None, bias=None, film=None, res=None, res_coef=1.0, passes=3):
    """srcs: one or two [B,C,H,W] tensors (channel concat); packed = packing.pack_weight(w) on the same device."""
    bpack, n, nc, nchunks = packed
    b, _, h, w = srcs[0].shape
    s1 = srcs[1] if len(srcs) > 1 else None
    _lib.call("fdn_pw_mma", _p(srcs[0]), srcs[0].shape[1], _p(s1), s1.shape[1] if s1 is not None else 0, _p(bpack), n, nc, nchunks,
              prologue, _p(ln[0]) if ln else None, _p(ln[1]) if ln else None, _p(aux), aux_bs, _p(stats), _p(bias),
              _p(film[0]) if film else None, _p(film[1]) if film else None, _p(res), float(res_coef), _p(out), b, h * w, passes,
              _stream())


def group_stats(x, stats, groups):
    b, gc, h, w = x.shape
    _lib.call("fdn_group_stats", _p(x), _p(stats), b, groups, gc // groups, h * w, _stream())


def chan_ln(x, out, gamma, beta, groups=1, mul=None, mul_bs=0, add=None, add_bs=0):
    b, gc, h, w = x.shape
    _lib.call("fdn_chan_ln", _p(x), _p(out), _p(gamma), _p(beta), _p(mul), mul_bs, _p(add), add_bs, b, groups, gc // groups,
              h * w, _stream())


def avgpool2(x, out):
    h, w = x.shape[-2:]
    _lib.call("fdn_avgpool2", _p(x), _p(out), x.numel() // (h * w), h, w, _stream())


def up2_bilinear(x, out):
    h, w = x.shape[-2:]
    _lib.call("fdn_up2_bilinear", _p(x), _p(out), x.numel() // (h * w), h, w, _stream())


def pixel_unshuffle(x, out, r):
    b, c, h, w = x.shape
    _lib.call("fdn_pixel_unshuffle", _p(x), _p(out), b, c, h, w, r, _stream())


def gamma_curve(x, illum, out, scale=40.0):
    _lib.call("fdn_gamma_curve", _p(x), _p(illum), _p(out), float(scale), x.numel(), _stream())


def fill_border(t, value, c):
    h, w = t.shape[-2:]
    _lib.call("fdn_fill_border", _p(t), _p(value), t.numel() // (h * w), c, h, w, _stream())


# ------------------------------------------------------------------------------------------------ spatial convs
def conv2d(x, w, out, bias=None, res=None, res_shift=0, stride=1, pad=1, act=0, head=0):
    b, cin, h, wd = x.shape
    cout, _, k, _ = w.shape
    _lib.call("fdn_conv2d", _p(x), _p(w), _p(bias), _p(res), res_shift, _p(out), b, cin, h, wd, cout, k, stride, pad, act, head,
              _stream())


def conv3x3_mma(x, wpack, out, bias=None, res=None):
    """Tensor-core 3x3 conv (stride 1, padding 1); wpack = packing.pack_conv3x3(weight)."""
    b, cin, h, wd = x.shape
    _lib.call("fdn_conv3x3_mma", _p(x), _p(wpack), _p(bias), _p(res), _p(out), b, cin, h, wd, out.shape[1], _stream())


def film_maps(img, wmul, wadd, omul, oadd):
    b, c, h, w = omul.shape
    _lib.call("fdn_film_maps", _p(img), _p(wmul), _p(wadd), _p(omul), _p(oadd), b, c, h, w, _stream())


def convt4s2(x, w, bias, out, act=1):
    b, cin, h, wd = x.shape
    _lib.call("fdn_convt4s2", _p(x), _p(w), _p(bias), _p(out), b, cin, w.shape[1], h, wd, act, _stream())


def dwconv3(x, w, out, mode=0):
    b, c, h, wd = x.shape
    _lib.call("fdn_dwconv3", _p(x), _p(w), _p(out), b, c, h, wd, mode, _stream())


# ------------------------------------------------------------------------------------------------ LPNet
def avgpool3s2(x, out):
    h, w = x.shape[-2:]
    _lib.call("fdn_avgpool3s2", _p(x), _p(out), x.numel() // (h * w), h, w, _stream())


def plane_mean(x, out):
    h, w = x.shape[-2:]
    _lib.call("fdn_plane_mean", _p(x), _p(out), x.numel() // (h * w), h * w, _stream())


def se_fc(m, w1, b1, w2, b2, s):
    b, c = m.shape
    _lib.call("fdn_se_fc", _p(m), _p(w1), _p(b1), _p(w2), _p(b2), _p(s), b, c, w1.shape[0], _stream())


def se_apply(x, s, shortcut, out):
    h, w = x.shape[-2:]
    _lib.call("fdn_se_apply", _p(x), _p(s), _p(shortcut), _p(out), x.numel() // (h * w), h * w, _stream())


def lpnet_head(m, w1, b1, w2, b2, gray, out):
    b, c = m.shape
    _lib.call("fdn_lpnet_head", _p(m), _p(w1), _p(b1), _p(w2), _p(b2), _p(gray), _p(out), b, c, _stream())


def gray_mean(x, out):
    b, _, h, w = x.shape
    _lib.call("fdn_gray_mean", _p(x), _p(out), b, h * w, _stream())


# ------------------------------------------------------------------------------------------------ image pre/post (inference scripts)
def _pu8(t):
    _device_ok(t)
    assert t.dtype == torch.uint8 and t.is_contiguous(), "image buffers are contiguous uint8"
    return t.data_ptr()


def pre_u8hwc(img, out):
    """img [B,h,w,3] uint8 BGR -> out [B,3,Hp,Wp] fp32 RGB in [0,1], reflect-padded (inference_fdn_lolblur.py:47-62)."""
    b, h, w, _ = img.shape
    _lib.call("fdn_pre_u8hwc_to_f32chw", _pu8(img), _p(out), b, h, w, out.shape[2], out.shape[3], _stream())


def post_u8hwc(x, img):
    """x [B,3,Hp,Wp] fp32 RGB -> img [B,h,w,3] uint8 BGR: crop, clamp, *255, round (tensor2img, img_util.py:36-98)."""
    b, h, w, _ = img.shape
    _lib.call("fdn_post_f32chw_to_u8hwc", _p(x), _pu8(img), b, h, w, x.shape[2], x.shape[3], _stream())


# ------------------------------------------------------------------------------------------------ validation metrics (basicsr/metrics/psnr_ssim.py)
def _pd(t):
    _device_ok(t)
    assert t.dtype == torch.float64 and t.is_contiguous(), "metric results / scratch are contiguous float64"
    return t.data_ptr()


def psnr(img1, img2, crop_border=0, test_y_channel=False):
    """calculate_psnr (psnr_ssim.py:8-70) per image: img1, img2 [B,C,H,W] fp32 on the device -> float64 [B] on the device."""
    b, c, h, w = img1.shape
    assert img2.shape == img1.shape, "Image shapes are different"
    out = torch.empty(b, dtype=torch.float64, device=img1.device)
    ws = torch.empty(4 * b, dtype=torch.float64, device=img1.device)
    _lib.call("fdn_psnr", _p(img1), _p(img2), _pd(out), _pd(ws), b, c, h, w, crop_border, 1 if test_y_channel else 0, _stream())
    return out


def ssim(img1, img2, crop_border=0, test_y_channel=False, ssim3d=True):
    """calculate_ssim (psnr_ssim.py:243-329) per image, float64 [B] on the device."""
    b, c, h, w = img1.shape
    assert img2.shape == img1.shape, "Image shapes are different"
    mode = 2 if test_y_channel else (0 if ssim3d else 1)
    out = torch.empty(b, dtype=torch.float64, device=img1.device)
    ws = torch.empty(4 * b, dtype=torch.float64, device=img1.device)
    _lib.call("fdn_ssim", _p(img1), _p(img2), _pd(out), _pd(ws), b, c, h, w, crop_border, mode, _stream())
    return out


# ------------------------------------------------------------------------------------------------ spectral losses (forward)
def diff(a, b, out):
    _lib.call("fdn_diff", _p(a), _p(b), _p(out), a.numel(), _stream())


def reduce_f64(a, b, mode):
    """mode 0: sum |a|; mode 1: sum (a - b)^2.  Returns a float64 scalar tensor on the device."""
    out = torch.empty(1, dtype=torch.float64, device=a.device)
    _lib.call("fdn_reduce_f64", _p(a), _p(b), _pd(out), a.numel(), mode, _stream())
    return out


def down8_bilinear(x, out):
    h, w = x.shape[-2:]
    _lib.call("fdn_down8_bilinear", _p(x), _p(out), x.numel() // (h * w), h, w, _stream())
