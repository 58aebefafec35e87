"""CPU oracle for the FDN / FDformer / MAR / LPNet inference forward.  TEST INFRASTRUCTURE ONLY.

This is a functional restatement (state_dict in, tensors out) of the reference's algorithm with torch
CPU ops as the arithmetic library.  It is what ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` check against and time; nothing in the
product package may import it.  It evaluates in the dtype of its inputs, so feeding float64 tensors
gives the fp64 oracle used for the per-block gates.

Parity pin: the reference ships no tests or golden vectors, so the pin is the reference
itself: ``oracle/validate_against_reference.py`` imports the four reference arch files in the build
container and checks every function below against them; ``tests/golden/make_golden.py`` stores the
reference's own outputs as fixtures that ``tests/test_oracle_golden.py`` replays on any machine.

Reference lines followed (all under basicsr/models/archs/):
  layer_norm      FDN_arch.py:326-342           rd (replace_denormals)  FDN_arch.py:548-553
  fdsa            FDN_arch.py:575-641           fdffn                   FDN_arch.py:453-475
  fcaffn          FDN_arch.py:405-429           transformer_block       FDN_arch.py:666-677
  fuse            FDN_arch.py:688-695           fdformer                FDN_arch.py:810-842
  fre_block       FDN_arch.py:88-100            process_block           FDN_arch.py:109-118, fdnlol24_arch.py:769-776
  fourier_fuse    FDN_arch.py:136-148           mar_core                FDN_arch.py:203-257, fdnlol24_arch.py:151-207
  mar             FDN_arch.py:269-286           fdn                     FDN_arch.py:869-921, fdnlol24_arch.py:981-1033
  lpnet           LPNet_arch.py:70-81, 114-134
"""
import torch
import torch.nn.functional as F

P = 8  # spectral patch size


# ----------------------------------------------------------------------------- helpers
def layer_norm(x, sd, p, eps=1e-5):
    """Per-pixel LayerNorm over the channel axis of NCHW, biased variance."""
    w, b = sd[p + "body.weight"], sd[p + "body.bias"]
    mu = x.mean(1, keepdim=True)
    var = ((x - mu) ** 2).mean(1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def rd(z, thr=1e-10):
    """Real and imaginary parts with |v| < thr become +thr (sign discarded)."""
    re = torch.where(z.real.abs() < thr, torch.full_like(z.real, thr), z.real)
    im = torch.where(z.imag.abs() < thr, torch.full_like(z.imag, thr), z.imag)
    return torch.complex(re, im)


def polar(mag, pha):
    return torch.complex(mag * torch.cos(pha), mag * torch.sin(pha))


def conv(x, sd, p, stride=1, padding=0, groups=1):
    return F.conv2d(x, sd[p + "weight"], sd.get(p + "bias"), stride=stride, padding=padding, groups=groups)


def to_patches(x):
    b, c, h, w = x.shape
    return x.view(b, c, h // P, P, w // P, P).permute(0, 1, 2, 4, 3, 5)


def from_patches(x):
    b, c, hp, wp, _, _ = x.shape
    return x.permute(0, 1, 2, 4, 3, 5).reshape(b, c, hp * P, wp * P)


def half(x):
    """nn.Upsample(scale_factor=0.5, bilinear, align_corners=False) == 2x2 mean."""
    return F.avg_pool2d(x, 2)


def up2_bilinear(x):
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)


def gelu(x):
    return F.gelu(x)


def gated_dw(x, sd, p):
    """dw3x3 C->2C (groups=C), chunk, gelu(x1)*x2."""
    c = x.shape[1]
    x1, x2 = conv(x, sd, p, padding=1, groups=c).chunk(2, dim=1)
    return gelu(x1) * x2


# ----------------------------------------------------------------------------- FDformer blocks
def fdsa(x, sd, p):
    e = sd[p + "fft"].shape[0]
    hid = conv(x, sd, p + "to_hidden.")
    hid = conv(hid, sd, p + "to_hidden_dw.", padding=1, groups=4 * e)
    q, k, v, vv = hid.chunk(4, dim=1)
    q, k, v = (torch.fft.rfft2(to_patches(t)) for t in (q, k, v))
    v = rd(v * sd[p + "fft"])
    qk_mag = rd(q * k).abs()
    theta = torch.angle(rd(q)) - torch.angle(rd(k))
    o1 = torch.fft.irfft2(polar(v.abs(), theta), s=(P, P))
    o2 = torch.fft.irfft2(polar(qk_mag, torch.angle(v)), s=(P, P))
    o3 = torch.fft.irfft2(polar(qk_mag, theta), s=(P, P))
    outs = [layer_norm(from_patches(o), sd, p + "norm%d." % (i + 1)) * vv for i, o in enumerate((o1, o2, o3))]
    return conv(torch.cat(outs, 1), sd, p + "project_out.")


def fdffn(x, sd, p):
    hd = sd[p + "ffta"].shape[0]
    x = conv(x, sd, p + "project_in.")
    sp = conv(x, sd, p + "space.0.", padding=1, groups=hd)
    sp = conv(gelu(sp), sd, p + "space.2.", padding=1, groups=hd)
    z = rd(torch.fft.rfft2(to_patches(x)))
    z = polar(z.abs() * sd[p + "ffta"], torch.angle(z) - sd[p + "fftp"])
    x = from_patches(torch.fft.irfft2(z, s=(P, P))) + sp
    return conv(gated_dw(x, sd, p + "dwconv."), sd, p + "project_out.")


def fcaffn(x, amp, pha, img, sd, p):
    c = x.shape[1]
    h, w = x.shape[-2:]
    z = rd(torch.fft.rfft2(x))
    z_p = torch.angle(z) - conv(pha, sd, p + "conv1_xp.")
    z_a = conv(amp, sd, p + "conv1_xa.") * z.abs()
    y = torch.fft.irfft2(polar(z_a, z_p), s=(h, w))
    y = layer_norm(y, sd, p + "norm.") * x + x
    y = conv(y, sd, p + "project_in.")
    mul = conv(conv(img, sd, p + "conv1_mul."), sd, p + "conv3_mul.", padding=1, groups=c)
    add = conv(conv(img, sd, p + "conv1_add."), sd, p + "conv3_add.", padding=1, groups=c)
    y = y * mul + add
    return conv(gated_dw(y, sd, p + "dwconv."), sd, p + "project_out.")


def transformer_block(x, side, sd, p):
    """side = (amp, pha, img) maps of this level; att/light inferred from the keys present."""
    if (p + "attn.fft") in sd:
        x = x + fdsa(layer_norm(x, sd, p + "norm1."), sd, p + "attn.")
    x = x + fdffn(layer_norm(x, sd, p + "norm2."), sd, p + "ffn.")
    if (p + "ffn2.project_in.weight") in sd:
        x = x + fcaffn(layer_norm(x, sd, p + "norm3."), side[0], side[1], side[2], sd, p + "ffn2.")
    return x


def stage(x, side, sd, p):
    i = 0
    while (p + "%d.norm2.body.weight" % i) in sd:
        x = transformer_block(x, side, sd, p + "%d." % i)
        i += 1
    return x


def fuse(enc, dec, sd, p):
    n = enc.shape[1]
    x = conv(torch.cat((enc, dec), 1), sd, p + "conv.")
    x = transformer_block(x, None, sd, p + "att_channel.")
    x = conv(x, sd, p + "conv2.")
    return x[:, :n] + x[:, n:]


def fdformer(img, side1, side2, side3, sd, p="net_p.", ori=None):
    """side_l = (amplitude map, phase map, MAR image) at level l."""
    x1 = conv(img, sd, p + "patch_embed.proj.", padding=1)
    x1 = stage(x1, side1, sd, p + "encoder_level1.")
    x2 = conv(half(x1), sd, p + "down1_2.body.1.", padding=1)
    x2 = stage(x2, side2, sd, p + "encoder_level2.")
    x3 = conv(half(x2), sd, p + "down2_3.body.1.", padding=1)
    x3 = stage(x3, side3, sd, p + "encoder_level3.")
    x3 = stage(x3, side3, sd, p + "decoder_level3.")
    y2 = conv(up2_bilinear(x3), sd, p + "up3_2.body.1.", padding=1)
    y2 = fuse(y2, x2, sd, p + "fuse2.")
    y2 = stage(y2, side2, sd, p + "decoder_level2.")
    y1 = conv(up2_bilinear(y2), sd, p + "up2_1.body.1.", padding=1)
    y1 = fuse(y1, x1, sd, p + "fuse1.")
    y1 = stage(y1, side1, sd, p + "decoder_level1.")
    y1 = stage(y1, side1, sd, p + "refinement.")
    return conv(y1, sd, p + "output.", padding=1) + (img if ori is None else ori)


# ----------------------------------------------------------------------------- MAR
def lrelu(x):
    return F.leaky_relu(x, 0.1)


def _mlp(x, sd, p):
    return conv(lrelu(conv(x, sd, p + "0.")), sd, p + "2.")


def _spectral_mlp(x, sd, p, h, w):
    """rfft2 -> (|.|, angle) -> per-bin channel MLPs -> polar -> irfft2(s=(h, w)) (slices the spectrum)."""
    z = torch.fft.rfft2(x)
    z = polar(_mlp(z.abs(), sd, p + "process1."), _mlp(torch.angle(z), sd, p + "process2."))
    return torch.fft.irfft2(z, s=(h, w))


def fre_block(x, sd, p):
    h, w = x.shape[-2:]
    return _spectral_mlp(conv(x, sd, p + "fpre."), sd, p, h, w) + x


def process_block(x, sd, p, variant):
    y = fre_block(x, sd, p + "frequency_process.")
    if variant == "lolv1":
        y = conv(y, sd, p + "cat.")
    return y + x


def fourier_fuse(x1, x2, x4, sd, p):
    x = torch.cat((x1, x2, x4), 1)
    h, w = x.shape[-2:]
    y = conv(x, sd, p + "fpre.0.")
    y = conv(y, sd, p + "fpre.1.", padding=1, groups=y.shape[1])      # 1x1 depthwise with padding 1 -> (h+2, w+2)
    return conv(_spectral_mlp(y, sd, p, h, w), sd, p + "fourier_out.", padding=1)


def nearest_down(x):
    return x[..., ::2, ::2]


def nearest_up(x):
    return x.repeat_interleave(2, -2).repeat_interleave(2, -1)


def mar_core(x, ratio, sd, p, variant="lolblur", use_ratio=True):
    x_2 = nearest_down(x)
    x_4 = nearest_down(x_2)
    z2 = process_block(conv(F.pixel_unshuffle(x, 2), sd, p + "f2.0."), sd, p + "f2.1.", variant)
    z4 = process_block(conv(F.pixel_unshuffle(x, 4), sd, p + "f1.0."), sd, p + "f1.1.", variant)
    x_ = process_block(conv(x, sd, p + "f3.0."), sd, p + "f3.1.", variant)
    if use_ratio:
        z2, z4, x_ = z2 * ratio, z4 * ratio, x_ * ratio
    res1 = process_block(x_, sd, p + "Encoder.0.", variant)
    z = lrelu(conv(res1, sd, p + "f3_down.main.0.", stride=2, padding=1))
    z = conv(conv(torch.cat((z, z2), 1), sd, p + "FAM2.merge1."), sd, p + "FAM2.merge2.", padding=1)
    res2 = process_block(z, sd, p + "Encoder.1.", variant)
    z = lrelu(conv(res2, sd, p + "f2_down.main.0.", stride=2, padding=1))
    z = conv(conv(torch.cat((z, z4), 1), sd, p + "FAM1.merge1."), sd, p + "FAM1.merge2.", padding=1)
    z = process_block(z, sd, p + "Encoder.2.", variant)

    z12, z21 = nearest_down(res1), nearest_up(res2)
    z42 = nearest_up(z)
    z41 = nearest_up(z42)
    res2 = fourier_fuse(z12, res2, z42, sd, p + "AFFs.1.")
    res1 = fourier_fuse(res1, z21, z41, sd, p + "AFFs.0.")

    outs = []
    z = process_block(z, sd, p + "Decoder.0.", variant)
    outs.append(torch.sigmoid(conv(z, sd, p + "ConvsOut.0.main.0.", padding=1) + x_4) + 1e-8)
    z = lrelu(F.conv_transpose2d(z, sd[p + "f2_up.main.0.weight"], sd[p + "f2_up.main.0.bias"], stride=2, padding=1))
    z = lrelu(conv(torch.cat((z, res2), 1), sd, p + "Convs.0.main.0."))
    z = process_block(z, sd, p + "Decoder.1.", variant)
    outs.append(torch.sigmoid(conv(z, sd, p + "ConvsOut.1.main.0.", padding=1) + x_2) + 1e-8)
    z = lrelu(F.conv_transpose2d(z, sd[p + "f3_up.main.0.weight"], sd[p + "f3_up.main.0.bias"], stride=2, padding=1))
    z = lrelu(conv(torch.cat((z, res1), 1), sd, p + "Convs.1.main.0."))
    z = process_block(z, sd, p + "Decoder.2.", variant)
    outs.append(torch.sigmoid(conv(z, sd, p + "out.main.0.", padding=1) + x) + 1e-8)
    return outs  # (1/4, 1/2, 1) illumination maps


def mar(x, ratio, sd, p="", variant="lolblur", use_ratio=True):
    """ratio: [B,1,1,1].  Returns the gamma-corrected pyramid (1/4, 1/2, 1)."""
    i3, i2, i1 = mar_core(x, ratio, sd, p + "net.", variant, use_ratio)
    x1 = x
    x2 = half(x1)
    x3 = half(x2)
    g = lambda img, i: 1.0 - torch.pow(1.0 - img, i * 40.0)
    return g(x3, i3), g(x2, i2), g(x1, i1)


# ----------------------------------------------------------------------------- FDN
def fdn(img, ratio_i, sd, variant="lolblur"):
    """img [B,3,H,W] in [0,1]; ratio_i [B,1].  Returns the reference's 4-tuple."""
    ratio = ratio_i.view(-1, 1, 1, 1)
    pyr = [img, half(img), half(half(img))]
    norms = ("norm1.", "norm2.", "norm3.")
    pha = [torch.angle(rd(torch.fft.rfft2(layer_norm(t, sd, n)))) for t, n in zip(pyr, norms)]
    q3, q2, q1 = mar(img, ratio, sd, "net_a.", variant)
    amp = [torch.fft.rfft2(layer_norm(t, sd, n)).abs() for t, n in zip((q1, q2, q3), norms)]
    out = fdformer(img, (amp[0], pha[0], q1), (amp[1], pha[1], q2), (amp[2], pha[2], q3), sd, "net_p.")
    if variant == "lolv1":
        return out, out, out, out
    return out, q1, q2, q3


# ----------------------------------------------------------------------------- LPNet
def _bn(x, sd, p, eps=1e-5):
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                        training=False, eps=eps)


def se_block(x, sd, p, stride):
    y = F.relu(_bn(conv(x, sd, p + "conv1.0.", stride=stride), sd, p + "conv1.1."))
    y = F.relu(_bn(conv(y, sd, p + "conv2.0.", padding=1), sd, p + "conv2.1."))
    y = _bn(conv(y, sd, p + "conv3.0."), sd, p + "conv3.1.")
    s = y.mean((2, 3), keepdim=True)
    s = torch.sigmoid(conv(F.relu(conv(s, sd, p + "se.1.")), sd, p + "se.3."))
    y = y * s
    if (p + "shortcut.0.weight") in sd:
        x = _bn(conv(x, sd, p + "shortcut.0.", stride=stride), sd, p + "shortcut.1.")
    return F.relu(y + x)


def gray_mean(x):
    """torchvision Grayscale (0.2989 R + 0.587 G + 0.114 B) then spatial mean -> [B,1]."""
    g = 0.2989 * x[:, 0] + 0.587 * x[:, 1] + 0.114 * x[:, 2]
    return g.mean((1, 2)).view(-1, 1)


def lpnet(x, sd, use_ori_i=False):
    g = gray_mean(x)
    y = F.relu(_bn(conv(x, sd, "conv1.0.", stride=2, padding=3), sd, "conv1.1."))
    y = F.avg_pool2d(y, 3, 2, 1)
    for name, num, stride in (("conv2", 3, 1), ("conv3", 3, 2), ("conv4", 6, 6)):
        for i in range(num):
            y = se_block(y, sd, "%s.%d." % (name, i), stride if i == 0 else 1)
    y = y.mean((2, 3))
    y = F.linear(y, sd["fc.0.weight"], sd["fc.0.bias"])
    y = torch.sigmoid(F.linear(y, sd["fc2.0.weight"], sd["fc2.0.bias"]))
    return g / y if use_ori_i else y


# ----------------------------------------------------------------------------- utilities for tests
def to_dtype(sd, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


def psnr(a, b, peak=1.0):
    mse = ((a.double() - b.double()) ** 2).mean().item()
    return float("inf") if mse == 0 else 10.0 * torch.log10(torch.tensor(peak * peak / mse)).item()
