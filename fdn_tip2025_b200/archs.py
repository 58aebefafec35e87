"""Drop-in nn.Modules for the reference's inference forward: FDN, FDN_lolv1, FDformer, MAR, I_predict_net.

Same class names, constructor arguments, forward signatures, return tuples and state_dict keys as
basicsr/models/archs/{FDN_arch,fdnlol24_arch,mar_arch,LPNet_arch}.py, so
``net.load_state_dict(torch.load(ckpt)["params"], strict=True)`` and the inference scripts work unchanged.
The modules only own parameters; ``forward`` walks the network and launches the sm_100a kernels of
libfdn_b200.so through ``ops`` (C ABI).  Inference only: no autograd, CUDA tensors only, no CPU fallback.

Unlike the reference constructors (FDN_arch.py:860-862, fdnlol24_arch.py:972-974) nothing is torch.load-ed from a
hard-coded path: ``net_a.*`` is filled by the full checkpoint like every other key.
"""
import contextlib
import os

import torch
import torch.nn as nn

from . import ops, packing, schema

__all__ = ["FDN", "FDN_lolv1", "FDformer", "MAR", "I_predict_net"]


# =====================================================================================================
# parameter tree + packed-weight cache
# =====================================================================================================
class _Node(nn.Module):
    """Anonymous container so dotted state_dict keys map onto nested modules."""


def _init_tensor(shape, kind):
    if kind == "ones" or kind == "buf_ones":
        return torch.ones(shape)
    if kind == "zeros" or kind == "buf_zeros":
        return torch.zeros(shape)
    if kind == "buf_long":
        return torch.zeros((), dtype=torch.long)
    if kind in ("conv", "linear"):
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
        bound = 1.0 / fan_in ** 0.5
        return (torch.rand(shape) * 2 - 1) * bound
    if kind.startswith("conv_bias:"):
        bound = 1.0 / int(kind.split(":")[1]) ** 0.5
        return (torch.rand(shape) * 2 - 1) * bound
    raise ValueError(kind)


class _Net(nn.Module):
    """Owns the parameters listed by a schema table and a cache of kernel-ready (packed) weights."""

    def __init__(self, table):
        super().__init__()
        for key, (shape, kind) in table.items():
            parts = key.split(".")
            node = self
            for name in parts[:-1]:
                if not hasattr(node, name):
                    node.add_module(name, _Node())
                node = getattr(node, name)
            t = _init_tensor(shape, kind)
            if kind.startswith("buf"):
                node.register_buffer(parts[-1], t)
            else:
                node.register_parameter(parts[-1], nn.Parameter(t))
        self._pack = {}
        self._cx = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.refresh())

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self.refresh()
        return out

    def refresh(self):
        """Drop packed weights (call after mutating parameters in place)."""
        self._pack = {}
        self._cx = None

    def _context(self):
        if self._cx is None:
            self._cx = _Ctx(self)
        return self._cx


class _Ctx:
    """Weights view used by the functional network walkers below."""

    def __init__(self, net, prefix=""):
        self.sd = {prefix + k: v.detach() for k, v in net.state_dict(keep_vars=True).items()}
        self.pack = net._pack

    def get(self, key):
        return self.sd[key]

    def cached(self, name, fn):
        t = self.pack.get(name)
        if t is None:
            with torch.no_grad():
                t = fn()
            t = tuple(x.contiguous().float() for x in t) if isinstance(t, tuple) else t.contiguous().float()
            self.pack[name] = t
        return t

    # ---- packers ------------------------------------------------------------------------------------
    def wt(self, key):
        """1x1 conv weight [N,K,1,1] -> [K][N]."""
        return self.cached("wt:" + key, lambda: self.sd[key].flatten(1).t())

    def packed(self, key):
        """1x1 conv weight packed for the tcgen05 kernel: (bpack, N, Nc, nchunks)."""
        name = "mma:" + key
        t = self.pack.get(name)
        if t is None:
            with torch.no_grad():
                t = packing.pack_weight(self.sd[key].flatten(1))
            self.pack[name] = t
        return t

    def packed_grouped(self, key, e):
        """FDSA project_out weight in the grouped K layout of the gate prologue (packing.grouped_layout)."""
        name = "mma_grouped:" + key
        t = self.pack.get(name)
        if t is None:
            with torch.no_grad():
                t = packing.pack_weight(self.sd[key].flatten(1), grouped_e=e)
            self.pack[name] = t
        return t

    def packed_fuse_out(self, p):
        name = "mma_fuse_out:" + p
        t = self.pack.get(name)
        if t is None:
            wt, bias = self.fuse_out(p)
            with torch.no_grad():
                t = packing.pack_weight(wt.t().contiguous())
            self.pack[name] = t
        return t

    def packed_conv3(self, key):
        """Dense 3x3 conv weight packed for the tensor-core kernel (packing.pack_conv3x3)."""
        name = "mma3x3:" + key
        t = self.pack.get(name)
        if t is None:
            with torch.no_grad():
                t = packing.pack_conv3x3(self.sd[key])
            self.pack[name] = t
        return t

    def flat(self, key):
        return self.cached("flat:" + key, lambda: self.sd[key].reshape(self.sd[key].shape[0], -1))

    def plain(self, key):
        return self.cached("plain:" + key, lambda: self.sd[key])

    def ln(self, p):
        return self.plain(p + "body.weight"), self.plain(p + "body.bias")

    def ln3(self, p):
        def f():
            g = torch.stack([self.sd[p + "norm%d.body.weight" % i] for i in (1, 2, 3)])
            b = torch.stack([self.sd[p + "norm%d.body.bias" % i] for i in (1, 2, 3)])
            return g, b
        return self.cached("ln3:" + p, f)

    def ffn_spec(self, p):
        """ffta * exp(-i fftp) as [C][8][5][2]."""
        def f():
            a, ph = self.sd[p + "ffta"].double().flatten(1), self.sd[p + "fftp"].double().flatten(1)
            return torch.stack((a * torch.cos(ph), -a * torch.sin(ph)), -1)
        return self.cached("ffnspec:" + p, f)

    def film(self, p, name):
        """conv3_x(conv1_x(img)) folded into one dense 3->C 3x3 conv (both are bias-free)."""
        def f():
            w1 = self.sd[p + "conv1_%s.weight" % name].double()[:, :, 0, 0]      # [C,3]
            w3 = self.sd[p + "conv3_%s.weight" % name].double()[:, 0]            # [C,3,3]
            return w1[:, :, None, None] * w3[:, None, :, :]
        return self.cached("film:" + p + name, f)

    def fuse_out(self, p):
        """conv2 followed by split/add folded: rows n and n+n_feat of weight and bias are summed."""
        def f():
            w = self.sd[p + "conv2.weight"].double().flatten(1)
            b = self.sd[p + "conv2.bias"].double()
            n = w.shape[0] // 2
            return (w[:n] + w[n:]).t(), b[:n] + b[n:]
        return self.cached("fuse_out:" + p, f)

    def spec_mlp(self, p):
        def f():
            parts = []
            for proc in ("process1", "process2"):
                for i in ("0", "2"):
                    parts.append(self.sd[p + "%s.%s.weight" % (proc, i)].flatten())
                    parts.append(self.sd[p + "%s.%s.bias" % (proc, i)].flatten())
            return torch.cat(parts)
        return self.cached("specmlp:" + p, f)

    def ffuse_pre(self, p):
        """fpre.0 (1x1) followed by fpre.1 (depthwise 1x1, padding 1): fold the per-channel scale into weight and bias."""
        def f():
            w0 = self.sd[p + "fpre.0.weight"].double().flatten(1)
            b0 = self.sd[p + "fpre.0.bias"].double()
            s = self.sd[p + "fpre.1.weight"].double().flatten()
            b1 = self.sd[p + "fpre.1.bias"].double()
            return (w0 * s[:, None]).t(), b0 * s + b1, b1
        return self.cached("ffuse:" + p, f)

    def bn_fold(self, conv_p, bn_p, eps=1e-5):
        def f():
            w = self.sd[conv_p + "weight"].double()
            g, b = self.sd[bn_p + "weight"].double(), self.sd[bn_p + "bias"].double()
            m, v = self.sd[bn_p + "running_mean"].double(), self.sd[bn_p + "running_var"].double()
            s = g / torch.sqrt(v + eps)
            return w * s.view(-1, 1, 1, 1), b - m * s
        return self.cached("bn:" + conv_p, f)


def _new(like, *shape):
    return torch.empty(shape, dtype=torch.float32, device=like.device)


# =====================================================================================================
# FDformer blocks                                                    (FDN_arch.py:381-475, 556-695)
# =====================================================================================================
def _gemm_mode():
    """'tf32x3' (default: tcgen05 with the 3xTF32 split, fp32-level accuracy), 'tf32' (single-pass TF32, reported
    separately) or 'ffma' (CUDA-core kernel).  The tensor-core kernel needs the device build of the library."""
    mode = os.environ.get("FDN_B200_GEMM", "tf32x3")
    if mode not in ("tf32x3", "tf32", "ffma"):
        raise RuntimeError("FDN_B200_GEMM must be tf32x3, tf32 or ffma")
    if mode != "ffma" and not ops.has_tcgen05():
        # never downgrade silently: the only library without the tensor-core kernel is the host emulation build of tests/emu
        raise RuntimeError("FDN_B200_GEMM=%s needs the sm_100a build of libfdn_b200.so (tcgen05 kernels); set FDN_B200_GEMM=ffma "
                           "explicitly to run the CUDA-core GEMM kernel" % mode)
    return mode


_MMA_OK = {}


def _mma_ok(k, n, prologue=0, stats_in_kernel=False):
    """Layers the tensor-core kernel cannot hold (K > 512, or a tile that does not fit in shared memory: FDformer dim 48 at level 3)
    run on the CUDA-core GEMM kernel of the same library - never on anything else."""
    key = (k, n, prologue, stats_in_kernel)
    ok = _MMA_OK.get(key)
    if ok is None:
        ok = _MMA_OK[key] = ops.pw_mma_supported(k, n, prologue, stats_in_kernel)
    return ok


def _conv1x1(cx, srcs, key, out, ln=None, bias=None, film=None, res=None):
    """1x1 convolution of the FDformer blocks: tcgen05 kernel, or the FFMA kernel when FDN_B200_GEMM=ffma."""
    mode = _gemm_mode()
    if mode == "ffma" or not _mma_ok(sum(t.shape[1] for t in srcs), out.shape[1], 1 if ln else 0):
        ops.pw_conv([(t, 0) for t in srcs], cx.wt(key), out, bias=bias, ln=ln, film=film, res=res, res_coef=1.0)
    else:
        ops.pw_mma(srcs, cx.packed(key), out, prologue=1 if ln else 0, ln=ln, bias=bias, film=film, res=res, res_coef=1.0,
                   passes=1 if mode == "tf32" else 3)


def _fdsa(cx, x, p):
    """x + FDSA(LN1(x)).  p = block prefix."""
    b, c, h, w = x.shape
    e = schema.expand_dim(c)
    hid = _new(x, b, 4 * e, h, w)
    _conv1x1(cx, [x], p + "attn.to_hidden.weight", hid, ln=cx.ln(p + "norm1."))
    o, vv = _new(x, b, 3 * e, h, w), _new(x, b, e, h, w)
    # depthwise 3x3, 8x8 rFFT, bin algebra and inverse FFT in one kernel; vv = convolved v_value group
    ops.fdsa_patch_dw(hid, cx.flat(p + "attn.to_hidden_dw.weight"), cx.flat(p + "attn.fft"), o, vv)
    del hid
    g3, b3 = cx.ln3(p + "attn.")
    out = _new(x, b, c, h, w)
    mode = _gemm_mode()
    in_kernel = not (e > 40 or os.environ.get("FDN_B200_GATE_STATS") == "prepass")
    if mode == "ffma" or not _mma_ok(3 * e, c, 2, in_kernel):
        ops.chan_ln(o, o, g3, b3, groups=3, mul=vv, mul_bs=e * h * w)
        ops.pw_conv([(o, 0)], cx.wt(p + "attn.project_out.weight"), out, res=x, res_coef=1.0)
    else:   # norm1..3, the v_value gate and project_out in one kernel.  LayerNorm statistics: computed by the kernel's producers when
        # all K blocks of a pixel tile fit its shared-memory ring (E <= 40, i.e. level 1), else from a small pre-pass
        stats = None
        if e > 40 or os.environ.get("FDN_B200_GATE_STATS") == "prepass":
            stats = _new(x, b, 3, 2, h * w)
            ops.group_stats(o, stats, 3)
        ops.pw_mma([o], cx.packed_grouped(p + "attn.project_out.weight", e), out, prologue=2, ln=(g3, b3), aux=vv, aux_bs=e * h * w,
                   stats=stats, res=x, res_coef=1.0, passes=1 if mode == "tf32" else 3)
    return out


def _fdffn(cx, x, p):
    """x + FDFFN(LN2(x))."""
    b, c, h, w = x.shape
    hd = schema.ffn_hidden(c)
    hid = _new(x, b, hd, h, w)
    _conv1x1(cx, [x], p + "ffn.project_in.weight", hid, ln=cx.ln(p + "norm2."))
    s1 = _new(x, b, hd, h, w)
    if os.environ.get("FDN_B200_FDFFN_FUSED", "1") != "0":
        # dw3x3 -> GELU -> dw3x3, the 8x8 patch-FFT branch and their sum in one kernel, one thread per (channel, patch) with rolling rows
        # in registers: the Hd-channel s1 tensor never exists in HBM (one write + one read less per FDFFN).  Same-box A/B on B200:
        # 348.8 vs 357.9 ms per 8-image step against the two kernels below (FDN_B200_FDFFN_FUSED=0).
        ops.fdffn_spatial(hid, cx.flat(p + "ffn.space.0.weight"), cx.flat(p + "ffn.space.2.weight"), cx.ffn_spec(p + "ffn."), s1)
        s2 = hid
    else:
        ops.dwconv3(hid, cx.flat(p + "ffn.space.0.weight"), s1, mode=1)
        t = _new(x, b, hd, h, w)
        # spectral branch + space.2 evaluated on the patch (one-pixel halo of s1): s2 never goes to HBM
        ops.fdffn_patch_dw(hid, s1, cx.flat(p + "ffn.space.2.weight"), cx.ffn_spec(p + "ffn."), t)
        s1, s2 = t, hid
    ops.dwconv3(s1, cx.flat(p + "ffn.dwconv.weight"), s2, mode=2)  # s2 <- gelu(x1) * x2
    out = _new(x, b, c, h, w)
    _conv1x1(cx, [s2], p + "ffn.project_out.weight", out, res=x)
    return out


def _fcaffn(cx, x, side, p):
    """x + FCAFFN(LN3(x), amp, pha, img).  side = (amp [B,3,H,Wf], pha [B,3,H,Wf], img [B,3,H,W])."""
    amp, pha, img = side
    b, c, h, w = x.shape
    wf = w // 2 + 1
    x1 = _new(x, b, c, h, w)
    g, bt = cx.ln(p + "norm3.")
    ops.chan_ln(x, x1, g, bt)
    spec = _new(x, b, c, h, wf, 2)
    ops.fft_rows_r2c(x1, spec)
    ops.fft_cols(spec, h * wf, wf, spec, h * wf, wf, b * c, h, wf, w, ops.COLS_FWD_MOD_INV, c, amp, pha,
                 cx.flat(p + "ffn2.conv1_xa.weight"), cx.flat(p + "ffn2.conv1_xp.weight"))
    y = _new(x, b, c, h, w)
    ops.fft_rows_c2r(spec, y, 1.0 / (h * w))
    del spec
    fmul, fadd = _new(x, b, c, h, w), _new(x, b, c, h, w)
    ops.film_maps(img, cx.film(p + "ffn2.", "mul"), cx.film(p + "ffn2.", "add"), fmul, fadd)
    t = _new(x, b, c, h, w)
    mode = _gemm_mode()
    if mode == "ffma" or not _mma_ok(c, c, 3):
        g, bt = cx.ln(p + "ffn2.norm.")
        ops.chan_ln(y, y, g, bt, mul=x1, mul_bs=c * h * w, add=x1, add_bs=c * h * w)
        ops.pw_conv([(y, 0)], cx.wt(p + "ffn2.project_in.weight"), t, film=(fmul, fadd))
    else:   # LN(irfft)*x1 + x1, project_in and the FiLM in one kernel
        ops.pw_mma([y], cx.packed(p + "ffn2.project_in.weight"), t, prologue=3, ln=cx.ln(p + "ffn2.norm."), aux=x1,
                   aux_bs=c * h * w, film=(fmul, fadd), passes=1 if mode == "tf32" else 3)
    ops.dwconv3(t, cx.flat(p + "ffn2.dwconv.weight"), y, mode=2)
    out = fmul
    _conv1x1(cx, [y], p + "ffn2.project_out.weight", out, res=x)
    return out


_NVTX = os.environ.get("FDN_B200_NVTX") == "1"      # NVTX ranges per sub-block (nsys / ncu --nvtx), off by default


@contextlib.contextmanager
def _range(name):
    if _NVTX:
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        yield


def _tblock(cx, x, side, p):
    if (p + "attn.fft") in cx.sd:
        with _range("FDSA " + p):
            x = _fdsa(cx, x, p)
    with _range("FDFFN " + p):
        x = _fdffn(cx, x, p)
    if (p + "ffn2.project_in.weight") in cx.sd:
        with _range("FCAFFN " + p):
            x = _fcaffn(cx, x, side, p)
    return x


def _stage(cx, x, side, p):
    i = 0
    while (p + "%d.norm2.body.weight" % i) in cx.sd:
        x = _tblock(cx, x, side, p + "%d." % i)
        i += 1
    return x


def _fuse(cx, enc, dec, p):
    b, n, h, w = enc.shape
    x = _new(enc, b, 2 * n, h, w)
    _conv1x1(cx, [enc, dec], p + "conv.weight", x, bias=cx.plain(p + "conv.bias"))
    x = _tblock(cx, x, None, p + "att_channel.")
    wt, bias = cx.fuse_out(p)
    out = _new(enc, b, n, h, w)
    mode = _gemm_mode()
    if mode == "ffma" or not _mma_ok(2 * n, n, 0):
        ops.pw_conv([(x, 0)], wt, out, bias=bias)
    else:
        ops.pw_mma([x], cx.packed_fuse_out(p), out, bias=bias, passes=1 if mode == "tf32" else 3)
    return out


def _conv3x3(cx, x, key, out):
    """Dense 3x3 conv (stride 1, padding 1, no bias) of the resamplers: tensor-core implicit GEMM (3xTF32) when the channel counts
    allow it, else the FFMA kernel (also under FDN_B200_GEMM=ffma)."""
    cout, cin = cx.sd[key].shape[:2]
    if _gemm_mode() != "ffma" and cin % 8 == 0 and packing.conv3x3_cn(cout):
        ops.conv3x3_mma(x, cx.packed_conv3(key), out)
    else:
        ops.conv2d(x, cx.plain(key), out, pad=1)


def _down(cx, x, key):
    b, c, h, w = x.shape
    t = _new(x, b, c, h // 2, w // 2)
    ops.avgpool2(x, t)
    out = _new(x, b, 2 * c, h // 2, w // 2)
    _conv3x3(cx, t, key, out)
    return out


def _up(cx, x, key):
    b, c, h, w = x.shape
    t = _new(x, b, c, 2 * h, 2 * w)
    ops.up2_bilinear(x, t)
    out = _new(x, b, c // 2, 2 * h, 2 * w)
    _conv3x3(cx, t, key, out)
    return out


def _fdformer(cx, img, side1, side2, side3, p, ori=None):
    b, _, h, w = img.shape
    c = cx.sd[p + "patch_embed.proj.weight"].shape[0]
    x1 = _new(img, b, c, h, w)
    ops.conv2d(img, cx.plain(p + "patch_embed.proj.weight"), x1, pad=1)
    x1 = _stage(cx, x1, side1, p + "encoder_level1.")
    x2 = _down(cx, x1, p + "down1_2.body.1.weight")
    x2 = _stage(cx, x2, side2, p + "encoder_level2.")
    x3 = _down(cx, x2, p + "down2_3.body.1.weight")
    x3 = _stage(cx, x3, side3, p + "encoder_level3.")
    x3 = _stage(cx, x3, side3, p + "decoder_level3.")
    y2 = _up(cx, x3, p + "up3_2.body.1.weight")
    del x3
    y2 = _fuse(cx, y2, x2, p + "fuse2.")
    del x2
    y2 = _stage(cx, y2, side2, p + "decoder_level2.")
    y1 = _up(cx, y2, p + "up2_1.body.1.weight")
    del y2
    y1 = _fuse(cx, y1, x1, p + "fuse1.")
    del x1
    y1 = _stage(cx, y1, side1, p + "decoder_level1.")
    y1 = _stage(cx, y1, side1, p + "refinement.")
    out = _new(img, b, cx.sd[p + "output.weight"].shape[0], h, w)
    ops.conv2d(y1, cx.plain(p + "output.weight"), out, res=(img if ori is None else ori), pad=1)
    return out


# =====================================================================================================
# MAR                                                                 (FDN_arch.py:75-286)
# =====================================================================================================
def _spectral_mlp_fwd(cx, x, p):
    """rows R2C + forward columns + per-bin MLPs, in place.  Returns spec [B,nc,H,Wf,2]."""
    b, nc, h, w = x.shape
    wf = w // 2 + 1
    spec = _new(x, b, nc, h, wf, 2)
    ops.fft_rows_r2c(x, spec)
    ops.fft_cols(spec, h * wf, wf, spec, h * wf, wf, b * nc, h, wf, w, ops.COLS_FWD)
    ops.spec_mlp(spec, h * wf, h * wf, b, nc, cx.spec_mlp(p))
    return spec


def _process_block(cx, x, p, variant, img_scale=None):
    b, nc, h, w = x.shape
    wf = w // 2 + 1
    f = _new(x, b, nc, h, w)
    fp = p + "frequency_process."
    ops.pw_conv([(x, 0)], cx.wt(fp + "fpre.weight"), f, bias=cx.plain(fp + "fpre.bias"))
    spec = _spectral_mlp_fwd(cx, f, fp)
    ops.fft_cols(spec, h * wf, wf, spec, h * wf, wf, b * nc, h, wf, w, ops.COLS_INV)
    if variant == "lolv1":      # cat(irfft + x) + x          fdnlol24_arch.py:769-776
        ops.fft_rows_c2r(spec, f, 1.0 / (h * w), res=x, res_coef=1.0, planes_per_image=nc)
        out = _new(x, b, nc, h, w)
        ops.pw_conv([(f, 0)], cx.wt(p + "cat.weight"), out, bias=cx.plain(p + "cat.bias"), res=x, res_coef=1.0,
                    img_scale=img_scale)
        return out
    # irfft + x + x                                            FDN_arch.py:100,118
    ops.fft_rows_c2r(spec, f, 1.0 / (h * w), res=x, res_coef=2.0, img_scale=img_scale, planes_per_image=nc)
    return f


def _fourier_fuse(cx, srcs, h, w, p):
    """srcs: [(tensor, shift)] x3 concatenated at (h, w)."""
    b = srcs[0][0].shape[0]
    wt, bias, border = cx.ffuse_pre(p)
    nc = wt.shape[1]
    hp, wp = h + 2, w + 2
    wpf, wf = wp // 2 + 1, w // 2 + 1
    ypad = _new(srcs[0][0], b, nc, hp, wp)
    ops.pw_conv(srcs, wt, ypad.view(-1)[wp + 1:], bias=bias, out_view=(h, w, nc * hp * wp, hp * wp, wp))
    ops.fill_border(ypad, border, nc)
    spec = _spectral_mlp_fwd(cx, ypad, p)
    spec2 = _new(ypad, b, nc, h, wf, 2)
    # irfft2(s=(h, w)) of a larger spectrum slices it to [:h, :w//2+1]
    ops.fft_cols(spec, hp * wpf, wpf, spec2, h * wf, wf, b * nc, h, wf, w, ops.COLS_INV)
    y = _new(ypad, b, nc, h, w)
    ops.fft_rows_c2r(spec2, y, 1.0 / (h * w))
    out = _new(ypad, b, nc, h, w)
    ops.conv2d(y, cx.plain(p + "fourier_out.weight"), out, bias=cx.plain(p + "fourier_out.bias"), pad=1)
    return out


def _conv_b(cx, x, p, cout, stride=1, pad=1, act=0, res=None, res_shift=0, head=0):
    b, _, h, w = x.shape
    k = cx.sd[p + "weight"].shape[-1]
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    out = _new(x, b, cout, ho, wo)
    ops.conv2d(x, cx.plain(p + "weight"), out, bias=cx.plain(p + "bias"), res=res, res_shift=res_shift, stride=stride, pad=pad,
               act=act, head=head)
    return out


def _pw_b(cx, srcs, p, h, w, act=0):
    wt = cx.wt(p + "weight")
    out = _new(srcs[0][0], srcs[0][0].shape[0], wt.shape[1], h, w)
    ops.pw_conv(srcs, wt, out, bias=cx.plain(p + "bias"), act=act)
    return out


def _mar_core(cx, x, ratio, p, variant, use_ratio):
    b, _, h, w = x.shape
    sc = ratio if use_ratio else None
    xp2 = _new(x, b, 12, h // 2, w // 2)
    ops.pixel_unshuffle(x, xp2, 2)
    xp4 = _new(x, b, 48, h // 4, w // 4)
    ops.pixel_unshuffle(x, xp4, 4)
    z2 = _process_block(cx, _pw_b(cx, [(xp2, 0)], p + "f2.0.", h // 2, w // 2), p + "f2.1.", variant, sc)
    z4 = _process_block(cx, _pw_b(cx, [(xp4, 0)], p + "f1.0.", h // 4, w // 4), p + "f1.1.", variant, sc)
    x_ = _process_block(cx, _pw_b(cx, [(x, 0)], p + "f3.0.", h, w), p + "f3.1.", variant, sc)
    res1 = _process_block(cx, x_, p + "Encoder.0.", variant)
    z = _conv_b(cx, res1, p + "f3_down.main.0.", 24, stride=2, act=1)
    z = _conv_b(cx, _pw_b(cx, [(z, 0), (z2, 0)], p + "FAM2.merge1.", h // 2, w // 2), p + "FAM2.merge2.", 24)
    res2 = _process_block(cx, z, p + "Encoder.1.", variant)
    z = _conv_b(cx, res2, p + "f2_down.main.0.", 48, stride=2, act=1)
    z = _conv_b(cx, _pw_b(cx, [(z, 0), (z4, 0)], p + "FAM1.merge1.", h // 4, w // 4), p + "FAM1.merge2.", 48)
    z = _process_block(cx, z, p + "Encoder.2.", variant)

    res2f = _fourier_fuse(cx, [(res1, -1), (res2, 0), (z, 1)], h // 2, w // 2, p + "AFFs.1.")
    res1f = _fourier_fuse(cx, [(res1, 0), (res2, 1), (z, 2)], h, w, p + "AFFs.0.")

    z = _process_block(cx, z, p + "Decoder.0.", variant)
    i3 = _conv_b(cx, z, p + "ConvsOut.0.main.0.", 3, res=x, res_shift=2, head=1)
    zu = _new(x, b, 24, h // 2, w // 2)
    ops.convt4s2(z, cx.plain(p + "f2_up.main.0.weight"), cx.plain(p + "f2_up.main.0.bias"), zu, act=1)
    z = _pw_b(cx, [(zu, 0), (res2f, 0)], p + "Convs.0.main.0.", h // 2, w // 2, act=1)
    z = _process_block(cx, z, p + "Decoder.1.", variant)
    i2 = _conv_b(cx, z, p + "ConvsOut.1.main.0.", 3, res=x, res_shift=1, head=1)
    zu = _new(x, b, 12, h, w)
    ops.convt4s2(z, cx.plain(p + "f3_up.main.0.weight"), cx.plain(p + "f3_up.main.0.bias"), zu, act=1)
    z = _pw_b(cx, [(zu, 0), (res1f, 0)], p + "Convs.1.main.0.", h, w, act=1)
    z = _process_block(cx, z, p + "Decoder.2.", variant)
    i1 = _conv_b(cx, z, p + "out.main.0.", 3, res=x, res_shift=0, head=1)
    return i3, i2, i1


def _pyramid(x):
    b, c, h, w = x.shape
    x2 = _new(x, b, c, h // 2, w // 2)
    ops.avgpool2(x, x2)
    x3 = _new(x, b, c, h // 4, w // 4)
    ops.avgpool2(x2, x3)
    return x, x2, x3


def _mar(cx, x, ratio, p, variant, use_ratio=True, pyr=None):
    """ratio: flat [B] tensor.  Returns (1/4, 1/2, 1) gamma-corrected images."""
    i3, i2, i1 = _mar_core(cx, x, ratio, p + "net.", variant, use_ratio)
    x1, x2, x3 = pyr if pyr is not None else _pyramid(x)
    outs = []
    for xi, ii in ((x3, i3), (x2, i2), (x1, i1)):
        o = torch.empty_like(xi)
        ops.gamma_curve(xi, ii, o, 40.0)
        outs.append(o)
    return tuple(outs)


# =====================================================================================================
# FDN                                                                 (FDN_arch.py:869-921)
# =====================================================================================================
def _spectral_map(cx, x, norm_p, mode):
    """LN(3) -> rfft2 -> angle(rd(.)) or abs : [B,3,H,Wf] real map."""
    b, c, h, w = x.shape
    wf = w // 2 + 1
    n = torch.empty_like(x)
    g, bt = cx.ln(norm_p)
    ops.chan_ln(x, n, g, bt)
    spec = _new(x, b, c, h, wf, 2)
    ops.fft_rows_r2c(n, spec)
    out = _new(x, b, c, h, wf)
    ops.fft_cols(spec, h * wf, wf, out, h * wf, wf, b * c, h, wf, w, mode)
    return out


def _fdn(cx, img, ratio, variant):
    pyr = _pyramid(img)
    norms = ("norm1.", "norm2.", "norm3.")
    with _range("FDN prologue (phase maps)"):
        pha = [_spectral_map(cx, t, n, ops.COLS_FWD_ANGLE) for t, n in zip(pyr, norms)]
    with _range("MAR"):
        q3, q2, q1 = _mar(cx, img, ratio, "net_a.", variant, True, pyr)
    with _range("FDN prologue (amplitude maps)"):
        amp = [_spectral_map(cx, t, n, ops.COLS_FWD_ABS) for t, n in zip((q1, q2, q3), norms)]
    with _range("FDformer"):
        out = _fdformer(cx, img, (amp[0], pha[0], q1), (amp[1], pha[1], q2), (amp[2], pha[2], q3), "net_p.")
    return out, q1, q2, q3


def _on_device(x):
    """Make the tensor's GPU the current device for the launches below (streams, twiddle tables and function attributes of the
    library are per device); CPU tensors only occur under the host emulation build of tests/emu."""
    return torch.cuda.device(x.device) if x.is_cuda else contextlib.nullcontext()


def _flat_ratio(ratio, b, dev):
    """[B,1] / [B,1,1,1] / scalar ratio -> flat [B] fp32 on the input's device (the reference broadcasts a single value)."""
    r = ratio.detach().float().reshape(-1).to(dev)
    if r.numel() == 1 and b > 1:
        r = r.expand(b)
    if r.numel() != b:
        raise RuntimeError("ratio must hold one value per image (or a single value): got %d values for %d images" % (r.numel(), b))
    return r.contiguous()


def _side(t, name, shape, dev):
    if t is None:
        raise RuntimeError("%s is required" % name)
    if not t.is_cuda or t.device != dev:
        raise RuntimeError("%s must be on %s (got %s)" % (name, dev, t.device))
    if tuple(t.shape) != tuple(shape):
        raise RuntimeError("%s must have shape %s, got %s" % (name, tuple(shape), tuple(t.shape)))
    return t.detach().float().contiguous()


def _check_input(x, multiple):
    ops._device_ok(x)
    if x.dim() != 4 or x.shape[1] != 3:
        raise RuntimeError("expected a [B,3,H,W] image tensor, got %s" % (tuple(x.shape),))
    if x.shape[2] % multiple or x.shape[3] % multiple:
        raise RuntimeError("H and W must be multiples of %d (pad the input as the inference scripts do), got %dx%d"
                           % (multiple, x.shape[2], x.shape[3]))
    return x.detach().float().contiguous()


def _micro_batch(b, h, w):
    env = os.environ.get("FDN_B200_MICRO_BATCH")
    if env:
        return max(1, min(b, int(env)))
    limit = max(1, min(b, 8, (8 * 1024 * 1024) // (h * w)))
    while b % limit:          # equal chunks: one CUDA-graph / workspace shape per call
        limit -= 1
    return limit


class _FDNBase(_Net):
    _dim = 32
    _variant = "lolblur"

    def __init__(self):
        super().__init__(schema.fdn_schema(self._dim))
        for k, prm in self.named_parameters():
            if k.startswith("net_a."):
                prm.requires_grad = False

    @torch.no_grad()
    def forward(self, inp_img, ori=None, device=None, ratio_i=None, mode=1):
        x = _check_input(inp_img, 32)
        if ratio_i is None:
            raise RuntimeError("ratio_i ([B,1] tensor) is required")       # the reference dereferences None here too
        ratio = _flat_ratio(ratio_i, x.shape[0], x.device)
        cx = self._context()
        b, _, h, w = x.shape
        mb = _micro_batch(b, h, w)
        outs = []
        with _on_device(x):       # kernels launch on the input's device, whatever the caller's current device is
            for s in range(0, b, mb):
                outs.append(_fdn(cx, x[s:s + mb].contiguous(), ratio[s:s + mb].contiguous(), self._variant))
        res = [torch.cat([o[i] for o in outs], 0) if len(outs) > 1 else outs[0][i] for i in range(4)]
        if self._variant == "lolv1":
            return res[0], res[0], res[0], res[0]
        return tuple(res)


class FDN(_FDNBase):
    """LOL-Blur network (FDN_arch.py:847-921): MAR + FDformer(dim=32)."""
    _dim = 32
    _variant = "lolblur"


class FDN_lolv1(_FDNBase):
    """LOL-v1 network (fdnlol24_arch.py:951-1033): MAR (cat conv live) + FDformer(dim=24); returns (out,)*4."""
    _dim = 24
    _variant = "lolv1"


class FDformer(_Net):
    """FDN_arch.py:753-842.  Stand-alone phase/restoration network; side maps are supplied by the caller."""

    def __init__(self, inp_channels=3, out_channels=3, dim=48, num_blocks=[6, 6, 12, 8], num_refinement_blocks=4,
                 ffn_expansion_factor=3, bias=False):
        if bias:
            raise NotImplementedError("the reference only instantiates FDformer with bias=False")
        super().__init__(schema.fdformer_schema(dim, tuple(num_blocks), num_refinement_blocks, inp_channels, out_channels))

    @torch.no_grad()
    def forward(self, inp_img, ori_img=None, x_high1=None, x_high2=None, x_high3=None, x_high12=None, x_high22=None,
                x_high32=None, x1=None, x2=None, x3=None):
        x = _check_input(inp_img, 32)
        b, _, h, w = x.shape
        sides = []
        for lvl, (amp, pha, img) in enumerate(((x_high1, x_high12, x1), (x_high2, x_high22, x2), (x_high3, x_high32, x3))):
            hl, wl = h >> lvl, w >> lvl
            sides.append((_side(amp, "x_high%d" % (lvl + 1), (b, 3, hl, wl // 2 + 1), x.device),
                          _side(pha, "x_high%d2" % (lvl + 1), (b, 3, hl, wl // 2 + 1), x.device),
                          _side(img, "x%d" % (lvl + 1), (b, 3, hl, wl), x.device)))
        ori = None if ori_img is None else _side(ori_img, "ori_img", x.shape, x.device)
        cx = self._context()
        with _on_device(x):
            return _fdformer(cx, x, sides[0], sides[1], sides[2], "", ori)


class MAR(_Net):
    """FDN_arch.py:261-286 / mar_arch.py:255-283 / fdnlol24_arch.py:211-248.

    ``variant='lolv1'`` selects the fdnlol24_arch ProcessBlock (its ``cat`` 1x1 conv is applied)."""

    def __init__(self, use_ratio=True, variant="lolblur"):
        super().__init__(schema.mar_schema())
        self.use_ratio = use_ratio
        self.variant = variant
        self.scale = 40.0

    # FDN_arch.MAR ignores its use_ratio argument (MAR_archa(use_ratio=True) and an unconditional multiply,
    # FDN_arch.py:213-219,264); mar_arch / fdnlol24_arch honour it.  The FDN_arch shim sets this to True.
    _always_ratio = False

    @torch.no_grad()
    def forward(self, x, ratio=None):
        x = _check_input(x, 4)
        use_ratio = self.use_ratio or self._always_ratio
        r = None
        if use_ratio:
            if ratio is None:
                raise RuntimeError("ratio is required")
            r = _flat_ratio(ratio, x.shape[0], x.device)
        with _on_device(x):
            return _mar(self._context(), x, r, "", self.variant, use_ratio)


class I_predict_net(_Net):
    """LPNet luminance predictor (LPNet_arch.py:86-134)."""

    def __init__(self, c=16):
        super().__init__(schema.lpnet_schema(c))
        self.c = c

    @torch.no_grad()
    def forward(self, x, use_ori_i=False):
        ops._device_ok(x)
        x = x.detach().float().contiguous()
        with _on_device(x):
            return self._forward(x, use_ori_i)

    def _forward(self, x, use_ori_i):
        cx = self._context()
        b, _, h, w = x.shape
        gray = None
        if use_ori_i:
            gray = _new(x, b)
            ops.gray_mean(x, gray)
        wgt, bias = cx.bn_fold("conv1.0.", "conv1.1.")
        c = self.c
        h1, w1 = (h + 6 - 7) // 2 + 1, (w + 6 - 7) // 2 + 1
        y = _new(x, b, c, h1, w1)
        ops.conv2d(x, wgt, y, bias=bias, stride=2, pad=3, act=2)
        h2, w2 = (h1 - 1) // 2 + 1, (w1 - 1) // 2 + 1
        t = _new(x, b, c, h2, w2)
        ops.avgpool3s2(y, t)
        y = t
        cin = c
        for si, (name, num, stride) in enumerate(schema.LPNET_STAGES):
            f1, f3 = c << si, c << (si + 1)
            for i in range(num):
                y = self._se_block(cx, y, "%s.%d." % (name, i), cin, f1, f3, stride if i == 0 else 1)
                cin = f3
        m = _new(x, b, cin)
        ops.plane_mean(y, m)
        out = _new(x, b, 1)
        ops.lpnet_head(m, cx.plain("fc.0.weight"), cx.plain("fc.0.bias"), cx.plain("fc2.0.weight"), cx.plain("fc2.0.bias"),
                       gray, out)
        return out

    @staticmethod
    def _se_block(cx, x, p, cin, f1, f3, stride):
        b, _, h, w = x.shape
        ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
        w1, b1 = cx.bn_fold(p + "conv1.0.", p + "conv1.1.")
        y1 = _new(x, b, f1, ho, wo)
        if stride == 1:
            ops.pw_conv([(x, 0)], cx.cached("wt1:" + p, lambda: w1.flatten(1).t()), y1, bias=b1, act=2)
        else:
            ops.conv2d(x, w1, y1, bias=b1, stride=stride, pad=0, act=2)
        w2, b2 = cx.bn_fold(p + "conv2.0.", p + "conv2.1.")
        y2 = _new(x, b, f1, ho, wo)
        ops.conv2d(y1, w2, y2, bias=b2, stride=1, pad=1, act=2)
        w3, b3 = cx.bn_fold(p + "conv3.0.", p + "conv3.1.")
        y3 = _new(x, b, f3, ho, wo)
        ops.pw_conv([(y2, 0)], cx.cached("wt3:" + p, lambda: w3.flatten(1).t()), y3, bias=b3)
        m = _new(x, b, f3)
        ops.plane_mean(y3, m)
        s = _new(x, b, f3)
        ops.se_fc(m, cx.flat(p + "se.1.weight"), cx.plain(p + "se.1.bias"), cx.flat(p + "se.3.weight"), cx.plain(p + "se.3.bias"), s)
        if (p + "shortcut.0.weight") in cx.sd:
            ws, bs = cx.bn_fold(p + "shortcut.0.", p + "shortcut.1.")
            sc = _new(x, b, f3, ho, wo)
            if stride == 1:
                ops.pw_conv([(x, 0)], cx.cached("wts:" + p, lambda: ws.flatten(1).t()), sc, bias=bs)
            else:
                ops.conv2d(x, ws, sc, bias=bs, stride=stride, pad=0)
        else:
            sc = x
        out = _new(x, b, f3, ho, wo)
        ops.se_apply(y3, s, sc, out)
        return out
