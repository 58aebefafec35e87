"""state_dict schema of the reference networks, stated declaratively.

The drop-in contract is: same keys, same shapes, so
``load_state_dict(ckpt["params"], strict=True)`` works with the reference checkpoints.
Nothing here builds ``nn.Conv2d`` objects; it only enumerates ``(key, shape, kind)``.

Reference constructors these tables restate:
  FDformer      basicsr/models/archs/FDN_arch.py:753-808  (fdnlol24_arch.py:826-881)
  TransformerBlock / FDSA / FDFFN / FCAFFN   FDN_arch.py:646-664, 556-573, 430-452, 381-402
  Fuse          FDN_arch.py:679-686
  MAR / MAR_archa / ProcessBlock / FreBlock / fourier_fuse / FAM / BasicConv
                FDN_arch.py:261-267, 149-201, 101-107, 75-86, 121-135, 52-57, 16-29
  FDN / FDN_lolv1   FDN_arch.py:847-867, fdnlol24_arch.py:951-979
  I_predict_net / SEBlock   LPNet_arch.py:86-112, 42-68
"""
from collections import OrderedDict

PATCH = 8
PATCH_BINS = PATCH // 2 + 1


def expand_dim(dim):
    """FDSA hidden width per q/k/v/v_value group (FDN_arch.py:561)."""
    return int(dim * 1.2)


def ffn_hidden(dim):
    """FDFFN hidden width (FDN_arch.py:434, r=2.7)."""
    return int(2.7 * dim)


def _ln(out, prefix, n):
    out[prefix + "body.weight"] = ((n,), "ones")
    out[prefix + "body.bias"] = ((n,), "zeros")


def _conv(out, prefix, cout, cin, k, bias):
    out[prefix + "weight"] = ((cout, cin, k, k), "conv")
    if bias:
        out[prefix + "bias"] = ((cout,), "conv_bias:%d" % (cin * k * k))


def transformer_block(out, p, dim, att, light):
    e, hd = expand_dim(dim), ffn_hidden(dim)
    if att:
        _ln(out, p + "norm1.", dim)
        out[p + "attn.fft"] = ((e, 1, 1, PATCH, PATCH_BINS), "ones")
        _conv(out, p + "attn.to_hidden.", 4 * e, dim, 1, False)
        _conv(out, p + "attn.to_hidden_dw.", 4 * e, 1, 3, False)
        _conv(out, p + "attn.project_out.", dim, 3 * e, 1, False)
        for i in (1, 2, 3):
            _ln(out, p + "attn.norm%d." % i, e)
    _ln(out, p + "norm2.", dim)
    out[p + "ffn.ffta"] = ((hd, 1, 1, PATCH, PATCH_BINS), "ones")
    out[p + "ffn.fftp"] = ((hd, 1, 1, PATCH, PATCH_BINS), "zeros")
    _conv(out, p + "ffn.space.0.", hd, 1, 3, False)
    _conv(out, p + "ffn.space.2.", hd, 1, 3, False)
    _conv(out, p + "ffn.dwconv.", 2 * hd, 1, 3, False)
    _conv(out, p + "ffn.project_in.", hd, dim, 1, False)
    _conv(out, p + "ffn.project_out.", dim, hd, 1, False)
    if light:
        _ln(out, p + "norm3.", dim)
        _conv(out, p + "ffn2.project_in.", dim, dim, 1, False)
        _conv(out, p + "ffn2.project_out.", dim, dim, 1, False)
        for nm in ("xa", "xp", "add", "mul"):
            _conv(out, p + "ffn2.conv1_%s." % nm, dim, 3, 1, False)
        for nm in ("add", "mul"):
            _conv(out, p + "ffn2.conv3_%s." % nm, dim, 1, 3, False)
        _ln(out, p + "ffn2.norm.", dim)
        _conv(out, p + "ffn2.dwconv.", 2 * dim, 1, 3, False)


def fuse(out, p, n_feat):
    transformer_block(out, p + "att_channel.", 2 * n_feat, att=False, light=False)
    _conv(out, p + "conv.", 2 * n_feat, 2 * n_feat, 1, True)
    _conv(out, p + "conv2.", 2 * n_feat, 2 * n_feat, 1, True)


def fdformer_schema(dim=48, num_blocks=(6, 6, 12, 8), num_refinement_blocks=4,
                    inp_channels=3, out_channels=3, prefix=""):
    out = OrderedDict()
    p = prefix
    c1, c2, c3 = dim, dim * 2, dim * 4
    _conv(out, p + "patch_embed.proj.", c1, inp_channels, 3, False)
    for i in range(num_blocks[0]):
        transformer_block(out, p + "encoder_level1.%d." % i, c1, True, True)
    _conv(out, p + "down1_2.body.1.", c2, c1, 3, False)
    for i in range(num_blocks[1]):
        transformer_block(out, p + "encoder_level2.%d." % i, c2, True, True)
    _conv(out, p + "down2_3.body.1.", c3, c2, 3, False)
    for i in range(num_blocks[2]):
        transformer_block(out, p + "encoder_level3.%d." % i, c3, True, True)
    for i in range(num_blocks[2]):
        transformer_block(out, p + "decoder_level3.%d." % i, c3, True, False)
    _conv(out, p + "up3_2.body.1.", c2, c3, 3, False)
    _conv(out, p + "reduce_chan_level2.", c2, c3, 1, False)          # dead parameter, must exist
    for i in range(num_blocks[1]):
        transformer_block(out, p + "decoder_level2.%d." % i, c2, True, False)
    _conv(out, p + "up2_1.body.1.", c1, c2, 3, False)
    for i in range(num_blocks[0]):
        transformer_block(out, p + "decoder_level1.%d." % i, c1, True, False)
    for i in range(num_refinement_blocks):
        transformer_block(out, p + "refinement.%d." % i, c1, True, False)
    fuse(out, p + "fuse2.", c2)
    fuse(out, p + "fuse1.", c1)
    _conv(out, p + "output.", out_channels, c1, 3, False)
    _ln(out, p + "norm.", 3)                                          # dead parameter, must exist
    return out


def _process_block(out, p, nc):
    _conv(out, p + "frequency_process.fpre.", nc, nc, 1, True)
    for proc in ("process1", "process2"):
        _conv(out, p + "frequency_process.%s.0." % proc, nc, nc, 1, True)
        _conv(out, p + "frequency_process.%s.2." % proc, nc, nc, 1, True)
    _conv(out, p + "cat.", nc, nc, 1, True)       # dead in FDN_arch/mar_arch, live in fdnlol24_arch


def _fourier_fuse(out, p, cin, cout):
    _conv(out, p + "fpre.0.", cout, cin, 1, True)
    _conv(out, p + "fpre.1.", cout, 1, 1, True)
    for proc in ("process1", "process2"):
        _conv(out, p + "%s.0." % proc, cout, cout, 1, True)
        _conv(out, p + "%s.2." % proc, cout, cout, 1, True)
    _conv(out, p + "fourier_out.", cout, cout, 3, True)


def mar_schema(prefix=""):
    out = OrderedDict()
    p = prefix + "net."
    b = 12
    for i, nc in enumerate((b, 2 * b, 4 * b)):
        _process_block(out, p + "Encoder.%d." % i, nc)
    for i, nc in enumerate((4 * b, 2 * b, b)):
        _process_block(out, p + "Decoder.%d." % i, nc)
    _conv(out, p + "Convs.0.main.0.", 2 * b, 4 * b, 1, True)
    _conv(out, p + "Convs.1.main.0.", b, 2 * b, 1, True)
    _conv(out, p + "ConvsOut.0.main.0.", 3, 4 * b, 3, True)
    _conv(out, p + "ConvsOut.1.main.0.", 3, 2 * b, 3, True)
    _fourier_fuse(out, p + "AFFs.0.", 7 * b, b)
    _fourier_fuse(out, p + "AFFs.1.", 7 * b, 2 * b)
    _conv(out, p + "FAM1.merge1.", 4 * b, 8 * b, 1, True)
    _conv(out, p + "FAM1.merge2.", 4 * b, 4 * b, 3, True)
    _conv(out, p + "f1.0.", 4 * b, 48, 1, True)
    _process_block(out, p + "f1.1.", 4 * b)
    _conv(out, p + "f2.0.", 2 * b, 12, 1, True)
    _process_block(out, p + "f2.1.", 2 * b)
    _conv(out, p + "f3.0.", b, 3, 1, True)
    _process_block(out, p + "f3.1.", b)
    _conv(out, p + "f3_down.main.0.", 2 * b, b, 3, True)
    _conv(out, p + "f2_down.main.0.", 4 * b, 2 * b, 3, True)
    # ConvTranspose2d weights are (in, out, k, k)
    out[p + "f2_up.main.0.weight"] = ((4 * b, 2 * b, 4, 4), "conv")
    out[p + "f2_up.main.0.bias"] = ((2 * b,), "conv_bias:%d" % (2 * b * 16))
    out[p + "f3_up.main.0.weight"] = ((2 * b, b, 4, 4), "conv")
    out[p + "f3_up.main.0.bias"] = ((b,), "conv_bias:%d" % (b * 16))
    _conv(out, p + "out.main.0.", 3, b, 3, True)
    _conv(out, p + "FAM2.merge1.", 2 * b, 4 * b, 1, True)
    _conv(out, p + "FAM2.merge2.", 2 * b, 2 * b, 3, True)
    return out


def fdn_schema(dim):
    """FDN (dim=32, FDN_arch.py:851-857) and FDN_lolv1 (dim=24, fdnlol24_arch.py:963-969)."""
    out = OrderedDict()
    out.update(mar_schema("net_a."))
    out.update(fdformer_schema(dim=dim, num_blocks=(6, 6, 10), num_refinement_blocks=4,
                               prefix="net_p."))
    for i in (1, 2, 3):
        _ln(out, "norm%d." % i, 3)
    return out


def _bn(out, p, n):
    out[p + "weight"] = ((n,), "ones")
    out[p + "bias"] = ((n,), "zeros")
    out[p + "running_mean"] = ((n,), "buf_zeros")
    out[p + "running_var"] = ((n,), "buf_ones")
    out[p + "num_batches_tracked"] = ((), "buf_long")


def _se_block(out, p, cin, filters, first):
    f1, f2, f3 = filters
    _conv(out, p + "conv1.0.", f1, cin, 1, False)
    _bn(out, p + "conv1.1.", f1)
    _conv(out, p + "conv2.0.", f2, f1, 3, False)
    _bn(out, p + "conv2.1.", f2)
    _conv(out, p + "conv3.0.", f3, f2, 1, False)
    _bn(out, p + "conv3.1.", f3)
    if first:
        _conv(out, p + "shortcut.0.", f3, cin, 1, False)
        _bn(out, p + "shortcut.1.", f3)
    _conv(out, p + "se.1.", f3 // 16, f3, 1, True)
    _conv(out, p + "se.3.", f3, f3 // 16, 1, True)


LPNET_STAGES = (("conv2", 3, 1), ("conv3", 3, 2), ("conv4", 6, 6))   # (name, blocks, stride)


def lpnet_schema(c=16):
    out = OrderedDict()
    _conv(out, "conv1.0.", c, 3, 7, False)
    _bn(out, "conv1.1.", c)
    cin = c
    for si, (name, num, _stride) in enumerate(LPNET_STAGES):
        f = (c << si, c << si, c << (si + 1))
        for i in range(num):
            _se_block(out, "%s.%d." % (name, i), cin if i == 0 else f[2], f, i == 0)
        cin = f[2]
    out["fc.0.weight"] = ((8 * c, 8 * c), "linear")
    out["fc.0.bias"] = ((8 * c,), "conv_bias:%d" % (8 * c))
    out["fc2.0.weight"] = ((1, 8 * c), "linear")
    out["fc2.0.bias"] = ((1,), "conv_bias:%d" % (8 * c))
    return out
