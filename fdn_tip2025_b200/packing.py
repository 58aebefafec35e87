"""Host-side weight packing for the tcgen05 1x1-convolution kernel (csrc/pw_mma.cu).

The B operand of tcgen05.mma is read from shared memory in the canonical K-major SWIZZLE_128B layout.  Packing the
weight once (at load_state_dict / first forward) into that exact image lets the kernel fill its B stage with a linear,
coalesced copy.  Layout: [chunk][k-block of 32][hi, lo][Nc rows][32 floats], where inside a row the 16-byte group g of
row n is stored at group position g ^ (n % 8); hi = tf32(w), lo = tf32(w - hi) (the 3xTF32 split).
"""
import torch


def tf32_round(x):
    """cvt.rna.tf32.f32: round to 10 mantissa bits, ties away from zero."""
    bits = x.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


def chunking(n):
    """Output-channel chunking: nchunks chunks of Nc (multiple of 16, <= 256) columns of TMEM."""
    nchunks = (n + 255) // 256
    per = (n + nchunks - 1) // nchunks
    nc = (per + 15) // 16 * 16
    return nc, nchunks


GROUP_BLOCK = 10     # csrc/pw_mma.cu MMA_EB: channels of each LayerNorm group per 32-row K block


def grouped_layout(w, e):
    """FDSA project_out weight [N, 3E] -> [N, nkb*32] in the kernel's grouped K order: block kb holds, for the three
    LayerNorm groups g, the channels g*E + kb*10 + el (el < 10) at row g*10 + el; rows 30, 31 and missing channels are zero."""
    n, k = w.shape
    assert k == 3 * e
    nkb = (e + GROUP_BLOCK - 1) // GROUP_BLOCK
    out = torch.zeros(n, nkb * 32, dtype=w.dtype, device=w.device)
    for kb in range(nkb):
        ne = min(GROUP_BLOCK, e - kb * GROUP_BLOCK)
        for g in range(3):
            out[:, kb * 32 + g * GROUP_BLOCK: kb * 32 + g * GROUP_BLOCK + ne] = w[:, g * e + kb * GROUP_BLOCK: g * e + kb * GROUP_BLOCK + ne]
    return out


def pack_weight(w, grouped_e=None):
    """w: [N, K] fp32 -> (bpack flat fp32 tensor, N, Nc, nchunks).  grouped_e = E selects the grouped K layout of prologue 2."""
    w = w.detach().float()
    if grouped_e is not None:
        w = grouped_layout(w, grouped_e)
    n, k = w.shape
    nc, nchunks = chunking(n)
    kpad = (k + 7) // 8 * 8
    nkb = (kpad + 31) // 32
    wp = torch.zeros(nchunks * nc, nkb * 32, dtype=torch.float32, device=w.device)
    # chunk c holds output channels [c*nc, (c+1)*nc) of the *padded* numbering == real numbering (padding only at the end
    # of each chunk would renumber channels, so pad at the very end and let every chunk but the last be full)
    wp[:n, :k] = w
    hi = tf32_round(wp)
    lo = tf32_round(wp - hi)
    rows = torch.arange(nc, device=w.device)
    grp = torch.arange(8, device=w.device)
    src_grp = grp[None, :] ^ (rows[:, None] & 7)            # stored position g holds source group g ^ (n%8)
    out = []
    for t in (hi, lo):
        t = t.view(nchunks, nc, nkb, 8, 4).permute(0, 2, 1, 3, 4)          # [chunk][kb][n][group][4]
        idx = src_grp[None, None, :, :, None].expand(nchunks, nkb, nc, 8, 4)
        out.append(torch.gather(t, 3, idx))
    packed = torch.stack(out, 2).contiguous()                # [chunk][kb][2][n][8][4]
    return packed.view(-1), n, nc, nchunks


# ---------------------------------------------------------------------------------------------------- 3x3 tensor-core conv
def conv3x3_cn(cout):
    """Output channels per CTA of csrc/conv_mma.cu (fdn_conv3x3_mma_cn): 32 or 24, whichever divides Cout, else 0."""
    for c in (32, 24):
        if cout % c == 0:
            return c
    return 0


def pack_conv3x3(w):
    """w [Cout, Cin, 3, 3] -> flat fp32 [Cout/CN][Cin/8][hi, lo][tap][8 channels][COP] (COP = 40, zero padded): the
    shared-memory image of one 8-channel K chunk of k_conv3x3_mma, so the kernel stages it with a linear copy.
    hi = tf32(w), lo = tf32(w - hi)."""
    w = w.detach().float()
    cout, cin = w.shape[:2]
    cn = conv3x3_cn(cout)
    assert cn and cin % 8 == 0 and tuple(w.shape[2:]) == (3, 3)
    cop = (cn + 31) // 32 * 32 + 8
    hi = tf32_round(w.contiguous())
    lo = tf32_round(w - hi)
    out = torch.zeros(cout // cn, cin // 8, 2, 9, 8, cop, dtype=torch.float32, device=w.device)
    for hl, t in enumerate((hi, lo)):
        v = t.reshape(cout // cn, cn, cin // 8, 8, 9)                    # [z][n][chunk][c][tap]
        out[:, :, hl, :, :, :cn] = v.permute(0, 2, 4, 3, 1)             # [z][chunk][tap][c][n]
    return out.reshape(-1).contiguous()
