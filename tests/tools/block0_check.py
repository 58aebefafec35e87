"""First FDSA block on a real (smooth, low-light) input: both GEMM modes against the fp64 oracle.  Dev tool, GPU only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from fdn_tip2025_b200 import archs, synth, ops
from oracle import fdn_oracle as O
h, w, b = 128, 160, 2
sd = synth.fdn_state_dict(dim=32, seed=7, damp=0.03)
sd64 = O.to_dtype(sd, torch.float64)
net = archs.FDN(); net.load_state_dict(sd, strict=True); net = net.cuda()
cx = net._context()
x = synth.low_light_images(b, h, w)
x0_64 = O.conv(x.double(), sd64, "net_p.patch_embed.proj.", padding=1)
p = "net_p.encoder_level1.0."
ref = x0_64 + O.fdsa(O.layer_norm(x0_64, sd64, p + "norm1."), sd64, p + "attn.")
hid64 = O.conv(O.conv(O.layer_norm(x0_64, sd64, p + "norm1."), sd64, p + "attn.to_hidden."), sd64, p + "attn.to_hidden_dw.", padding=1, groups=152)
x0 = x0_64.float().cuda()
for mode in ("ffma", "tf32x3", "tf32"):
    os.environ["FDN_B200_GEMM"] = mode
    out = archs._fdsa(cx, x0, p)
    torch.cuda.synchronize()
    d = (out.double().cpu() - ref).abs()
    print("%-7s FDSA block0 vs fp64: max %.2e  rel-L2 %.2e  frac>1e-5 %.2e" % (mode, d.max(), (d.norm() / ref.norm()), (d > 1e-5).double().mean()))
# fp32 oracle (CPU) as a third independent fp32 evaluation
ref32 = x0_64.float() + O.fdsa(O.layer_norm(x0_64.float(), sd, p + "norm1."), sd, p + "attn.")
d = (ref32.double() - ref).abs()
print("oracle32 FDSA block0 vs fp64: max %.2e  rel-L2 %.2e  frac>1e-5 %.2e" % (d.max(), d.norm() / ref.norm(), (d > 1e-5).double().mean()))
q64 = torch.fft.rfft2(O.to_patches(hid64[:, :38]))
print("fraction of q bins with |q| < 1e-5 * max:", (q64.abs() < 1e-5 * q64.abs().max()).double().mean().item(), " min |q| %.2e max %.2e" % (q64.abs().min(), q64.abs().max()))

# ---- stage by stage (FFMA path so each stage is a separate kernel)
print("---- stages")
os.environ["FDN_B200_GEMM"] = "ffma"
e = 38
hid = torch.empty(b, 4 * e, h, w, device="cuda")
ops.pw_conv([(x0, 0)], cx.wt(p + "attn.to_hidden.weight"), hid, ln=cx.ln(p + "norm1."))
hid_pre64 = O.conv(O.layer_norm(x0_64, sd64, p + "norm1."), sd64, p + "attn.to_hidden.")
d = (hid.double().cpu() - hid_pre64).abs(); print("to_hidden (LN+1x1): max %.2e rel-L2 %.2e" % (d.max(), d.norm() / hid_pre64.norm()))
hid_pre32 = O.conv(O.layer_norm(x0_64.float(), sd, p + "norm1."), sd, p + "attn.to_hidden.")
d = (hid_pre32.double() - hid_pre64).abs(); print("   cpu fp32          : max %.2e rel-L2 %.2e" % (d.max(), d.norm() / hid_pre64.norm()))
hid_dw = torch.empty_like(hid)
ops.dwconv3(hid, cx.flat(p + "attn.to_hidden_dw.weight"), hid_dw, mode=0)
d = (hid_dw.double().cpu() - hid64).abs(); print("dwconv: max %.2e rel-L2 %.2e" % (d.max(), d.norm() / hid64.norm()))
# feed the *exact* fp64 dw output (rounded to fp32) to the spectral kernel to isolate it
o = torch.empty(b, 3 * e, h, w, device="cuda")
ops.fdsa_patch(hid64.float().cuda(), cx.flat(p + "attn.fft"), o)
q, k, v, vv = hid64.float().double().chunk(4, 1)
q, k, v = (torch.fft.rfft2(O.to_patches(t)) for t in (q, k, v))
v = O.rd(v * sd64[p + "attn.fft"]); qkm = O.rd(q * k).abs(); th = torch.angle(O.rd(q)) - torch.angle(O.rd(k))
o64 = torch.cat([O.from_patches(torch.fft.irfft2(z, s=(8, 8))) for z in (O.polar(v.abs(), th), O.polar(qkm, torch.angle(v)), O.polar(qkm, th))], 1)
d = (o.double().cpu() - o64).abs(); print("fdsa_patch on exact input: max %.2e rel-L2 %.2e (|o| max %.2e)" % (d.max(), d.norm() / o64.norm(), o64.abs().max()))
o2 = torch.empty_like(o)
ops.fdsa_patch(hid_dw, cx.flat(p + "attn.fft"), o2)
d = (o2.double().cpu() - o64).abs(); print("fdsa_patch on GPU fp32 input: max %.2e rel-L2 %.2e" % (d.max(), d.norm() / o64.norm()))
print("---- cpu fp32 comparisons")
hid32 = O.conv(hid_pre32, sd, p + "attn.to_hidden_dw.", padding=1, groups=152)
d = (hid32.double() - hid64).abs(); print("cpu fp32 hid_dw: max %.2e rel-L2 %.2e" % (d.max(), d.norm() / hid64.norm()))
def spectral(hidt, dt):
    sdd = sd64 if dt == torch.float64 else sd
    q, k, v, vv = hidt.to(dt).chunk(4, 1)
    q, k, v = (torch.fft.rfft2(O.to_patches(t)) for t in (q, k, v))
    v = O.rd(v * sdd[p + "attn.fft"]); qkm = O.rd(q * k).abs(); th = torch.angle(O.rd(q)) - torch.angle(O.rd(k))
    return torch.cat([O.from_patches(torch.fft.irfft2(z, s=(8, 8))) for z in (O.polar(v.abs(), th), O.polar(qkm, torch.angle(v)), O.polar(qkm, th))], 1)
o_cpu32 = spectral(hid32, torch.float32)
d = (o_cpu32.double() - o64).abs(); print("cpu fp32 spectral on cpu fp32 hid_dw vs o64: max %.2e rel-L2 %.2e" % (d.max(), d.norm() / o64.norm()))
o_cpu64_on32 = spectral(hid32, torch.float64)
d = (o_cpu64_on32 - o64).abs(); print("fp64 spectral on cpu fp32 hid_dw vs o64 (pure input-noise amplification): max %.2e rel-L2 %.2e" % (d.max(), d.norm() / o64.norm()))
o_gpu_on_cpu = torch.empty_like(o)
ops.fdsa_patch(hid32.cuda(), cx.flat(p + "attn.fft"), o_gpu_on_cpu)
d = (o_gpu_on_cpu.double().cpu() - o64).abs(); print("gpu fdsa_patch on cpu fp32 hid_dw vs o64: max %.2e rel-L2 %.2e" % (d.max(), d.norm() / o64.norm()))
o_64_on_gpu = spectral(hid_dw.cpu(), torch.float64)
d = (o_64_on_gpu - o64).abs(); print("fp64 spectral on GPU fp32 hid_dw vs o64: max %.2e rel-L2 %.2e" % (d.max(), d.norm() / o64.norm()))
dd = (hid_dw.cpu().double() - hid64); dc = (hid32.double() - hid64)
print("noise stats: gpu max %.2e, cpu max %.2e; correlation of gpu and cpu noise %.3f" % (dd.abs().max(), dc.abs().max(), (dd * dc).sum() / (dd.norm() * dc.norm())))
print("---- locate worst element of GPU-kernel-on-GPU-input")
err = (o2.double().cpu() - o_64_on_gpu).abs()      # same input, fp32 kernel vs fp64 math
print("kernel fp32 vs fp64 math on the same GPU input: max %.2e rel-L2 %.2e" % (err.max(), err.norm() / o_64_on_gpu.norm()))
idx = err.flatten().argmax().item()
bb, cc, yy, xx = [int(v) for v in torch.unravel_index(torch.tensor(idx), err.shape)]
g, ee, py, px = cc // e, cc % e, yy // 8, xx // 8
print("worst at b=%d group=%d e=%d patch=(%d,%d) err=%.3e" % (bb, g, ee, py, px, err.flatten()[idx]))
hd = hid_dw.cpu().double()
pat = lambda ch: hd[bb, ch, py * 8:py * 8 + 8, px * 8:px * 8 + 8]
Q, K, V = (torch.fft.rfft2(pat(r * e + ee)) for r in range(3))
print("|Q| bins:\n", Q.abs()); print("|K| bins:\n", K.abs())
Q32, K32 = (torch.fft.rfft2(pat(r * e + ee).float()) for r in range(2))
print("fp32 torch FFT rel err of Q bins:\n", ((Q32.to(torch.complex128) - Q).abs() / Q.abs()))
print("q patch:\n", pat(ee))
