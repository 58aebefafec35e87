"""Where do two GEMM modes (default tcgen05 3xTF32 vs FFMA) start to differ inside one FDN forward?  Dev tool, GPU only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fdn_tip2025_b200 import archs, synth

h, w, b = int(os.environ.get("H", 128)), int(os.environ.get("W", 160)), int(os.environ.get("B", 2))
sd = synth.fdn_state_dict(dim=32, seed=7, damp=0.03)
net = archs.FDN(); net.load_state_dict(sd, strict=True); net = net.cuda()
x = synth.low_light_images(b, h, w).cuda(); ratio = torch.full((b, 1), 0.35).cuda()
recs = {}
orig = {n: getattr(archs, n) for n in ("_fdsa", "_fdffn", "_fcaffn")}
def wrap(name):
    f = orig[name]
    def g(cx, x, *a):
        out = f(cx, x, *a)
        recs[mode].append((name + ":" + a[-1], out.detach().clone()))
        return out
    return g
for n in orig: setattr(archs, n, wrap(n))
for mode in ("ffma", "tf32x3"):
    os.environ["FDN_B200_GEMM"] = mode
    recs[mode] = []
    out = net(x, ratio_i=ratio)
    torch.cuda.synchronize()
    recs[mode].append(("final", out[0].clone()))
for (n, a), (_, c) in zip(recs["ffma"], recs["tf32x3"]):
    d = (a - c).abs().max().item(); s = a.abs().max().item()
    print("%-48s maxdiff %.2e rel %.1e" % (n, d, d / s))
