"""Run one encoder TransformerBlock (FDSA + FDFFN + FCAFFN) of a given level on B images; used under ncu.  Dev tool, GPU only.

    python tools/one_block.py [level 1|2|3] [batch] [reps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fdn_tip2025_b200 import archs, synth

lvl = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
H, W = 640 >> (lvl - 1), 1120 >> (lvl - 1)
C = 32 << (lvl - 1)
net = archs.FDN()
net.load_state_dict(synth.fdn_state_dict(dim=32, seed=0, damp=0.03), strict=True)
net = net.cuda().eval()
cx = net._context()
g = torch.Generator().manual_seed(5)
x = torch.randn(B, C, H, W, generator=g).cuda()
wf = W // 2 + 1
side = (torch.rand(B, 3, H, wf, generator=g).cuda() * 50, (torch.rand(B, 3, H, wf, generator=g).cuda() - 0.5) * 6,
        torch.rand(B, 3, H, W, generator=g).cuda())
with torch.no_grad():
    for _ in range(reps):
        y = archs._tblock(cx, x, side, "net_p.encoder_level%d.0." % lvl)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
