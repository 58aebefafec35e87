"""Per (kernel, operand-bytes) timing of one FDN forward: CUDA events around every C-ABI launch.  Dev tool, GPU only.

    python tools/kernel_breakdown.py [batch] [H] [W]

Rows are grouped by entry point and algorithmic bytes of the launch (which identifies the level/shape), so the
table shows where the forward's time goes per level and at what fraction of the HBM peak each shape runs.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fdn_tip2025_b200 import _lib, archs, ops, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H = int(sys.argv[2]) if len(sys.argv) > 2 else 640
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1120
net = archs.FDN()
net.load_state_dict(synth.fdn_state_dict(dim=32, seed=0, damp=0.03), strict=True)
net = net.cuda().eval()
x = synth.low_light_images(B, H, W).cuda()
ratio = torch.full((B, 1), 0.35, device="cuda")
for _ in range(2):
    net(x, ratio_i=ratio)
torch.cuda.synchronize()
recs = []


def hook(name, fn, cargs):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn()
    e1.record()
    recs.append((name, e0, e1, ops.pending_bytes))
    ops.pending_bytes = 0
    return rc


ops.count_bytes = True
ops.pending_bytes = 0
_lib.profile_hook = hook
net(x, ratio_i=ratio)
torch.cuda.synchronize()
_lib.profile_hook = None
agg = {}
for name, e0, e1, nb in recs:
    a = agg.setdefault((name, nb), [0.0, 0])
    a[0] += e0.elapsed_time(e1)
    a[1] += 1
tot = sum(a[0] for a in agg.values())
print("FDN %dx%d batch %d: %.2f ms in kernels (%.2f ms / image), %d launches" % (W, H, B, tot, tot / B, len(recs)))
print("%-22s %10s %5s %9s %7s %8s" % ("kernel", "MB/launch", "n", "ms total", "share", "GB/s"))
rows = []
for (name, nb), (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    gbs = nb * n / (ms * 1e-3) / 1e9 if ms > 0 else 0
    rows.append({"kernel": name, "mb": nb / 1e6, "n": n, "ms": ms, "share": ms / tot, "gbs": gbs})
    if ms / tot >= 0.002:
        print("%-22s %10.1f %5d %9.3f %7.4f %8.1f" % (name.replace("fdn_", ""), nb / 1e6, n, ms, ms / tot, gbs))
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/kernel_breakdown.json", "w") as f:
    json.dump({"batch": B, "h": H, "w": W, "total_ms": tot, "rows": rows}, f)
