"""Per-shape timing of the 1x1-convolution kernels (tcgen05 vs FFMA) at the BASELINE sizes.  Dev tool, GPU only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fdn_tip2025_b200 import ops, packing

B = int(os.environ.get("B", "4"))
only = os.environ.get("ONLY")
shapes = [("L1 to_hidden", 716800, 32, 152, 1), ("L1 fdsa_out", 716800, 114, 32, 2), ("L1 ffn_in", 716800, 32, 86, 1),
          ("L1 ffn_out", 716800, 86, 32, 0), ("L1 fca_in", 716800, 32, 32, 3), ("L1 fca_out", 716800, 32, 32, 0),
          ("L2 to_hidden", 179200, 64, 304, 1), ("L2 fdsa_out", 179200, 228, 64, 2), ("L2 ffn_in", 179200, 64, 172, 1),
          ("L2 ffn_out", 179200, 172, 64, 0), ("L3 to_hidden", 44800, 128, 612, 1), ("L3 fdsa_out", 44800, 459, 128, 2),
          ("L3 ffn_in", 44800, 128, 345, 1), ("L3 ffn_out", 44800, 345, 128, 0)]
dev = "cuda"
for name, hw, k, n, pro in shapes:
    if only and only not in name:
        continue
    h, w = hw // 1120 * 2 if False else 640, hw // 640
    x = torch.randn(B, k, h, w, device=dev)
    wgt = torch.randn(n, k, device=dev) / k ** 0.5
    packed = packing.pack_weight(wgt, grouped_e=(k // 3 if pro == 2 else None))
    out = torch.empty(B, n, h, w, device=dev)
    res = torch.randn(B, n, h, w, device=dev)
    kw = {}
    if pro == 1:
        kw = dict(prologue=1, ln=(torch.ones(k, device=dev), torch.zeros(k, device=dev)))
    elif pro == 2:
        e = k // 3
        hid = torch.randn(B, 4 * e, h, w, device=dev)
        stats = torch.zeros(B, 3, 2, h * w, device=dev) + 1
        kw = dict(prologue=2, ln=(torch.ones(3, e, device=dev), torch.zeros(3, e, device=dev)), aux=hid.view(-1)[3 * e * h * w:], aux_bs=4 * e * h * w, stats=stats)
    elif pro == 3:
        kw = dict(prologue=3, ln=(torch.ones(k, device=dev), torch.zeros(k, device=dev)), aux=torch.randn(B, k, h, w, device=dev), aux_bs=k * h * w)
    def run_mma(passes):
        ops.pw_mma([x], packed, out, res=res if pro != 1 else None, passes=passes, **kw)
    def run_ffma():
        ops.pw_conv([(x, 0)], wgt.t().contiguous(), out, ln=kw.get("ln") if pro == 1 else None, res=res if pro != 1 else None)
    def t(fn, n=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    byt = (k + n + (n if pro != 1 else 0) + (k // 3 if pro == 2 else 0) + (k if pro == 3 else 0)) * hw * B * 4
    t3, t1 = t(lambda: run_mma(3)), t(lambda: run_mma(1))
    tf = t(run_ffma) if pro in (0, 1) else float("nan")
    print("%-14s K=%3d N=%3d pro=%d  3xTF32 %.3f ms (%.0f GB/s)  TF32 %.3f ms  FFMA %.3f ms   ideal@6.5TB/s %.3f ms"
          % (name, k, n, pro, t3, byt / t3 / 1e6, t1, tf, byt / 6.5e9), flush=True)
