"""Per-role barrier wait cycles of the tcgen05 1x1-conv kernel for the main layer shapes (fdn_pw_mma_set_debug).  Dev tool.
Needs the profiling build:  FDN_MMA_PROFILE=1 python -m fdn_tip2025_b200.build --force   (rebuild without it afterwards)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fdn_tip2025_b200 import ops, packing, _lib
B = 4
dev = "cuda"
shapes = [("L1 to_hidden", 716800, 32, 152, 1), ("L1 fdsa_out", 716800, 114, 32, 2), ("L1 ffn_in", 716800, 32, 86, 1), ("L1 ffn_out", 716800, 86, 32, 0),
          ("L2 to_hidden", 179200, 64, 304, 1), ("L2 fdsa_out", 179200, 228, 64, 2), ("L2 ffn_out", 179200, 172, 64, 0),
          ("L3 to_hidden", 44800, 128, 612, 1), ("L3 fdsa_out", 44800, 459, 128, 2), ("L3 ffn_in", 44800, 128, 345, 1), ("L3 ffn_out", 44800, 345, 128, 0)]
dbg = torch.zeros(8, dtype=torch.int64, device=dev)
for name, hw, k, n, pro in shapes:
    h, w = 640, hw // 640
    x = torch.randn(B, k, h, w, device=dev); wgt = torch.randn(n, k, device=dev) / k ** 0.5
    packed = packing.pack_weight(wgt, grouped_e=(k // 3 if pro == 2 else None))
    out = torch.empty(B, n, h, w, device=dev); res = torch.randn(B, n, h, w, device=dev)
    kw = {}
    if pro == 1: kw = dict(prologue=1, ln=(torch.ones(k, device=dev), torch.zeros(k, device=dev)))
    elif pro == 2:
        e = k // 3
        kw = dict(prologue=2, ln=(torch.ones(3, e, device=dev), torch.zeros(3, e, device=dev)), aux=torch.randn(B, e, h, w, device=dev), aux_bs=e * h * w,
                  stats=torch.ones(B, 3, 2, h * w, device=dev))
    run = lambda: ops.pw_mma([x], packed, out, res=res if pro != 1 else None, **kw)
    run(); torch.cuda.synchronize()
    dbg.zero_(); _lib.call("fdn_pw_mma_set_debug", dbg.data_ptr())
    run(); torch.cuda.synchronize()
    _lib.call("fdn_pw_mma_set_debug", None)
    d = dbg.cpu().tolist(); tot = max(d[6], 1)
    print("%-13s total %.0f kcyc/CTA | producer waits raw_full %4.1f%% a_empty %4.1f%% | epilogue waits acc_full %4.1f%% | MMA waits a_full %4.1f%% acc_empty %4.1f%% | loader waits raw_empty %4.1f%%"
          % (name, tot / 148 / 1e3, 100 * d[0] / tot, 100 * d[1] / tot, 100 * d[2] / tot, 100 * d[3] / tot, 100 * d[4] / tot, 100 * d[5] / tot))
