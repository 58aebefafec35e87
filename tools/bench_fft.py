"""FFT-stage micro-benchmark: every global-FFT kernel shape of the FDN 1120x640 forward, timed alone with CUDA events.  Dev tool, GPU only.

    python tools/bench_fft.py [batch] [tag]          # FDN_B200_LIB=<path> times another build of the library (A/B runs)

Prints one line per (kernel, shape) with the time per launch and the algorithmic GB/s, then the FFT-stage total per image weighted by
the instance counts of one forward (22 FCAFFN, 9 FreBlock, 2 fourier_fuse, 6 prologue transforms) and its fraction
of the measured HBM peak - the same definition bench.py's `fft_stage` uses, without running the rest of the network.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fdn_tip2025_b200 import _lib

if os.environ.get("FDN_B200_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["FDN_B200_LIB"])
from fdn_tip2025_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
TAG = sys.argv[2] if len(sys.argv) > 2 else "fft"
H, W = 640, 1120
dev = "cuda"
PEAK = 6545.6
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


rows = []
stage_ms = 0.0


def rec(name, shape, ms, nbytes, count):
    """count = launches of this shape per forward (per micro-batch of B images)"""
    global stage_ms
    rows.append({"kernel": name, "shape": shape, "ms": ms, "GBps": nbytes / ms / 1e6, "per_forward": count})
    stage_ms += ms * count
    print("%-28s %-22s %8.3f ms %8.1f GB/s  x%d" % (name, shape, ms, nbytes / ms / 1e6, count), flush=True)


def fcaffn(level, c, n):
    h, w = H >> level, W >> level
    wf = w // 2 + 1
    x = torch.randn(B, c, h, w, device=dev)
    y = torch.empty_like(x)
    spec = torch.empty(B, c, h, wf, 2, device=dev)
    amp = torch.rand(B, 3, h, wf, device=dev) * 3
    pha = (torch.rand(B, 3, h, wf, device=dev) - 0.5) * 6
    wxa = torch.randn(c * 3, device=dev)
    wxp = torch.randn(c * 3, device=dev)
    ops.fft_prepare(h, w)
    sh = "L%d %dx%dx%d" % (level + 1, c, h, w)
    rb, sb = x.numel() * 4, spec.numel() * 4
    rec("rows_r2c", sh, timeit(lambda: ops.fft_rows_r2c(x, spec)), rb + sb, n)
    rec("cols_fwd_mod_inv", sh, timeit(lambda: ops.fft_cols(spec, h * wf, wf, spec, h * wf, wf, B * c, h, wf, w, ops.COLS_FWD_MOD_INV,
                                                             c, amp, pha, wxa, wxp)), 2 * sb + 2 * amp.numel() * 4, n)
    rec("rows_c2r", sh, timeit(lambda: ops.fft_rows_c2r(spec, y, 1.0 / (h * w))), rb + sb, n)


def freblock(level, nc, n):
    h, w = H >> level, W >> level
    wf = w // 2 + 1
    x = torch.randn(B, nc, h, w, device=dev)
    y = torch.empty_like(x)
    spec = torch.empty(B, nc, h, wf, 2, device=dev)
    sh = "MAR L%d %dx%dx%d" % (level + 1, nc, h, w)
    rb, sb = x.numel() * 4, spec.numel() * 4
    rec("rows_r2c", sh, timeit(lambda: ops.fft_rows_r2c(x, spec)), rb + sb, n)
    rec("cols_fwd", sh, timeit(lambda: ops.fft_cols(spec, h * wf, wf, spec, h * wf, wf, B * nc, h, wf, w, ops.COLS_FWD)), 2 * sb, n)
    rec("cols_inv", sh, timeit(lambda: ops.fft_cols(spec, h * wf, wf, spec, h * wf, wf, B * nc, h, wf, w, ops.COLS_INV)), 2 * sb, n)
    rec("rows_c2r+res", sh, timeit(lambda: ops.fft_rows_c2r(spec, y, 1.0 / (h * w), res=x, res_coef=2.0, planes_per_image=nc)), 2 * rb + sb, n)


def fourier_fuse(level, nc):
    h, w = H >> level, W >> level
    hp, wp = h + 2, w + 2
    wpf, wf = wp // 2 + 1, w // 2 + 1
    x = torch.randn(B, nc, hp, wp, device=dev)
    spec = torch.empty(B, nc, hp, wpf, 2, device=dev)
    spec2 = torch.empty(B, nc, h, wf, 2, device=dev)
    y = torch.empty(B, nc, h, w, device=dev)
    ops.fft_prepare(hp, wp)
    sh = "fuse L%d %dx%dx%d" % (level + 1, nc, hp, wp)
    rb, sb = x.numel() * 4, spec.numel() * 4
    rec("rows_r2c", sh, timeit(lambda: ops.fft_rows_r2c(x, spec)), rb + sb, 1)
    rec("cols_fwd", sh, timeit(lambda: ops.fft_cols(spec, hp * wpf, wpf, spec, hp * wpf, wpf, B * nc, hp, wpf, wp, ops.COLS_FWD)), 2 * sb, 1)
    rec("cols_inv(slice)", sh, timeit(lambda: ops.fft_cols(spec, hp * wpf, wpf, spec2, h * wf, wf, B * nc, h, wf, w, ops.COLS_INV)),
        2 * spec2.numel() * 4, 1)
    rec("rows_c2r", sh, timeit(lambda: ops.fft_rows_c2r(spec2, y, 1.0 / (h * w))), y.numel() * 4 + spec2.numel() * 4, 1)


def prologue(level):
    h, w = H >> level, W >> level
    wf = w // 2 + 1
    x = torch.randn(B, 3, h, w, device=dev)
    spec = torch.empty(B, 3, h, wf, 2, device=dev)
    out = torch.empty(B, 3, h, wf, device=dev)
    sh = "prologue L%d 3x%dx%d" % (level + 1, h, w)
    rec("rows_r2c", sh, timeit(lambda: ops.fft_rows_r2c(x, spec)), x.numel() * 4 + spec.numel() * 4, 2)
    rec("cols_angle", sh, timeit(lambda: ops.fft_cols(spec, h * wf, wf, out, h * wf, wf, B * 3, h, wf, w, ops.COLS_FWD_ANGLE)),
        spec.numel() * 4 + out.numel() * 4, 1)
    rec("cols_abs", sh, timeit(lambda: ops.fft_cols(spec, h * wf, wf, out, h * wf, wf, B * 3, h, wf, w, ops.COLS_FWD_ABS)),
        spec.numel() * 4 + out.numel() * 4, 1)


print("lib:", _lib.LIB_PATH, " FDN_FFT_V =", os.environ.get("FDN_FFT_V"), " batch", B)
fcaffn(0, 32, 6)
fcaffn(1, 64, 6)
fcaffn(2, 128, 10)
freblock(0, 12, 3)
freblock(1, 24, 3)
freblock(2, 48, 3)
fourier_fuse(0, 12)
fourier_fuse(1, 24)
for lv in range(3):
    prologue(lv)
alg = 2.679e9 * B
print("FFT stage (without spec_mlp): %.3f ms per %d images -> %.1f GB/s algorithmic = %.2f %% of %.0f GB/s" % (
    stage_ms, B, alg / stage_ms / 1e6, 100 * alg / stage_ms / 1e6 / PEAK, PEAK))
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"lib": _lib.LIB_PATH, "variant": os.environ.get("FDN_FFT_V"), "batch": B, "stage_ms": stage_ms,
           "stage_GBps": alg / stage_ms / 1e6, "frac": alg / stage_ms / 1e6 / PEAK, "rows": rows},
          open("gpurun_out/bench_fft_%s.json" % TAG, "w"), indent=1)
