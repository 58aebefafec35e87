"""FDN inference throughput on B200 (BASELINE.json metric: FDN images/s at 1120x640 on 1/2/4/8 GPUs).

    python bench.py --gpus 1 --steps K --warmup W                       # this repo's CUDA path
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                                # CPU reference arm (oracle port, host cores)

A step = one FDN forward over the per-GPU batch (default 8 images of 1120x640, i.e. BASELINE config 3's batch 64
sharded over 8 GPUs; weak scaling: the per-GPU batch is fixed as N grows).  Ranks shard by image, there is no
collective on the data path; timing is CUDA events on the launching stream, max over ranks.
Prints ONE JSON line (see the keys at the bottom).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

B_ALG_PER_PIXEL = 19985.0          # ideal-fused HBM bytes per pixel of one FDN(dim 32) forward
FFT_STAGE_BYTES_1120x640 = 2.679e9  # global-FFT stages, per image


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(h_full, w_full, steps, warmup, sample_hw=(256, 320)):
    """Oracle port of the reference forward on the host cores; returns 1120x640-equivalent images/s and details."""
    from fdn_tip2025_b200 import synth
    from oracle import fdn_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.fdn_state_dict(dim=32, seed=0, damp=0.03)
    sh, sw = sample_hw
    x = synth.low_light_images(1, sh, sw)
    ratio = torch.full((1, 1), 0.35)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.fdn(x, ratio, sd, "lolblur")
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    t = sum(times) / len(times)
    scale = (sh * sw) / float(h_full * w_full)
    return scale / t, t, cores, "1 image of %dx%d per step (%.1f%% of a %dx%d frame), per-pixel scaled; fp32 torch CPU, %d threads" % (
        sw, sh, 100 * scale, w_full, h_full, torch.get_num_threads())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, t, cores, sample = cpu_reference_rate(args.height, args.width, max(1, args.steps), max(0, min(args.warmup, 1)))
    line = {
        "impl": "reference", "metric": "FDN images/sec at 1120x640", "value": val, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "FDN (LOL-Blur, dim 32) forward, %dx%d, CPU reference arm" % (args.width, args.height)},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="images per GPU per step")
    ap.add_argument("--height", type=int, default=640)
    ap.add_argument("--width", type=int, default=1120)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from fdn_tip2025_b200 import _lib, archs, sharding, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    warmup = max(3, args.warmup)
    steps = max(1, args.steps)
    H, W, B = args.height, args.width, args.batch

    net = archs.FDN()
    net.load_state_dict(synth.fdn_state_dict(dim=32, seed=0, damp=0.03), strict=True)
    net = net.to(dev).eval()
    lp = archs.I_predict_net()
    lp.load_state_dict(synth.lpnet_state_dict(seed=3), strict=True)
    lp = lp.to(dev).eval()

    # image i of the global batch goes to rank i mod world (reference validation rule, image_restoration_model.py:731)
    idx = sharding.shard_indices(rank, world, B)
    host = torch.cat([synth.low_light_images(1, H, W, first_index=i) for i in idx], 0).pin_memory()
    x = host.to(dev, non_blocking=True)
    ratio = lp(x)                                   # LPNet is outside the timed region: the metric is the FDN forward
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        t = sharding.max_over_ranks(torch.tensor([ms], device=dev))
        barrier()
        return t.item()

    def step_resident():
        net(x, ratio_i=ratio)

    for _ in range(warmup):
        step_resident()
    sampler = ClockSampler(local)
    sampler.start()
    n0 = _lib.launch_count
    ms_total = timed(step_resident, steps)
    launches = _lib.launch_count - n0
    clocks = sampler.stop()
    ms_step = ms_total / steps
    value = world * B / (ms_step * 1e-3)

    # ---- end to end through the public API with host buffers (pinned H2D of the step's images, D2H of the result)
    out_host = torch.empty(B, 3, H, W).pin_memory()

    def step_e2e():
        xd = host.to(dev, non_blocking=True)
        out = net(xd, ratio_i=ratio)[0]
        out_host.copy_(out, non_blocking=True)

    step_e2e()
    e2e_ms = timed(step_e2e, steps) / steps
    e2e = {"value": world * B / (e2e_ms * 1e-3), "unit": "images/s", "h2d_bytes_per_step": host.numel() * 4,
           "d2h_bytes_per_step": out_host.numel() * 4}

    # ---- per-kernel timing of one step (CUDA events around every launch, outside the timed region) -> roofline
    roofline, families, fft_stage = None, None, None
    peak, peak_src = measured_peaks()
    if rank == 0 and not args.no_profile:
        from fdn_tip2025_b200 import ops
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
        step_resident()
        torch.cuda.synchronize()
        _lib.profile_hook = None
        ops.count_bytes = False
        fam = {}
        for name, e0, e1, nb in recs:
            f = fam.setdefault(name, [0.0, 0, 0])
            f[0] += e0.elapsed_time(e1)
            f[1] += 1
            f[2] += nb
        tot = sum(f[0] for f in fam.values())
        families = {k: {"ms": round(v[0], 3), "launches": v[1], "share": round(v[0] / tot, 4),
                        "alg_GBps": round(v[2] / (v[0] * 1e-3) / 1e9, 1) if v[0] > 0 else None}
                    for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0])}
        # FFT stage (BASELINE metric): every global rfft2 -> spectral op -> irfft2 group; its algorithmic
        # bytes per image are fixed by the architecture, its time is the sum of the kernels that implement it
        fft_ms = sum(v[0] for k, v in fam.items() if k in ("fdn_fft_rows_r2c", "fdn_fft_cols", "fdn_fft_rows_c2r", "fdn_spec_mlp"))
        if fft_ms > 0 and (H, W) == (640, 1120):
            fft_gbps = FFT_STAGE_BYTES_1120x640 * B / (fft_ms * 1e-3) / 1e9
            fft_stage = {"alg_bytes_per_image": FFT_STAGE_BYTES_1120x640, "ms_per_step": round(fft_ms, 3), "achieved_GBps": round(fft_gbps, 1),
                         "frac_of_measured_peak": round(fft_gbps / peak, 4), "frac_of_nominal_8TBps": round(fft_gbps / 8000.0, 4),
                         "share_of_step": round(fft_ms / tot, 4), "kernels": "fdn_fft_rows_r2c + fdn_fft_cols + fdn_fft_rows_c2r + fdn_spec_mlp"}
        top = max(fam.items(), key=lambda kv: kv[1][0])
        ach = top[1][2] / (top[1][0] * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tr_path = os.path.join(ROOT, "profiles", "r1_traffic_ratio.json")
        if os.path.exists(tr_path):
            with open(tr_path) as f:
                tr = json.load(f)
            if top[0] in tr:        # DRAM bytes / algorithmic bytes measured with ncu --set full for this kernel family
                traffic = tr[top[0]]["ratio"] * top[1][2] / top[1][1]
                traffic_src = "profiles/r1_traffic_ratio.json: ncu dram bytes / algorithmic bytes = %.2f" % tr[top[0]]["ratio"]
        roofline = {"kernel": top[0], "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "share_of_step": top[1][0] / tot,
                    "bytes_per_launch": top[1][2] / top[1][1], "avg_launch_ms": top[1][0] / top[1][1]}

    fwd_alg = B_ALG_PER_PIXEL * H * W * (value) / 1e9 / world       # GB/s per GPU if the forward were ideally fused

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, t, cores, sample = cpu_reference_rate(H, W, 1, 1)
        cpu_baseline = {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": "FDN images/sec at 1120x640", "value": value, "unit": "images/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "FDN (LOL-Blur, dim 32) forward, %dx%d, %d images per GPU per step (BASELINE config 3: batch 64 "
                                   "sharded by image)" % (W, H, B),
                       "weights": "synthetic, seed 0, net_p project_out x0.03", "precision": "fp32 I/O; 1x1 convs on tcgen05 in 3xTF32 (FDN_B200_GEMM=%s), everything else fp32 FFMA" % os.environ.get("FDN_B200_GEMM", "tf32x3"),
                       "l2": "per-step working set (>1 GB of activations) exceeds the 126 MB L2; no explicit flush",
                       "micro_batch": archs._micro_batch(B, H, W), "ratio_i": "I_predict_net output, computed outside the timed region"},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "forward_roofline": {"alg_bytes_per_image": B_ALG_PER_PIXEL * H * W, "achieved_GBps_per_gpu": fwd_alg,
                                 "frac_of_measured_peak": fwd_alg / peak},
            "fft_stage": fft_stage, "kernel_families": families, "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
