"""FDN inference throughput on B200 (BASELINE.json metric: FDN images/s at 1120x640 on 1/2/4/8 GPUs).

    python bench.py --gpus 1 --steps K --warmup W                       # this repo's CUDA path
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                                # the reference's own CPU forward on the host cores
    python bench.py --config lolv1_600x400_b8 | lpnet_1120x640 | fdn_4k | lolblur_1120x640_b64    # the other BASELINE configs

A step = one forward over the per-GPU batch (default: 8 images of 1120x640 through FDN, i.e. BASELINE config 3's batch 64
sharded over 8 GPUs; weak scaling: the per-GPU batch is fixed as N grows).  Ranks shard by image, there is no
collective on the data path; timing is CUDA events on the launching stream, max over ranks.
Prints ONE JSON line (see the keys at the bottom).
"""
import argparse
import contextlib
import io
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

# ideal-fused HBM bytes per pixel of one forward, and the global-FFT stage bytes per image
CONFIGS = {
    # name: (module, H, W, images per GPU per step, B_alg bytes per pixel, description)
    "lolblur_1120x640": ("FDN", 640, 1120, 8, 19985.0, "BASELINE config 3: FDN (LOL-Blur, dim 32) 1120x640, batch 64 sharded by image over 8 GPUs"),
    "lolblur_1120x640_b64": ("FDN", 640, 1120, 64, 19985.0, "BASELINE config 3 on one GPU: FDN 1120x640, the whole batch of 64"),
    "lolv1_600x400_b8": ("FDN_lolv1", 416, 608, 8, 15442.0, "BASELINE config 2: FDN_lolv1 (dim 24), 600x400 reflect-padded to 608x416, batch 8"),
    "lpnet_1120x640": ("I_predict_net", 640, 1120, 8, 70e6 / (640 * 1120), "BASELINE config 4: I_predict_net (LPNet) 1120x640, 8 images per GPU"),
    "fdn_4k": ("FDN", 2176, 3840, 1, 19985.0, "BASELINE config 5: FDN 3840x2160 padded to 3840x2176, one image per GPU, no tiling"),
    "fdn_256": ("FDN", 256, 256, 1, 19985.0, "BASELINE config 1 shape: FDN 256x256, one image"),
}
FFT_STAGE_BYTES_1120x640 = 2.679e9
BENCH_DAMP = 0.005      # net_p project_out scale of the synthetic weights: well conditioned, so the parity key is meaningful (timing is weight independent)


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


def _synthetic_weights(kind):
    from fdn_tip2025_b200 import synth
    if kind == "I_predict_net":
        return synth.lpnet_state_dict(seed=3)
    return synth.fdn_state_dict(dim=32 if kind == "FDN" else 24, seed=0, damp=BENCH_DAMP)


def _checkpoint_weights(kind):
    """Real checkpoints are used unchanged when the driver supplies them: checkpoint/FDN_lolblur.pth,
    FDN_lolv1.pth, LPNet_lolblur.pth next to bench.py or under FDN_CHECKPOINT_DIR."""
    name = {"FDN": "FDN_lolblur.pth", "FDN_lolv1": "FDN_lolv1.pth", "I_predict_net": "LPNet_lolblur.pth"}[kind]
    for d in (os.environ.get("FDN_CHECKPOINT_DIR"), os.path.join(ROOT, "checkpoint")):
        if d and os.path.isfile(os.path.join(d, name)):
            return torch.load(os.path.join(d, name), map_location="cpu")["params"], os.path.join(d, name)
    return None, None


def bench_weights(kind):
    sd, path = _checkpoint_weights(kind)
    if sd is not None:
        return sd, "checkpoint %s" % path
    return _synthetic_weights(kind), ("synthetic, seed 3" if kind == "I_predict_net" else "synthetic, seed 0, net_p project_out x%g" % BENCH_DAMP)


def cpu_reference_rate(kind, h_full, w_full, sample_hw, repeats=1, warmup=0):
    """The reference's own forward (oracle/_ref, staged by oracle/stage_ref.py) on the host cores - or the oracle port when the
    reference files are not staged.  Returns (full-frame-equivalent images/s, seconds per sample, cores, kind, sample text)."""
    from fdn_tip2025_b200 import synth
    from oracle import fdn_oracle as O
    from oracle import ref_loader as R
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = _synthetic_weights(kind)
    sh, sw = sample_hw
    x = synth.low_light_images(1, sh, sw)
    ratio = torch.full((1, 1), 0.35)
    use_ref = R.available()
    if use_ref:
        net = R.build(kind)
        net.load_state_dict(sd, strict=True)
        fwd = (lambda: R.run(net, x)) if kind == "I_predict_net" else (lambda: R.run(net, x, ratio_i=ratio))
    elif kind == "I_predict_net":
        fwd = lambda: O.lpnet(x, sd)
    else:
        fwd = lambda: O.fdn(x, ratio, sd, "lolblur" if kind == "FDN" else "lolv1")
    times = []
    with torch.no_grad():
        for i in range(warmup + repeats):
            t0 = time.perf_counter()
            fwd()
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    t = sum(times) / len(times)
    scale = (sh * sw) / float(h_full * w_full)
    what = "the reference's %s.forward (oracle/_ref, unmodified arch file)" % kind if use_ref else "oracle port of %s.forward" % kind
    if scale == 1.0:
        sample = "%s on 1 full %dx%d frame per step, %d timed, fp32 torch CPU (%s), %d threads" % (what, w_full, h_full, repeats, torch.__version__, torch.get_num_threads())
    else:
        sample = "%s on 1 frame of %dx%d (%.1f%% of a %dx%d frame), per-pixel scaled, fp32 torch CPU, %d threads" % (
            what, sw, sh, 100 * scale, w_full, h_full, torch.get_num_threads())
    return scale / t, t, cores, ("reference" if use_ref else "port"), sample


def metric_name(config_name, kind, H, W):
    return "FDN images/sec at 1120x640" if config_name.startswith("lolblur") else "%s images/sec at %dx%d" % (kind, W, H)


def run_reference(args, cfg):
    """Reference arm: the reference's own CPU implementation on this box's host cores, SAME config (a full frame of the configured size;
    one timed forward - a 1120x640 frame takes minutes on CPU - unless FDN_REF_SAMPLE=HxW asks for a per-pixel-scaled sample)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, H, W, B, _, desc = cfg
    sample_hw = (H, W)
    if os.environ.get("FDN_REF_SAMPLE"):
        sample_hw = tuple(int(v) for v in os.environ["FDN_REF_SAMPLE"].lower().split("x"))
    elif kind != "I_predict_net" and H * W > 1200 * 700:
        sample_hw = (640, 1120)          # 4K: a 1120x640 frame, scaled per pixel (a 4K CPU forward takes the better part of an hour)
    repeats = 1 if kind != "I_predict_net" else max(1, min(args.steps, 5))
    val, t, cores, how, sample = cpu_reference_rate(kind, H, W, sample_hw, repeats=repeats, warmup=0 if kind != "I_predict_net" else 1)
    line = {
        "impl": "reference", "metric": metric_name(args.config, kind, H, W), "value": val, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc + " - CPU reference arm", "same_config": sample_hw == (H, W), "timed_forwards": repeats},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": how, "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def gpu_eager_reference(kind, H, W, dev):
    """The like-for-like GPU incumbent: the reference module itself, .cuda(), PyTorch eager (cuDNN / cuFFT /
    ATen), batch 1.  Reported next to the result, never part of the timed region."""
    from fdn_tip2025_b200 import synth
    from oracle import ref_loader as R
    if not R.available():
        return None
    try:
        net = R.build(kind)
        net.load_state_dict(_synthetic_weights(kind), strict=True)
        net = net.to(dev)
        x = synth.low_light_images(1, H, W).to(dev)
        ratio = torch.full((1, 1), 0.35, device=dev)
        fwd = (lambda: R.run(net, x)) if kind == "I_predict_net" else (lambda: R.run(net, x, ratio_i=ratio, device=dev))
        fwd()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            fwd()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        del net
        torch.cuda.empty_cache()
        return {"value": 1e3 / ms, "unit": "images/s", "ms_per_image": ms, "batch": 1,
                "what": "reference %s module (oracle/_ref) on cuda:0, PyTorch %s eager, cudnn.allow_tf32=%s, matmul.allow_tf32=%s" % (
                    kind, torch.__version__, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)}
    except Exception as exc:        # e.g. out of memory at 4K: report, do not fail the bench
        torch.cuda.empty_cache()
        return {"unavailable": "%s: %s" % (type(exc).__name__, str(exc)[:200])}


def output_parity(kind, net, x_dev, ratio_dev, sd):
    """One image of the timed batch against the fp64 oracle on the host, outside the timed region (north-star gate: max-abs <= 1e-3,
    PSNR >= 50 dB)."""
    from oracle import fdn_oracle as O
    t0 = time.perf_counter()
    x1, r1 = x_dev[:1], ratio_dev[:1]
    sd64 = O.to_dtype(sd, torch.float64)
    if kind == "I_predict_net":
        got = net(x1).double().cpu()
        ref = O.lpnet(x1.double().cpu(), sd64)
    else:
        got = net(x1, ratio_i=r1)[0].double().cpu()
        ref = O.fdn(x1.double().cpu(), r1.double().cpu(), sd64, "lolblur" if kind == "FDN" else "lolv1")[0]
    d = (got - ref).abs()
    return {"max_abs": d.max().item(), "psnr_db": O.psnr(got, ref) if kind != "I_predict_net" else None, "frac_gt_1e-3": (d > 1e-3).double().mean().item(),
            "against": "oracle/fdn_oracle.py in float64 on the host (pinned to the reference: oracle/VALIDATION.txt), image 0 of the timed batch",
            "gate": "max_abs <= 1e-3 and psnr >= 50 dB", "pass": bool(d.max().item() <= 1e-3), "oracle_seconds": round(time.perf_counter() - t0, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="lolblur_1120x640", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=None, help="images per GPU per step (default: the config's)")
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-eager", action="store_true")
    args = ap.parse_args()
    kind, H, W, B, balg_px, desc = CONFIGS[args.config]
    H, W, B = args.height or H, args.width or W, args.batch or B
    cfg = (kind, H, W, B, balg_px, desc)
    if args.impl == "reference":
        return run_reference(args, cfg)

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

    sd, weights_desc = bench_weights(kind)
    net = getattr(archs, kind)()
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval()
    lp = archs.I_predict_net()
    lp.load_state_dict(bench_weights("I_predict_net")[0], strict=True)
    lp = lp.to(dev).eval()

    # image i of the global batch goes to rank i mod world (reference validation rule, image_restoration_model.py:731)
    idx = sharding.shard_indices(rank, world, B)
    host = torch.cat([synth.low_light_images(1, H, W, first_index=i) for i in idx], 0).pin_memory()
    x = host.to(dev, non_blocking=True)
    is_lp = kind == "I_predict_net"

    def ratio_of(xd):
        """ratio_i as the inference scripts form it (inference_fdn_lolblur.py:65,71; inference_fdn_lolv1.py:58-64)."""
        r = lp(xd)
        if kind == "FDN_lolv1":
            gray = (0.2989 * xd[:, 0] + 0.587 * xd[:, 1] + 0.114 * xd[:, 2]).mean(dim=(1, 2)).view(-1, 1)
            r = gray / r
        return r

    ratio = ratio_of(x)
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
        return net(x) if is_lp else net(x, ratio_i=ratio)

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

    # ---- end to end through the public API with host buffers: pinned H2D of the step's frames, LPNet -> ratio_i -> FDN exactly as
    # the inference scripts call them, D2H of the restored frames
    out_host = (torch.empty(B, 1) if is_lp else torch.empty(B, 3, H, W)).pin_memory()

    def step_e2e():
        xd = host.to(dev, non_blocking=True)
        out = net(xd) if is_lp else net(xd, ratio_i=ratio_of(xd))[0]
        out_host.copy_(out, non_blocking=True)

    step_e2e()
    e2e_ms = timed(step_e2e, steps) / steps
    e2e = {"value": world * B / (e2e_ms * 1e-3), "unit": "images/s", "h2d_bytes_per_step": host.numel() * 4,
           "d2h_bytes_per_step": out_host.numel() * 4,
           "api": "I_predict_net(x)" if is_lp else "ratio = I_predict_net(x); %s(x, ratio_i=ratio) with pinned host input and output" % kind}

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
        fft_names = ("fdn_fft_rows_r2c", "fdn_fft_cols", "fdn_fft_rows_c2r", "fdn_spec_mlp", "fdn_fft_plane")
        fft_ms = sum(v[0] for k, v in fam.items() if k in fft_names)
        if fft_ms > 0 and (H, W) == (640, 1120) and kind == "FDN":
            fft_gbps = FFT_STAGE_BYTES_1120x640 * B / (fft_ms * 1e-3) / 1e9
            fft_stage = {"alg_bytes_per_image": FFT_STAGE_BYTES_1120x640, "ms_per_step": round(fft_ms, 3), "achieved_GBps": round(fft_gbps, 1),
                         "frac_of_measured_peak": round(fft_gbps / peak, 4), "frac_of_nominal_8TBps": round(fft_gbps / 8000.0, 4),
                         "share_of_step": round(fft_ms / tot, 4), "kernels": " + ".join(k for k in fft_names if k in fam)}
        top = max(fam.items(), key=lambda kv: kv[1][0])
        ach = top[1][2] / (top[1][0] * 1e-3) / 1e9
        traffic, traffic_src = None, None
        for tr_name in ("r2_traffic_ratio.json", "r1_traffic_ratio.json"):
            tr_path = os.path.join(ROOT, "profiles", tr_name)
            if os.path.exists(tr_path):
                with open(tr_path) as f:
                    tr = json.load(f)
                if top[0] in tr:        # DRAM bytes / algorithmic bytes measured with ncu --set full for this kernel family
                    traffic = tr[top[0]]["ratio"] * top[1][2] / top[1][1]
                    traffic_src = "profiles/%s: ncu dram bytes / algorithmic bytes = %.2f" % (tr_name, tr[top[0]]["ratio"])
                    break
        roofline = {"kernel": top[0], "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "share_of_step": top[1][0] / tot,
                    "bytes_per_launch": top[1][2] / top[1][1], "avg_launch_ms": top[1][0] / top[1][1],
                    "moved_bytes_per_image_all_kernels": sum(v[2] for v in fam.values()) / B}

    fwd_alg = balg_px * H * W * (value) / 1e9 / world       # GB/s per GPU if the forward were ideally fused

    cpu_baseline, parity, eager = None, None, None
    if rank == 0 and world == 1:
        if not args.no_parity and H * W <= 1200 * 700:
            parity = output_parity(kind, net, x, ratio, sd)
        if not args.no_eager and H * W <= 1200 * 700:
            eager = gpu_eager_reference(kind, H, W, dev)
        if not args.no_cpu_baseline:
            shw = (H, W) if is_lp else (256, 320)
            v, t, cores, how, sample = cpu_reference_rate(kind, H, W, shw, repeats=1, warmup=1)
            cpu_baseline = {"value": v, "unit": "images/s", "cores": cores, "kind": how, "sample": sample}

    if rank == 0:
        line = {
            "metric": metric_name(args.config, kind, H, W), "value": value, "unit": "images/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s; %dx%d, %d images per GPU per step" % (desc, W, H, B), "name": args.config,
                       "weights": weights_desc, "precision": "fp32 I/O; 1x1 convs on tcgen05 in 3xTF32 (FDN_B200_GEMM=%s), everything else fp32" % os.environ.get("FDN_B200_GEMM", "tf32x3"),
                       "l2": "per-step working set (>1 GB of activations) exceeds the 126 MB L2; no explicit flush",
                       "micro_batch": None if is_lp else archs._micro_batch(B, H, W),
                       "ratio_i": "I_predict_net output (script semantics), computed outside the timed region for `value`, inside it for `e2e`"},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "forward_roofline": {"alg_bytes_per_image": balg_px * H * W, "achieved_GBps_per_gpu": fwd_alg,
                                 "frac_of_measured_peak": fwd_alg / peak},
            "fft_stage": fft_stage, "parity": parity, "gpu_eager_reference": eager, "kernel_families": families, "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
