"""GPU parity tests: the CUDA path (libfdn_b200.so through the C ABI) against the CPU oracle.  Run with -m gpu."""
import os

import pytest
import torch

import parity_cases as P

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# FFT sizes of every BASELINE config, incl. the (H+2)x(W+2) fourier_fuse transforms
FFT_SIZES = [(8, 8), (64, 64), (256, 256), (258, 258), (130, 130), (104, 152), (208, 304), (416, 608), (418, 610), (210, 306),
             (160, 280), (320, 560), (640, 1120), (322, 562), (642, 1122), (30, 14), (13, 22)]


@pytest.mark.parametrize("h,w", FFT_SIZES)
def test_fft_sizes(cuda_dev, h, w):
    P.case_rfft2_irfft2(cuda_dev, h, w)


def test_fft_4k_roundtrip(cuda_dev):
    """Largest config (3840x2160 padded to 2176 rows) through the size-independent round-trip property."""
    P.case_rfft2_irfft2(cuda_dev, 2176, 3840, planes=1)
    P.case_rfft2_irfft2(cuda_dev, 1090, 1922, planes=1)


def test_irfft2_nonhermitian(cuda_dev):
    P.case_irfft2_nonhermitian(cuda_dev, 64, 96)
    P.case_irfft2_nonhermitian(cuda_dev, 160, 280)


@pytest.mark.parametrize("h,w", [(640, 1120), (320, 560), (160, 280), (256, 256), (64, 64), (416, 608), (104, 152), (16, 24)])
def test_fcaffn_fft_stage(cuda_dev, h, w):
    """rows R2C -> columns forward + FCAFFN modulation + inverse -> rows C2R at every level of the BASELINE configs, vs torch float64;
    also with phases beyond the fast sincos range (library path of the column kernel)."""
    P.case_fcaffn_fft_stage(cuda_dev, h, w, b=2, c=3)
    P.case_fcaffn_fft_stage(cuda_dev, h, w, b=1, c=2, big_phase=True)


def test_pw_conv(cuda_dev):
    P.case_pw_conv(cuda_dev)


def test_conv2d(cuda_dev):
    P.case_conv2d(cuda_dev)


def test_conv3x3_tensor_core(cuda_dev):
    P.case_conv3x3_mma(cuda_dev)
    P.case_conv3x3_mma(cuda_dev, shapes=((64, 32, 640, 1120),))       # up2_1 at the benchmarked size


def test_convt_dw_misc(cuda_dev):
    P.case_convt_dw_misc(cuda_dev)


@pytest.mark.parametrize("dim,h,w", [(32, 64, 96), (64, 32, 48), (128, 16, 24), (24, 32, 64), (48, 16, 32), (96, 8, 16)])
def test_transformer_block(cuda_dev, dim, h, w):
    P.case_tblock(cuda_dev, dim, h, w, True, True, seed=dim)


def test_fuse_and_resamplers(cuda_dev):
    P.case_fuse_resample(cuda_dev, 32, 32, 48)
    P.case_fuse_resample(cuda_dev, 64, 16, 24)


@pytest.mark.parametrize("variant", ["lolblur", "lolv1"])
def test_mar(cuda_dev, variant):
    P.case_mar(cuda_dev, 64, 96, variant)


def test_mar_nonsmooth_size(cuda_dev):
    P.case_mar(cuda_dev, 104, 152, "lolv1")    # 1/4 scale of the LOL-v1 shape: radices 13 and 19, fuse sizes 106x154


def test_lpnet_synthetic(cuda_dev):
    P.case_lpnet(cuda_dev, 256, 256)
    P.case_lpnet(cuda_dev, 160, 96)


def test_lpnet_real_checkpoint_fixture(cuda_dev):
    """Known answers of the real LPNet checkpoints stored in tests/golden/lpnet_kat.pt with the weights."""
    path = os.path.join(GOLDEN, "lpnet_kat.pt")
    if not os.path.exists(path):
        pytest.skip("fixture not generated")
    from fdn_tip2025_b200 import archs
    fx = torch.load(path)
    for name, item in fx.items():
        net = archs.I_predict_net()
        net.load_state_dict(item["params"], strict=True)
        net = net.to(cuda_dev)
        for seed, h, w, expect in item["kats"]:
            x = torch.rand(2, 3, h, w, generator=torch.Generator().manual_seed(seed)) * 0.2
            y = net(x.to(cuda_dev)).cpu().flatten()
            assert torch.allclose(y, torch.tensor(expect), atol=2e-6), (name, seed, y, expect)


E2E_CASES = [("FDN", 64, 96, 1), ("FDN", 128, 160, 2), ("FDN_lolv1", 96, 160, 1)]


@pytest.mark.parametrize("kind,h,w,b", E2E_CASES)
def test_fdn_end_to_end(cuda_dev, kind, h, w, b):
    """Default mode (tcgen05 3xTF32 GEMMs) against the fp64 oracle: PSNR gate + bounded chaotic events (see case_fdn)."""
    P.case_fdn(cuda_dev, kind, h, w, b=b, strict=False)


STRICT_DAMP = 0.005     # tests/golden/make_golden.py::STRICT_DAMP


@pytest.mark.parametrize("kind,h,w,b", E2E_CASES)
def test_fdn_end_to_end_ffma_strict(cuda_dev, kind, h, w, b, monkeypatch):
    """All GEMMs on the fp32 FFMA kernel: the strict north-star gate (max-abs <= 1e-3, PSNR >= 50 dB) vs the fp64 oracle.

    Weights: net_p project_out damped by 0.005.  With the 0.03 of the other end-to-end tests the gate depends on luck: whether a
    weight seed meets an isolated chaotic FDSA sign event (a few 1e-3 on ~1 % of the pixels) is decided by
    the rounding order of the evaluation, not by kernel accuracy - profiles/r1_v7_strict_gate_seed_sweep.txt shows two FFT
    implementations of identical accuracy meeting one event each in six seeds, on different seeds.  At 0.005 such an event moves
    the output by well under 1e-3, so the gate measures what it is meant to: fp32-level agreement through all 50 blocks."""
    monkeypatch.setenv("FDN_B200_GEMM", "ffma")
    for seed in (7, 8):
        P.case_fdn(cuda_dev, kind, h, w, b=b, strict=True, seed=seed, damp=STRICT_DAMP)


@pytest.mark.parametrize("kind,h,w,b", E2E_CASES)
def test_fdn_end_to_end_tf32x3_strict(cuda_dev, kind, h, w, b, monkeypatch):
    """The same strict gate on the product's default GEMM path (tcgen05, 3xTF32): if the split is as accurate as FFMA it must pass
    the identical test."""
    monkeypatch.setenv("FDN_B200_GEMM", "tf32x3")
    for seed in (7, 8):
        P.case_fdn(cuda_dev, kind, h, w, b=b, strict=True, seed=seed, damp=STRICT_DAMP)


# ---- the benchmarked configurations themselves (BASELINE.json configs 1-3), default GEMM mode, against the fp64 oracle on the host
@pytest.mark.parametrize("kind,h,w", [("FDN", 640, 1120), ("FDN_lolv1", 416, 608), ("FDN", 256, 256)])
def test_fdn_full_size_strict(cuda_dev, kind, h, w, monkeypatch):
    """North-star gate (max-abs <= 1e-3 on [0,1] outputs, PSNR >= 50 dB) at the sizes bench.py times, one image, the weights
    bench.py uses (seed 0, project_out x 0.005)."""
    monkeypatch.delenv("FDN_B200_GEMM", raising=False)
    rep = []
    P.case_fdn(cuda_dev, kind, h, w, b=1, strict=True, seed=0, damp=STRICT_DAMP, report=rep)
    print("full-size parity", rep)


FULL_BLOCKS = [(32, 640, 1120), (64, 320, 560), (128, 160, 280), (24, 416, 608), (48, 208, 304), (96, 104, 152)]


@pytest.mark.parametrize("dim,h,w", FULL_BLOCKS)
def test_transformer_block_full_size(cuda_dev, dim, h, w):
    """FDSA + FDFFN + FCAFFN at every level of the 1120x640 (dim 32) and 608x416 (dim 24) configurations: thousands of pixel tiles
    per CTA through the persistent tcgen05 kernel (ring / TMEM phase wrap), rel-L2 <= 1e-5 vs the fp64 oracle.  FDSA is gated per
    8x8 patch (parity_cases.compare_patchwise): at these sizes any fp32 evaluation, torch's included, meets an isolated sign event."""
    rep = []
    P.case_tblock(cuda_dev, dim, h, w, True, True, seed=dim + 1, b=1, report=rep, patchwise=True)
    print("full-size blocks", rep)


def test_fcaffn_block_4k_level2(cuda_dev):
    """One FCAFFN + FDFFN block at level 2 of the 3840x2160 configuration (1088x1920, generic radix-17 FFT path)."""
    P.case_tblock(cuda_dev, 64, 1088, 1920, False, True, seed=3, b=1)


@pytest.mark.parametrize("name", ["zeros", "half", "ones", "one_hot_pixel"])
def test_fdn_edge_inputs(cuda_dev, name):
    """Degenerate frames: exactly-zero / constant / saturated planes make whole 8x8 spectra exactly zero, so every bin goes
    through replace_denormals (+1e-10, phase pi/4) and the rsqrt moduli of the FDSA algebra (patch_spectral.cu)."""
    h, w = 64, 96
    x = {"zeros": torch.zeros(1, 3, h, w), "half": torch.full((1, 3, h, w), 0.5), "ones": torch.ones(1, 3, h, w)}.get(name)
    if x is None:
        x = torch.zeros(1, 3, h, w)
        x[0, :, 17, 41] = 1.0
    P.case_fdn(cuda_dev, "FDN", h, w, b=1, strict=True, seed=7, damp=STRICT_DAMP, images=x)


def test_ratio_broadcast_and_device_guard(cuda_dev):
    """A single ratio value broadcasts over the batch like the reference's [1,1,1,1] tensor; mismatched counts raise."""
    from fdn_tip2025_b200 import archs, synth
    net = archs.MAR()
    net.load_state_dict(synth.mar_state_dict(seed=5), strict=True)
    net = net.to(cuda_dev)
    x = synth.low_light_images(2, 32, 48).to(cuda_dev)
    a = net(x, torch.tensor([[0.3]], device=cuda_dev))
    b = net(x, torch.tensor([[0.3], [0.3]], device=cuda_dev))
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    with pytest.raises(RuntimeError):
        net(x, torch.tensor([[0.3], [0.3], [0.3]], device=cuda_dev))


def test_reference_module_paths_resolve(cuda_dev):
    """The shims under integration/ expose the reference's module paths and class names (archs/__init__.py:43-46 looks classes up
    by name); a star-import provides everything inference_fdn_lolblur.py uses."""
    import importlib.util
    import sys
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "integration", "basicsr", "models", "archs")
    ns = {}
    for mod in ("FDN_arch", "fdnlol24_arch", "mar_arch", "LPNet_arch"):
        spec = importlib.util.spec_from_file_location("shim_" + mod, os.path.join(root, mod + ".py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        for k in getattr(m, "__all__"):
            ns[k] = getattr(m, k)
    for name in ("FDN", "FDN_lolv1", "FDformer", "MAR", "I_predict_net", "transforms"):
        assert name in ns, name
    from fdn_tip2025_b200 import synth
    net = ns["FDN"]()                                    # define_network({'type': 'FDN'})-style construction: no arguments
    net.load_state_dict(synth.fdn_state_dict(dim=32, seed=4, damp=0.005), strict=True)
    net = net.to(cuda_dev).eval()
    x = synth.low_light_images(1, 32, 32).to(cuda_dev)
    out = net(x, ratio_i=torch.tensor([[0.3]], device=cuda_dev), device=cuda_dev)
    assert len(out) == 4 and out[0].shape == x.shape


@pytest.mark.parametrize("dim", [48, 32])
def test_fdformer_standalone(cuda_dev, dim):
    """The stand-alone FDformer module (FDN_arch.py:753-842) with its reference default dim = 48 (level-3 layers wider than the
    tensor-core kernel's K limit run on the CUDA-core GEMM) and with dim = 32, side maps supplied by the caller, vs the fp64 oracle."""
    from fdn_tip2025_b200 import archs, schema, synth
    table = schema.fdformer_schema(dim)
    sd = synth.make_state_dict(table, seed=9)
    for k in sd:                                            # damp the residual branches like the FDN fixtures
        if k.endswith("project_out.weight"):
            sd[k] = sd[k] * 0.005
    net = archs.FDformer(dim=dim)
    net.load_state_dict(sd, strict=True)
    net = net.to(cuda_dev).eval()
    b, h, w = 1, 64, 96
    img = synth.low_light_images(b, h, w)
    sides, sides64 = [], []
    for lvl in range(3):
        hl, wl = h >> lvl, w >> lvl
        amp = P.rnd(b, 3, hl, wl // 2 + 1, seed=lvl + 1).abs() * 3
        pha = P.rnd(b, 3, hl, wl // 2 + 1, seed=lvl + 4) * 3.14
        im = P.rnd(b, 3, hl, wl, seed=lvl + 7).abs()
        sides.append(tuple(P.dev32(t, cuda_dev) for t in (amp, pha, im)))
        sides64.append(tuple(t.float().double() for t in (amp, pha, im)))
    got = net(img.to(cuda_dev), x_high1=sides[0][0], x_high12=sides[0][1], x1=sides[0][2], x_high2=sides[1][0], x_high22=sides[1][1],
              x2=sides[1][2], x_high3=sides[2][0], x_high32=sides[2][1], x3=sides[2][2])
    ref = P.O.fdformer(img.double(), sides64[0], sides64[1], sides64[2], P._sd64(sd), p="")
    d = (got.double().cpu() - ref).abs().max().item()
    assert d <= 1e-3 and P.O.psnr(got.cpu(), ref) >= 50.0, (dim, d)
    with pytest.raises(RuntimeError):                       # side maps are validated (shape, device)
        net(img.to(cuda_dev), x_high1=sides[1][0], x_high12=sides[0][1], x1=sides[0][2], x_high2=sides[1][0], x_high22=sides[1][1],
            x2=sides[1][2], x_high3=sides[2][0], x_high32=sides[2][1], x3=sides[2][2])


def _golden_replay(cuda_dev, strict):
    path = os.path.join(GOLDEN, "fdn_golden_strict.pt" if strict else "fdn_golden.pt")
    if not os.path.exists(path):
        pytest.skip("fixture not generated")
    from fdn_tip2025_b200 import archs, synth
    fx = torch.load(path)
    for name, item in fx.items():
        sd = synth.fdn_state_dict(dim=item["dim"], seed=item["seed"], damp=item["damp"])
        net = getattr(archs, item["kind"])()
        net.load_state_dict(sd, strict=True)
        net = net.to(cuda_dev)
        x = synth.low_light_images(item["b"], item["h"], item["w"])
        got = net(x.to(cuda_dev), ratio_i=item["ratio"].to(cuda_dev))
        for g, r in zip(got, item["outputs"]):
            d = (g.cpu().double() - r.double()).abs()
            if strict:
                assert d.max().item() <= 1e-3, (name, d.max().item())
            else:
                # isolated chaotic events (FDSA phase of a rounding-level bin) may exceed 1e-3 at a
                # few pixels for ANY other fp32 evaluation order - the fp32 CPU oracle shows the same on fdn_96x64_b2
                assert d.max().item() <= 5e-2, (name, d.max().item())
                assert (d > 1e-3).double().mean().item() <= 5e-2, (name, (d > 1e-3).double().mean().item())
        assert P.O.psnr(got[0].cpu(), item["outputs"][0]) >= 50.0


def test_fdn_golden_fixture(cuda_dev):
    """Replay the reference's own fp32 outputs (tests/golden/make_golden.py) with the default tcgen05 3xTF32 GEMMs."""
    _golden_replay(cuda_dev, strict=False)


def test_fdn_golden_fixture_ffma_strict(cuda_dev, monkeypatch):
    """The reference's own fp32 outputs for the strictly damped weights (fdn_golden_strict.pt, project_out x 0.005; see
    test_fdn_end_to_end_ffma_strict) with every GEMM on the fp32 FFMA kernel: max-abs <= 1e-3 and PSNR >= 50 dB on every output."""
    monkeypatch.setenv("FDN_B200_GEMM", "ffma")
    _golden_replay(cuda_dev, strict=True)


def test_full_size_properties(cuda_dev):
    """BASELINE full size (640x1120): FCAFFN stage linearity in the amplitude map and FFT round trip."""
    P.case_rfft2_irfft2(cuda_dev, 640, 1120, planes=2)
    P.case_tblock(cuda_dev, 32, 320, 560, False, True, seed=11)


# FDformer 1x1-conv shapes: (K, N, prologue)
MMA_SHAPES = [(32, 152, 1), (114, 32, 2), (32, 86, 1), (86, 32, 0), (32, 32, 3), (64, 304, 1), (228, 64, 2), (172, 64, 0),
              (128, 612, 1), (459, 128, 2), (345, 128, 0), (128, 128, 3), (96, 460, 1), (345, 96, 2), (259, 96, 0), (192, 192, 0)]


@pytest.mark.parametrize("k,n,prologue", MMA_SHAPES)
def test_tcgen05_conv1x1(cuda_dev, k, n, prologue):
    """tcgen05 / TMEM 1x1 convolution in 3xTF32 mode: fp32-level accuracy against torch float64."""
    P.case_pw_mma(cuda_dev, k, n, hw=(16, 24), prologue=prologue, two_src=(prologue == 0 and k % 3 == 0))
    P.case_pw_mma(cuda_dev, k, n, hw=(12, 19), prologue=prologue)          # ragged last pixel tile (HW % 128 != 0)


def test_tcgen05_conv1x1_single_pass_tf32(cuda_dev):
    """Single-pass TF32 variant: stated tolerance 1e-3 relative (10-bit mantissa operands)."""
    P.case_pw_mma(cuda_dev, 128, 256, hw=(64, 64), prologue=1, passes=1, tol=(1e-3, 5e-3))


def test_ffma_path_still_matches(cuda_dev, monkeypatch):
    monkeypatch.setenv("FDN_B200_GEMM", "ffma")
    P.case_tblock(cuda_dev, 32, 32, 48, True, True, seed=5)
    P.case_fuse_resample(cuda_dev, 32, 16, 24)


@pytest.mark.parametrize("fused", ["1", "0"])
def test_fdffn_fused_variant(cuda_dev, monkeypatch, fused):
    """FDFFN middle section: the fused per-patch kernel (dw-GELU-dw + patch FFT + sum, the default) and the two-kernel form
    (FDN_B200_FDFFN_FUSED=0) against the fp64 oracle, on sizes whose patch rows wrap inside a warp."""
    monkeypatch.setenv("FDN_B200_FDFFN_FUSED", fused)
    P.case_tblock(cuda_dev, 32, 40, 72, False, False, seed=21)
    P.case_tblock(cuda_dev, 64, 32, 32, False, False, seed=22)
    P.case_tblock(cuda_dev, 24, 8, 8, False, False, seed=23)        # a single patch: every halo element is image border


def test_image_pre_post_bit_exact(cuda_dev):
    """uint8 HWC BGR <-> fp32 CHW RGB, reflect pad, crop, clamp, round: bit-exact against the scripts' host code."""
    P.case_imgio(cuda_dev)
    P.case_imgio(cuda_dev, h=400, w=600, b=1)        # LOL-v1 frame -> 416x608
    P.case_imgio(cuda_dev, h=33, w=64, b=3)


@pytest.mark.parametrize("variant,kind", [("lolblur", "FDN"), ("lolv1", "FDN_lolv1")])
def test_inference_pipeline_matches_script_semantics(cuda_dev, variant, kind):
    """InferencePipeline (uint8 in, uint8 out, everything on the device) against the script's own sequence of host steps around the
    same modules - must be bit-identical - and against the fp64 oracle run through the same steps (at most 1 LSB, rarely)."""
    from fdn_tip2025_b200 import archs, pipeline, synth
    h, w = 50, 70
    dim = 32 if kind == "FDN" else 24
    sd = synth.fdn_state_dict(dim=dim, seed=4, damp=0.005)
    net = getattr(archs, kind)()
    net.load_state_dict(sd, strict=True)
    net = net.to(cuda_dev).eval()
    lsd = synth.lpnet_state_dict(seed=3)
    lp = archs.I_predict_net()
    lp.load_state_dict(lsd, strict=True)
    lp = lp.to(cuda_dev).eval()
    img = (synth.low_light_images(1, h, w)[0].permute(1, 2, 0) * 255).round().to(torch.uint8).flip(-1).contiguous()   # BGR frame
    got = torch.from_numpy(pipeline.InferencePipeline(net, lp, variant)(img.numpy()))
    # the script's host steps around our modules
    x = P.script_pre(img[None]).to(cuda_dev)
    ratio = lp(x)
    if variant == "lolv1":
        gray = (0.2989 * x[:, 0] + 0.587 * x[:, 1] + 0.114 * x[:, 2]).mean(dim=(1, 2)).view(1, 1)
        ratio_in = gray / ratio
    else:
        ratio_in = ratio
    want = P.script_post(net(x, ratio_i=ratio_in)[0], h, w)[0]
    if variant == "lolblur":
        assert torch.equal(got, want)
    else:       # the gray mean is reduced in a different order on the device: allow the last bit of ratio_i to differ
        assert (got.int() - want.int()).abs().max().item() <= 1
    # fp64 oracle through the same steps
    x64 = P.script_pre(img[None]).double()
    r64 = P.O.lpnet(x64, P.O.to_dtype(lsd, torch.float64))
    if variant == "lolv1":
        r64 = (0.2989 * x64[:, 0] + 0.587 * x64[:, 1] + 0.114 * x64[:, 2]).mean(dim=(1, 2)).view(1, 1) / r64
    ref = P.script_post(P.O.fdn(x64, r64, P._sd64(sd), variant)[0], h, w)[0]
    d = (got.int() - ref.int()).abs()
    assert d.max().item() <= 1 and (d > 0).float().mean().item() <= 1e-2


def test_inference_pipeline_cuda_graph_replay(cuda_dev):
    """CUDA-graph replay of the device-side sequence is bit-identical to the eager launches, also on the second frame."""
    from fdn_tip2025_b200 import archs, pipeline, synth
    net = archs.FDN()
    net.load_state_dict(synth.fdn_state_dict(dim=32, seed=4, damp=0.005), strict=True)
    net = net.to(cuda_dev).eval()
    lp = archs.I_predict_net()
    lp.load_state_dict(synth.lpnet_state_dict(seed=3), strict=True)
    lp = lp.to(cuda_dev).eval()
    eager = pipeline.InferencePipeline(net, lp, "lolblur")
    graphed = pipeline.InferencePipeline(net, lp, "lolblur", use_graphs=True)
    for i in range(3):
        img = (synth.low_light_images(1, 50, 70, first_index=i)[0].permute(1, 2, 0) * 255).round().to(torch.uint8).flip(-1).contiguous().numpy()
        assert (eager(img) == graphed(img)).all()


def test_micro_batching_is_exact(cuda_dev, monkeypatch):
    """A batch processed in micro-batches (FDN_B200_MICRO_BATCH) gives bit-identical images to one image at a time: no op mixes images."""
    from fdn_tip2025_b200 import archs, synth
    net = archs.FDN()
    net.load_state_dict(synth.fdn_state_dict(dim=32, seed=5, damp=0.005), strict=True)
    net = net.to(cuda_dev).eval()
    x = synth.low_light_images(3, 64, 96).to(cuda_dev)
    ratio = torch.tensor([[0.3], [0.35], [0.4]], device=cuda_dev)
    monkeypatch.setenv("FDN_B200_MICRO_BATCH", "2")
    whole = net(x, ratio_i=ratio)
    monkeypatch.setenv("FDN_B200_MICRO_BATCH", "1")
    for i in range(3):
        one = net(x[i:i + 1], ratio_i=ratio[i:i + 1])
        for a, b in zip(whole, one):
            assert torch.equal(a[i:i + 1], b)
