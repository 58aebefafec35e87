"""Pin the CPU oracle to the reference's own outputs stored under tests/golden/ (runs anywhere, no GPU)."""
import os

import pytest
import torch

from fdn_tip2025_b200 import synth
from oracle import fdn_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.skip("fixture %s missing (run tests/golden/make_golden.py in the build container)" % name)
    return torch.load(path)


@pytest.mark.parametrize("name", ["fdn_64x96", "fdn_lolv1_64x96", "fdn_96x64_b2"])
def test_oracle_fdn_matches_reference_outputs(name):
    item = _load("fdn_golden.pt")[name]
    sd = synth.fdn_state_dict(dim=item["dim"], seed=item["seed"], damp=item["damp"])
    x = synth.low_light_images(item["b"], item["h"], item["w"])
    variant = "lolblur" if item["kind"] == "FDN" else "lolv1"
    # Primary pin: the fp64 oracle reproduces the reference's fp32 outputs to the reference's own rounding noise.
    got64 = O.fdn(x.double(), item["ratio"].double(), O.to_dtype(sd, torch.float64), variant)
    for i, (g, r) in enumerate(zip(got64, item["outputs"])):
        assert g.shape == r.shape
        # MAR's phase MLPs make its fp32 noise floor ~3e-5; the restored image is tighter
        tol = 2e-6 if (i == 0 or variant == "lolv1") else 2e-4
        assert (g - r.double()).abs().max().item() <= tol, (i, (g - r.double()).abs().max().item())
    assert O.psnr(got64[0], item["outputs"][0]) >= 110.0
    # The fp32 oracle is a *different* fp32 evaluation order of the same graph.  The network is chaotic at FDSA bins
    # whose modulus is at rounding level (angle of a near-zero bin), so two valid
    # fp32 evaluations may differ by O(1e-3) at isolated events (fdn_96x64_b2 has one in encoder_level3.2.attn:
    # 2.4e-3).  Bound it loosely and require PSNR well above the 50 dB gate.
    got = O.fdn(x, item["ratio"], sd, variant)
    assert (got[0] - item["outputs"][0]).abs().max().item() <= 5e-3
    assert O.psnr(got[0], item["outputs"][0]) >= 70.0


@pytest.mark.parametrize("name", ["fdn_64x96", "fdn_lolv1_64x96", "fdn_96x64_b2"])
def test_oracle_fdn_matches_strictly_damped_reference_outputs(name):
    """fdn_golden_strict.pt (project_out x 0.005): chaotic events stay far below 1e-3, so even the fp32 oracle - a different fp32
    evaluation order than the reference - must meet the strict north-star gate against the reference's fp32 outputs."""
    item = _load("fdn_golden_strict.pt")[name]
    sd = synth.fdn_state_dict(dim=item["dim"], seed=item["seed"], damp=item["damp"])
    x = synth.low_light_images(item["b"], item["h"], item["w"])
    variant = "lolblur" if item["kind"] == "FDN" else "lolv1"
    got = O.fdn(x, item["ratio"], sd, variant)
    for g, r in zip(got, item["outputs"]):
        assert (g - r).abs().max().item() <= 1e-3
    assert O.psnr(got[0], item["outputs"][0]) >= 50.0


@pytest.mark.parametrize("name", ["mar_lolblur", "mar_lolv1"])
def test_oracle_mar_matches_reference_outputs(name):
    item = _load("mar_golden.pt")[name]
    sd = synth.mar_state_dict(seed=item["seed"])
    x = synth.low_light_images(2, item["h"], item["w"])
    got = O.mar(x, item["ratio"].view(2, 1, 1, 1), sd, "", item["variant"])
    for g, r in zip(got, item["outputs"]):
        assert (g - r).abs().max().item() <= 2e-5


def test_oracle_blocks_match_reference_hooks():
    fx = _load("block_golden.pt")
    sd = synth.fdn_state_dict(dim=fx["dim"], seed=fx["seed"], damp=None)
    for mod_name, item in fx["blocks"].items():
        p = mod_name + "."
        ins = item["inputs"]
        if item["kind"] == "fdsa":
            got = O.fdsa(ins[0], sd, p)
        elif item["kind"] == "fdffn":
            got = O.fdffn(ins[0], sd, p)
        elif item["kind"] == "fcaffn":
            got = O.fcaffn(ins[0], ins[1], ins[2], ins[3], sd, p)
        else:
            got = O.fuse(ins[0], ins[1], sd, p)
        ref = item["output"]
        rel = ((got - ref).norm() / ref.norm()).item()
        assert rel <= 2e-6, (mod_name, rel)


def test_oracle_lpnet_known_answers():
    fx = _load("lpnet_kat.pt")
    for name, item in fx.items():
        for seed, h, w, expect in item["kats"]:
            x = torch.rand(2, 3, h, w, generator=torch.Generator().manual_seed(seed)) * 0.2
            y = O.lpnet(x, item["params"]).flatten()
            assert torch.allclose(y, torch.tensor(expect), atol=2e-6), (name, seed, y)


def test_oracle_edge_cases():
    """Algebraic identities the kernels rely on hold in the oracle's arithmetic."""
    z = torch.complex(torch.tensor([0.0, 1e-12, -1e-11, 2.0, -3.0]), torch.tensor([0.0, -5e-11, 1.0, 1e-10, 0.0]))
    r = O.rd(z)
    assert torch.equal(r.real, torch.tensor([1e-10, 1e-10, 1e-10, 2.0, -3.0]))
    assert torch.equal(r.imag, torch.tensor([1e-10, 1e-10, 1.0, 1e-10, 1e-10]))
    x = torch.rand(2, 3, 8, 12, dtype=torch.float64)
    assert torch.allclose(O.half(x), torch.nn.functional.interpolate(x, scale_factor=0.5, mode="bilinear", align_corners=False))
    assert torch.equal(O.nearest_down(x), torch.nn.functional.interpolate(x, scale_factor=0.5))
    assert torch.equal(O.nearest_up(x), torch.nn.functional.interpolate(x, scale_factor=2))
    # irfft2 with s smaller than the spectrum slices it (fourier_fuse)
    s = torch.fft.rfft2(torch.rand(10, 14, dtype=torch.float64))
    assert torch.allclose(torch.fft.irfft2(s, s=(8, 12)), torch.fft.irfft2(s[:8, :7], s=(8, 12)))
