"""Pin the oracle and the state_dict schema to the reference itself (build container only).

Run:  python oracle/validate_against_reference.py
Checks, for FDN / FDN_lolv1 / MAR (three files) / I_predict_net:
  * schema.py enumerates exactly the reference's state_dict keys and shapes;
  * a synthetic state_dict loads into the reference with strict=True;
  * oracle outputs == reference outputs on the same inputs (fp32: ~1e-6; fp64: ~1e-12);
  * the real LPNet checkpoints reproduce the known answers in BASELINE.md.
Writes oracle/VALIDATION.txt (committed) with the measured differences.
"""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fdn_tip2025_b200 import schema, synth  # noqa: E402
from oracle import fdn_oracle as O  # noqa: E402
from oracle import ref_loader as R  # noqa: E402

LINES = []


def log(s):
    print(s, flush=True)
    LINES.append(s)


def check_schema(net, table, name):
    ref = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    mine = {k: tuple(s) for k, (s, _) in table.items()}
    missing = sorted(set(ref) - set(mine))
    extra = sorted(set(mine) - set(ref))
    bad = sorted(k for k in set(ref) & set(mine) if ref[k] != mine[k])
    log("schema %-14s keys ref=%d mine=%d missing=%d extra=%d shape-mismatch=%d"
        % (name, len(ref), len(mine), len(missing), len(extra), len(bad)))
    assert not missing and not extra and not bad, (missing[:5], extra[:5], bad[:5])


def maxdiff(a, b):
    return (a.double() - b.double()).abs().max().item()


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)

    # ---- LPNet
    net = R.build("I_predict_net")
    check_schema(net, schema.lpnet_schema(), "I_predict_net")
    sd = synth.lpnet_state_dict(seed=3)
    net.load_state_dict(sd, strict=True)
    x = synth.low_light_images(2, 96, 160)
    log("lpnet synthetic fp32 max|oracle-ref| = %.3e" % maxdiff(O.lpnet(x, sd), R.run(net, x)))
    log("lpnet use_ori_i fp32 max|oracle-ref| = %.3e" % maxdiff(O.lpnet(x, sd, True), R.run(net, x, use_ori_i=True)))
    for ck, kats in (("LPNet_lolblur.pth", ((0, 256, 256, 0.2576584, 0.2509128), (1, 416, 608, 0.3614573, 0.3529572))),
                     ("LPNet_lolv1.pth", ((0, 256, 256, 0.3792925, 0.3849431), (1, 416, 608, 0.3662928, 0.3647223)))):
        real = torch.load(os.path.join(R.REF_ROOT, "checkpoint", ck), map_location="cpu")["params"]
        for seed, h, w, a0, a1 in kats:
            xi = torch.rand(2, 3, h, w, generator=torch.Generator().manual_seed(seed)) * 0.2
            y = O.lpnet(xi, real)
            log("lpnet %s seed %d %dx%d oracle=(%.7f, %.7f) KAT=(%.7f, %.7f)" % (ck, seed, h, w, y[0, 0], y[1, 0], a0, a1))
            assert abs(y[0, 0].item() - a0) < 2e-6 and abs(y[1, 0].item() - a1) < 2e-6

    # ---- MAR (three reference files share one schema)
    for kind, variant in (("MAR", "lolblur"), ("MAR_lolv1", "lolv1"), ("MAR_standalone", "lolblur")):
        net = R.build(kind)
        check_schema(net, schema.mar_schema(), kind)
        sd = synth.mar_state_dict(seed=5)
        net.load_state_dict(sd, strict=True)
        x = synth.low_light_images(2, 64, 96)
        ratio = torch.tensor([[0.3], [0.45]]).view(2, 1, 1, 1)
        ref = R.run(net, x, ratio)
        mine = O.mar(x, ratio, sd, "", variant)
        log("%s fp32 max|oracle-ref| = %s" % (kind, ["%.2e" % maxdiff(a, b) for a, b in zip(mine, ref)]))
        net64 = copy.deepcopy(net).double()
        ref64 = R.run(net64, x.double(), ratio.double())
        mine64 = O.mar(x.double(), ratio.double(), O.to_dtype(sd, torch.float64), "", variant)
        d = [maxdiff(a, b) for a, b in zip(mine64, ref64)]
        log("%s fp64 max|oracle-ref| = %s" % (kind, ["%.2e" % v for v in d]))
        assert max(d) < 1e-9

    # ---- FDN / FDN_lolv1
    for kind, dim, variant in (("FDN", 32, "lolblur"), ("FDN_lolv1", 24, "lolv1")):
        net = R.build(kind)
        check_schema(net, schema.fdn_schema(dim), kind)
        sd = synth.fdn_state_dict(dim=dim, seed=7, damp=0.03)
        net.load_state_dict(sd, strict=True)
        x = synth.low_light_images(1, 64, 96)
        ratio = torch.tensor([[0.35]])
        ref = R.run(net, x, ratio_i=ratio)
        mine = O.fdn(x, ratio, sd, variant)
        log("%s fp32 max|oracle-ref| = %s  psnr(out)=%.1f dB"
            % (kind, ["%.2e" % maxdiff(a, b) for a, b in zip(mine, ref)], O.psnr(mine[0], ref[0])))
        # fp64: the reference calls .float() before every FFT, so neutralise it while running in double
        net64 = copy.deepcopy(net).double()
        real_float = torch.Tensor.float
        torch.Tensor.float = lambda self, *a, **k: self
        try:
            ref64 = R.run(net64, x.double(), ratio_i=ratio.double())
        finally:
            torch.Tensor.float = real_float
        mine64 = O.fdn(x.double(), ratio.double(), O.to_dtype(sd, torch.float64), variant)
        d = [maxdiff(a, b) for a, b in zip(mine64, ref64)]
        log("%s fp64 max|oracle-ref| = %s" % (kind, ["%.2e" % v for v in d]))
        log("%s reference noise floor fp32-vs-fp64 = %.2e ; oracle fp32-vs-fp64 = %.2e"
            % (kind, maxdiff(ref[0], ref64[0]), maxdiff(mine[0], mine64[0])))
        assert max(d) < 1e-7

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "VALIDATION.txt"), "w") as f:
        f.write("oracle/validate_against_reference.py, torch %s, reference at %s\n" % (torch.__version__, R.REF_ROOT))
        f.write("\n".join(LINES) + "\n")
    log("OK")


if __name__ == "__main__":
    main()
