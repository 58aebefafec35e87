"""Drop-in boundary: class names, signatures, state_dict schema, ABI header vs binding vs built library (no GPU)."""
import ctypes
import inspect
import os

import pytest
import torch

from abi_util import header_signatures
from fdn_tip2025_b200 import _lib, archs, build, schema, synth


def test_header_matches_ctypes_binding():
    h = header_signatures()
    assert set(h) - {"fdn_last_error_string"} == set(_lib.SIGNATURES)
    for name, codes in _lib.SIGNATURES.items():
        assert h[name] == codes, name


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()            # nvcc cross-compiles for sm_100a without a GPU
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in header_signatures():
        assert hasattr(lib, name), name
    lib.fdn_abi_version.restype = ctypes.c_int
    assert lib.fdn_abi_version() == 1
    lib.fdn_is_device_build.restype = ctypes.c_int
    assert lib.fdn_is_device_build() == 1
    # argument validation happens on the host before any launch: no GPU needed to see the error path
    lib.fdn_fdffn_patch.restype = ctypes.c_int
    lib.fdn_last_error_string.restype = ctypes.c_char_p
    rc = lib.fdn_fdffn_patch(None, None, None, None, 1, 1, 8, 8, None)
    assert rc < 0 and b"bad arguments" in lib.fdn_last_error_string()


def test_sass_is_sm100a():
    path = build.build()
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.parametrize("cls,dim", [(archs.FDN, 32), (archs.FDN_lolv1, 24)])
def test_fdn_state_dict_schema(cls, dim):
    net = cls()
    sd = net.state_dict()
    table = schema.fdn_schema(dim)
    assert len(sd) == 1503 and set(sd) == set(table)
    for k, (shape, _) in table.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    net.load_state_dict(synth.fdn_state_dict(dim=dim, seed=1), strict=True)
    assert all(not p.requires_grad for k, p in net.named_parameters() if k.startswith("net_a."))
    with pytest.raises(RuntimeError):
        bad = synth.fdn_state_dict(dim=dim, seed=1)
        bad.pop("net_p.output.weight")
        net.load_state_dict(bad, strict=True)


def test_other_schemas():
    assert len(archs.MAR().state_dict()) == 168
    assert sum(p.numel() for p in archs.MAR().parameters()) == 143013
    lp = archs.I_predict_net()
    assert len(lp.state_dict()) == 292
    assert lp.state_dict()["conv1.1.num_batches_tracked"].dtype == torch.int64
    assert sum(p.numel() for p in archs.FDN().parameters()) == 8030489
    assert sum(p.numel() for p in archs.FDN_lolv1().parameters()) == 4909805
    fd = archs.FDformer(dim=32, num_blocks=[6, 6, 10])
    assert len(fd.state_dict()) == 1503 - 168 - 6


def test_real_lpnet_checkpoint_loads_strict():
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lpnet_kat.pt")
    if not os.path.exists(path):
        pytest.skip("fixture missing")
    params = torch.load(path)["LPNet_lolblur.pth"]["params"]
    archs.I_predict_net().load_state_dict(params, strict=True)


def test_signatures_match_reference():
    f = inspect.signature(archs.FDN.forward)
    assert list(f.parameters) == ["self", "inp_img", "ori", "device", "ratio_i", "mode"]
    assert [p.default for p in f.parameters.values()][2:] == [None, None, None, 1]
    assert list(inspect.signature(archs.FDN.__init__).parameters) == ["self"]
    f = inspect.signature(archs.FDformer.__init__)
    assert list(f.parameters)[1:] == ["inp_channels", "out_channels", "dim", "num_blocks", "num_refinement_blocks",
                                      "ffn_expansion_factor", "bias"]
    assert f.parameters["dim"].default == 48 and f.parameters["num_blocks"].default == [6, 6, 12, 8]
    f = inspect.signature(archs.FDformer.forward)
    assert list(f.parameters)[1:] == ["inp_img", "ori_img", "x_high1", "x_high2", "x_high3", "x_high12", "x_high22", "x_high32",
                                      "x1", "x2", "x3"]
    assert list(inspect.signature(archs.MAR.forward).parameters) == ["self", "x", "ratio"]
    assert inspect.signature(archs.MAR.__init__).parameters["use_ratio"].default is True
    assert list(inspect.signature(archs.I_predict_net.forward).parameters) == ["self", "x", "use_ori_i"]
    assert inspect.signature(archs.I_predict_net.__init__).parameters["c"].default == 16


def test_no_cpu_fallback():
    """CPU tensors are rejected loudly; nothing under the package imports the oracle."""
    with pytest.raises(RuntimeError, match="CUDA"):
        archs.I_predict_net()(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError, match="CUDA"):
        archs.FDN()(torch.zeros(1, 3, 32, 32), ratio_i=torch.ones(1, 1))
    pkg = os.path.dirname(archs.__file__)
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "import oracle" not in src and "from oracle" not in src and "fdn_oracle" not in src, f


def test_shape_validation():
    from fdn_tip2025_b200 import ops
    saved = ops._device_ok
    ops._device_ok = lambda t: None
    try:
        with pytest.raises(RuntimeError, match="multiples of 32"):
            archs.FDN()(torch.zeros(1, 3, 40, 64), ratio_i=torch.ones(1, 1))
        with pytest.raises(RuntimeError, match="ratio_i"):
            archs.FDN()(torch.zeros(1, 3, 32, 64))
    finally:
        ops._device_ok = saved
