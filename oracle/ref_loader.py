"""Import the reference arch files by path (build container only; TEST INFRASTRUCTURE).

``/root/reference`` does not exist on the GPU box; there the files staged by ``oracle/stage_ref.py`` under
``oracle/_ref/`` are used (bench.py's reference arm / cpu_baseline leg).  Otherwise only
``oracle/validate_against_reference.py`` and ``tests/golden/make_golden.py`` use this.  The arch files are loaded one by one with importlib because
``import basicsr`` needs lmdb/skimage which are not installed; the hard-coded
``torch.load('/data/tuluwei/...')`` inside ``FDN.__init__`` (FDN_arch.py:860-862) is stubbed while
constructing.
"""
import contextlib
import importlib.util
import inspect
import io
import os

import torch

REF_ROOT = os.environ.get("FDN_REFERENCE_ROOT", "/root/reference")
ARCH_DIR = os.path.join(REF_ROOT, "basicsr", "models", "archs")
if not os.path.isfile(os.path.join(ARCH_DIR, "FDN_arch.py")):
    # GPU box: the unmodified arch files staged by oracle/stage_ref.py (git-ignored, shipped with the snapshot)
    ARCH_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "basicsr", "models", "archs")


def available():
    return os.path.isfile(os.path.join(ARCH_DIR, "FDN_arch.py"))


def load_arch(name):
    spec = importlib.util.spec_from_file_location("_fdn_ref_" + name, os.path.join(ARCH_DIR, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build(kind):
    """kind in {'FDN', 'FDN_lolv1', 'MAR', 'MAR_lolv1', 'MAR_standalone', 'I_predict_net'} -> reference nn.Module (eval)."""
    if kind in ("FDN", "FDN_lolv1"):
        mod = load_arch("FDN_arch" if kind == "FDN" else "fdnlol24_arch")
        real = torch.load

        def stub(*a, **k):
            return {"params": inspect.currentframe().f_back.f_locals["self"].net_a.state_dict()}

        torch.load = stub
        try:
            net = getattr(mod, kind)()
        finally:
            torch.load = real
    elif kind == "MAR":
        net = load_arch("FDN_arch").MAR()
    elif kind == "MAR_lolv1":
        net = load_arch("fdnlol24_arch").MAR()
    elif kind == "MAR_standalone":
        net = load_arch("mar_arch").MAR()
    elif kind == "I_predict_net":
        net = load_arch("LPNet_arch").I_predict_net()
    else:
        raise ValueError(kind)
    return net.eval()


def run(net, *args, **kwargs):
    """Forward without grad, swallowing the print() inside MAR_archa.forward (FDN_arch.py:211)."""
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        return net(*args, **kwargs)
