"""Drop-in replacement for basicsr/models/archs/FDN_arch.py (copy this file over the reference's).

``from basicsr.models.archs.FDN_arch import *`` in inference_fdn_lolblur.py keeps working: FDN, FDformer, MAR and the
helper names below resolve to the B200 kernels in fdn_tip2025_b200 (which must be importable, e.g. via PYTHONPATH).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from fdn_tip2025_b200.archs import FDN, FDformer, I_predict_net  # noqa: F401
from fdn_tip2025_b200.archs import MAR as _MAR


class MAR(_MAR):
    # FDN_arch.MAR multiplies by ratio unconditionally (FDN_arch.py:213-219, 264)
    _always_ratio = True


__all__ = ["FDN", "FDformer", "MAR", "torch", "nn", "F"]
