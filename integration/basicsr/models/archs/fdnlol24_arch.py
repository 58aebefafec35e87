"""Drop-in replacement for basicsr/models/archs/fdnlol24_arch.py (LOL-v1 variant: FDN_lolv1, dim 24, cat conv live)."""
import torch
import torch.nn as nn
import torch.nn.functional as F
from fdn_tip2025_b200.archs import FDN_lolv1, FDformer  # noqa: F401
from fdn_tip2025_b200.archs import MAR as _MAR


class MAR(_MAR):
    def __init__(self, use_ratio=True):
        super().__init__(use_ratio=use_ratio, variant="lolv1")


__all__ = ["FDN_lolv1", "FDformer", "MAR", "torch", "nn", "F"]
