"""Drop-in replacement for basicsr/models/archs/mar_arch.py (stand-alone MAR used for MAR pre-training checkpoints)."""
from fdn_tip2025_b200.archs import MAR  # noqa: F401

__all__ = ["MAR"]
