"""Drop-in replacement for basicsr/models/archs/LPNet_arch.py.  The inference scripts obtain ``transforms`` through the
star-import of this module (LPNet_arch.py:84, inference_fdn_lolblur.py:35), so it is re-exported."""
from torchvision import transforms  # noqa: F401
from fdn_tip2025_b200.archs import I_predict_net  # noqa: F401

__all__ = ["I_predict_net", "transforms"]
