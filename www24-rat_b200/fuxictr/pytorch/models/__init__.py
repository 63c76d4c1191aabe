"""`getattr(fuxictr.pytorch.models, params["model"])` is how run_expid.py:75 resolves the model class; the four RAT
variants share one engine-backed implementation (rat.py) and differ only in how the engine sequences its kernels."""
from .base_model import BaseModel
from .rat import RAT_m0, RAT_m1, RAT_m2, RAT_m3

__all__ = ["BaseModel", "RAT_m0", "RAT_m1", "RAT_m2", "RAT_m3"]
