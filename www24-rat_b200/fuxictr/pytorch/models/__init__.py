from .base_model import BaseModel
from .rat import RAT_m0, RAT_m1, RAT_m2, RAT_m3
