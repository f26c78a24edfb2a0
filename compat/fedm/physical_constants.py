"""fedm.physical_constants -> fedm_b200.physical_constants"""
from fedm_b200.physical_constants import *      # noqa: F401,F403
from fedm_b200 import physical_constants as _m
globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
