"""fedm.utils -> fedm_b200.utils"""
from fedm_b200.utils import *      # noqa: F401,F403
from fedm_b200 import utils as _m
globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
