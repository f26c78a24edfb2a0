"""fedm.functions -> fedm_b200.functions"""
from fedm_b200.functions import *      # noqa: F401,F403
from fedm_b200 import functions as _m
globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
