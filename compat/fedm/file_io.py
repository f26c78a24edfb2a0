"""fedm.file_io -> fedm_b200.file_io"""
from fedm_b200.file_io import *      # noqa: F401,F403
from fedm_b200 import file_io as _m
globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
