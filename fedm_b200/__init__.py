"""fedm_b200 -- B200-native implementation of FEDM's per-time-step hot path.

The product path is the CUDA library ``libfedmx.so`` (sources in ``fedm_b200/csrc``, C ABI in
``include/fedmx.h``) driven from Python through ctypes; PyTorch tensors are used only as device buffers.
There is no CPU fallback: importing :mod:`fedm_b200.lib` raises if the library is missing.
"""
__version__ = "0.1.0"
