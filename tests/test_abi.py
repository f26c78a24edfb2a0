"""CPU checks of the drop-in boundary: libfedmx.so loads and exports every symbol include/fedmx.h declares;
host-only entry points (controllers) are bit-exact with the reference's Python arithmetic."""
import ctypes
import re
from pathlib import Path

import pytest

from fedm_b200 import lib as L

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "fedmx.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fedmx_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n
        assert n in L.SIGNATURES, f"{n} has no ctypes signature"
    assert set(L.SIGNATURES) == set(names)
    assert b"sm_100a" in lib.fedmx_version()


def test_ctx_create_fails_cleanly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = L.load()
    h = ctypes.c_void_p()
    assert lib.fedmx_ctx_create(0, ctypes.byref(h)) != 0
    assert not h.value


def test_pid_controller_bit_exact_on_golden_logs(golden_dir):
    from fedm_b200.hotpath import pid_step
    for case, ttol, dt_min, dt_max in (("streamer_discharge", 1e-3, 1e-15, 5e-12), ("glow_discharge", 2e-3, 1e-15, 1e-8)):
        rows = [[float(t) for t in ln.split()] for ln in (golden_dir / f"{case}_relative_error.log").read_text().splitlines()]
        max_error = [1, 1, 1]
        for i, (err, _, dt) in enumerate(rows):
            max_error[0] = err
            nxt = pid_step(dt, max_error, ttol, dt_min, dt_max)
            max_error[2], max_error[1] = max_error[1], max_error[0]
            if i + 1 < len(rows):
                assert nxt == rows[i + 1][2]


def test_hotpath_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fedm_b200.hotpath import HotPath
    with pytest.raises(RuntimeError):
        HotPath(0)
