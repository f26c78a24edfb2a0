"""The assembly kernels of row_kernel.cuh, run on the CPU: tests/csrc/cuda_emul stands in for the CUDA run time (device memory =
host memory, a kernel = a plain function executed once per (block, thread)), libfedmx's own fedmx_topology.cu builds the SELL /
fan structures, and a GENERATED model (the form compiler's output) is assembled row-wise and cell-wise on small meshes with
marked exterior facets, Dirichlet rows, a partly filled last scratch block and ghost vertices.  The two paths must agree to
rounding, the register / shared-memory accumulator variants bit for bit (tests/csrc/emul_assembly.cpp).  This is the CPU-side
guard for the kernels' index logic; the numbers themselves are checked on the device by the `-m gpu` tests."""
import shutil
import subprocess
from pathlib import Path

import pytest

from fedm_b200 import dolfin as df
from fedm_b200 import formcompiler as fc

from test_formcompiler_cpu import _tof_form, numpy_storage  # noqa: F401  (autouse fixture: numpy vector store for Functions)

ROOT = Path(__file__).resolve().parent.parent


def _emulate(tmp_path, form_cu):
    gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    exe = tmp_path / "emul"
    cmd = [gxx, "-std=c++17", "-O1", "-x", "c++", "-I", str(ROOT / "tests" / "csrc" / "cuda_emul"), "-I", str(ROOT / "fedm_b200" / "csrc"),
           f'-DFORM="{form_cu}"', "-o", str(exe), str(ROOT / "tests" / "csrc" / "emul_assembly.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "bad 0" in r.stdout.splitlines()[-1], r.stdout[-3000:]
    return r.stdout


def test_cellwise_and_rowwise_kernels_agree_under_emulation_scalar_form(tmp_path):
    df.parameters["form_compiler"]["quadrature_degree"] = None
    form, u_new = _tof_form(df.RectangleMesh(df.Point(0, 0), df.Point(1, 1), 2, 2))
    cf = fc.CompiledForm(form, u_new, 4)
    assert cf.cellwise
    src = tmp_path / "tof_form.cu"
    src.write_text(cf.source)
    out = _emulate(tmp_path, src)
    assert out.count("cellwise 1") == 3


def test_cellwise_and_rowwise_kernels_agree_under_emulation_cached_systems(tmp_path):
    """The multi-component plug-ins built by __graft_entry__.build() (streamer: n_eq = 3, cell-wise with facet integrals; glow:
    n_eq = 5, row-wise only -- its accumulator variants are still compared)."""
    forms = [f for f in sorted(fc.CACHE.glob("form_*.cu")) if "NEQ = 1;" not in f.read_text() and "emul" not in f.name
             and "CELLS_JAC" in f.read_text()]
    if not forms:
        pytest.skip("no cached multi-component plug-ins (run __graft_entry__.build())")
    seen_cellwise = False
    for f in forms:
        out = _emulate(tmp_path, f)
        seen_cellwise = seen_cellwise or "cellwise 1" in out
    assert seen_cellwise or all("NEQ = 3;" not in f.read_text() for f in forms)
