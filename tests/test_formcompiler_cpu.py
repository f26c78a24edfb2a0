"""CPU tests of the script-facing tool chain that replaces FFC/UFL/dijitso for the hot path (no GPU needed):
the C++ Expression-string parser, UFL-lite (degree estimation as UFL's SumDegreeEstimator, lhs/rhs, action, test-function
split) and the form compiler's code generation + nvcc cross-compilation of a plug-in."""
import ctypes
import math

import numpy as np
import pytest

from fedm_b200 import cexpr
from fedm_b200 import dolfin as df
from fedm_b200 import functions as F
from fedm_b200 import ufl as U


def test_expression_strings_of_the_reference_scripts_evaluate_like_numpy():
    x0, x1 = np.linspace(0.0, 2.5e-4, 7), np.linspace(0.0, 5e-4, 7)
    p = dict(D=0.12, w=1.7e5, alpha=5009.51, t=2.5e-9, pi=math.pi)
    # fedm_tof.py:116 (analytic swarm), fedm_gd.py:272 (powered electrode), fedm_streamer.py:177 (seed)
    cases = {
        "std::log(exp(-(pow(x[1]-w*t, 2)+pow(x[0], 2))/(4.0*D*t)+alpha*w*t)/pow(4*D*t*pi,1.5))":
            np.log(np.exp(-((x1 - p["w"] * p["t"]) ** 2 + x0 ** 2) / (4.0 * p["D"] * p["t"]) + p["alpha"] * p["w"] * p["t"]) / (4 * p["D"] * p["t"] * math.pi) ** 1.5),
        "U0*(1-exp(-t/1e-9))": -250.0 * (1 - math.exp(-2.5)) * np.ones(7),
        "std::log(1e13 + 5e18*exp(-(pow(x[0], 2) + pow(x[1]-1e-2, 2))/pow(0.4e-3, 2)))":
            np.log(1e13 + 5e18 * np.exp(-(x0 ** 2 + (x1 - 1e-2) ** 2) / 0.4e-3 ** 2)),
        "-x[0] + 2.0*-3.0 + DOLFIN_EPS": -x0 - 6.0 + 3.0e-16,
    }
    for text, ref in cases.items():
        got = cexpr.evaluate(cexpr.parse(text), (x0, x1), dict(p, U0=-250.0))
        assert np.allclose(np.broadcast_to(got, ref.shape), ref, rtol=1e-15, atol=0.0), text
    with pytest.raises(SyntaxError):
        cexpr.parse("pow(x[0], )")


@pytest.fixture(autouse=True)
def numpy_storage():
    """Functions need a vector store: on the CPU the oracle's numpy backend provides it (the product default, CUDA, refuses to
    start without a GPU)."""
    from oracle.ufl_backend import NumpyBackend
    df.set_backend(NumpyBackend())
    df.parameters["form_compiler"]["quadrature_degree"] = None
    yield
    df.set_backend(None)
    df.parameters["form_compiler"]["quadrature_degree"] = None


def _tof_form(mesh):
    V = df.FunctionSpace(mesh, "P", 1)
    W = df.VectorFunctionSpace(mesh, "P", 1)
    u, v = df.TrialFunction(V), df.TestFunction(V)
    u_old, u_old1, u_new = df.Function(V), df.Function(V), df.Function(V)
    dt = df.Expression("time_step", time_step=1e-12, degree=0)
    dt_old = df.Expression("time_step", time_step=1e30, degree=0)
    r = df.Expression("x[0]", degree=1)
    w = df.interpolate(df.Constant(("0", 1.7e5)), W)
    D = df.interpolate(df.Constant(0.12), V)
    f = df.Expression("exp(-(pow(x[1]-w*t, 2)+pow(x[0], 2))/(4.0*D*t))", D=0.12, w=1.7e5, t=2.5e-9, degree=2)
    Gamma = -df.grad(D * df.exp(u)) + w * df.exp(u)
    form = F.weak_form_balance_equation_log_representation("drift-diffusion-reaction", dt, dt_old, df.dx, u, u_old, u_old1, v, f, Gamma, r)
    return df.action(form, u_new), u_new


def test_degree_estimation_matches_ufl_on_the_time_of_flight_form():
    """FFC integrates the ToF form with the degree UFL estimates (no global quadrature degree is set): 8, i.e. the 25-point
    collapsed Gauss-Jacobi rule that the golden VTU pins."""
    df.parameters["form_compiler"]["quadrature_degree"] = None
    form, _ = _tof_form(df.RectangleMesh(df.Point(0, 0), df.Point(1, 1), 2, 2))
    total = U.Const(0.0)
    for it in form.integrals:
        total = total + it.integrand
    assert U.estimate_degree(total) == 8


def test_lhs_rhs_and_action_on_the_poisson_form():
    mesh = df.RectangleMesh(df.Point(0, 0), df.Point(1, 1), 2, 2)
    V = df.FunctionSpace(mesh, "P", 1)
    u, v = df.TrialFunction(V), df.TestFunction(V)
    form = F.weak_form_Poisson_equation(df.dx, u, v, df.Constant(3.0))
    a, L = df.lhs(form), df.rhs(form)
    assert all(U._has_trial(i.integrand) for i in a.integrals if not isinstance(i.integrand, U.Const))
    assert not any(U._has_trial(i.integrand) for i in L.integrals)
    phi = df.Function(V)
    res = df.action(form, phi)
    assert not any(U._has_trial(i.integrand) for i in res.integrals)
    split = U.test_function_split(sum((i.integrand for i in res.integrals), U.Const(0.0)), 1)
    s, fx, fy = split[0]
    assert not (isinstance(fx, U.Const) and fx.value == 0.0) and not (isinstance(fy, U.Const) and fy.value == 0.0)   # grad(u).grad(v)
    assert not (isinstance(s, U.Const) and s.value == 0.0)                                                          # -f v


def test_form_compiler_generates_and_cross_compiles_a_plugin(tmp_path, monkeypatch):
    """UFL-lite -> CUDA source -> nvcc (sm_100a) -> .so exporting the plug-in entry points libfedmx dlopen()s; the cache key
    covers the generated source and the library headers."""
    from fedm_b200 import formcompiler as fc
    df.parameters["form_compiler"]["quadrature_degree"] = None
    form, u_new = _tof_form(df.RectangleMesh(df.Point(0, 0), df.Point(1, 1), 2, 2))
    cf = fc.CompiledForm(form, u_new, None)
    assert cf.n_eq == 1 and [g[:3] for g in cf.groups] == [("dx", None, 8)]
    assert "QC0[25][4]" in cf.source and "struct JitModel" in cf.source and "dexp(" in cf.source
    assert cf.key == fc.CompiledForm(form, u_new, None).key                       # deterministic
    assert cf.key != fc.CompiledForm(form, u_new, 2).key                          # another quadrature rule = another kernel
    monkeypatch.setattr(fc, "CACHE", tmp_path)
    so = cf.build()
    lib = ctypes.CDLL(str(so))
    assert lib.fedmx_plugin_neq() == 1 and lib.fedmx_plugin_nf() == len(cf.fields)
    assert hasattr(lib, "fedmx_plugin_launch") and lib.fedmx_plugin_args_size() > 0
    assert cf.build() == so                                                       # second call hits the cache


def _host_run(tmp_path, source, *defs):
    """g++-compile a host test program against the library's dual-number headers (CUDA qualifiers defined away)."""
    import shutil
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    exe = tmp_path / "host_test"
    cmd = [gxx, "-std=c++17", "-O1", "-I", str(root / "tools" / "jitcheck" / "stub"), "-I", str(root / "fedm_b200" / "csrc"),
           *defs, "-o", str(exe), str(root / "tests" / "csrc" / source)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    return r.stdout


def test_component_sparse_duals_match_finite_differences_and_dense_duals(tmp_path):
    """csrc/cdual.cuh (what generated kernels compute with): every operator against central differences, bit-equal with the
    dense Dual<9> on the live lanes, and lane counts follow the component masks."""
    assert "bad 0" in _host_run(tmp_path, "cdual_fd.cpp")


def test_generated_rows_have_the_exact_jacobian_of_their_residual(tmp_path):
    """J = derivative(F, u) (fedm_streamer.py:296-297): the generated row<true> of the time-of-flight form and of every cached
    multi-component form (streamer n_eq = 3, glow n_eq = 5) against central differences of row<false>, compiled for the host."""
    from fedm_b200 import formcompiler as fc
    df.parameters["form_compiler"]["quadrature_degree"] = None
    form, u_new = _tof_form(df.RectangleMesh(df.Point(0, 0), df.Point(1, 1), 2, 2))
    cf = fc.CompiledForm(form, u_new, 4)
    src = tmp_path / "tof_form.cu"
    src.write_text(cf.source)
    sources = [src] + [f for f in sorted(fc.CACHE.glob("form_*.cu")) if "NEQ = 1;" not in f.read_text()
                       and '#include "cdual.cuh"' in f.read_text()]
    for f in sources:
        out = _host_run(tmp_path, "jit_row_fd.cpp", f'-DFORM="{f}"')
        assert "bad 0" in out, (f.name, out[-2000:])


def test_cellwise_path_is_generated_only_where_it_pays():
    """Cell-wise assembly (row_kernel.cuh) is emitted for expensive integrands (the ToF form: 25 quadrature points with exp) and
    not for a Poisson residual; the plug-in reports its scratch need through fedmx_plugin_cell_doubles."""
    from fedm_b200 import formcompiler as fc
    df.parameters["form_compiler"]["quadrature_degree"] = None
    mesh = df.RectangleMesh(df.Point(0, 0), df.Point(1, 1), 2, 2)
    form, u_new = _tof_form(mesh)
    tof = fc.CompiledForm(form, u_new, None)
    assert tof.cellwise and "static constexpr bool CELLS = true;" in tof.source
    assert "REC_JAC = 4, REC_RES = 2" in tof.source                      # n_eq + 3 n_eq^2 = 4 doubles per (cell, test vertex)
    assert "fedmx_plugin_cell_doubles" in tof.source and "FX = FX + " in tof.source
    V = df.FunctionSpace(mesh, "P", 1)
    u, v = df.TrialFunction(V), df.TestFunction(V)
    phi = df.Function(V)
    poisson = fc.CompiledForm(df.action(F.weak_form_Poisson_equation(df.dx, u, v, df.Constant(3.0)), phi), phi, None)
    assert not poisson.cellwise and "static constexpr bool CELLS = false;" in poisson.source
    # dependency masks: every sub-expression of the scalar ToF form can only depend on component 0
    for it in form.integrals:
        assert tof.deps(it.integrand) <= frozenset([0])


def _write_dolfin_xml(path, coords, cells):
    lines = ['<?xml version="1.0"?>', '<dolfin xmlns:dolfin="http://fenicsproject.org">', '  <mesh celltype="triangle" dim="2">',
             f'    <vertices size="{len(coords)}">']
    lines += [f'      <vertex index="{i}" x="{x!r}" y="{y!r}" />' for i, (x, y) in enumerate(coords.tolist())]
    lines += ["    </vertices>", f'    <cells size="{len(cells)}">']
    lines += [f'      <triangle index="{i}" v0="{a}" v1="{b}" v2="{c}" />' for i, (a, b, c) in enumerate(cells.tolist())]
    lines += ["    </cells>", "  </mesh>", "</dolfin>"]
    path.write_text("\n".join(lines))


def test_dolfin_xml_mesh_reader_and_circle_boundaries(tmp_path):
    """Mesh('mesh.xml') (fedm-streamer.py:116) and Marking_boundaries with a 'circle' entry (fedm/functions.py:86-124)."""
    from fedm_b200.mesh import rectangle
    tri = rectangle((0.0, 0.0), (1.0, 1.0), 6, 6, "crossed")
    cells = tri.cells[:, ::-1].copy()                         # unsorted vertex order in the file: dolfin sorts on read
    _write_dolfin_xml(tmp_path / "mesh.xml", tri.coords, cells)
    mesh = df.Mesh(str(tmp_path / "mesh.xml"))
    assert mesh.num_vertices() == tri.num_vertices and mesh.num_cells() == tri.num_cells
    assert np.array_equal(mesh.coordinates(), tri.coords) and np.array_equal(mesh.cells(), np.sort(tri.cells, axis=1))
    # bottom edge z = 0 as a line, and the unit circle around the origin clipped to z <= 0 (touches the boundary at (1, 0) only)
    markers = F.Marking_boundaries(mesh, [["line", 0.0, 0.0, 0.0, 1.0], ["circle", 0.0, 0.0, 1.0]], gap_length=1.0)
    vals = markers.array()
    xy = mesh.coordinates()
    mids = 0.5 * (xy[markers.facets[:, 0]] + xy[markers.facets[:, 1]])
    assert set(np.unique(vals)) <= {0, 1, 2}
    assert (vals[np.abs(mids[:, 1]) < 1e-14] == 1).all() and (vals[np.abs(mids[:, 1]) > 1e-14] == 0).all()
    with pytest.raises(ValueError):
        F.Marking_boundaries(mesh, [["ellipse", 0, 0, 1]])


def test_streamer_script_runs_on_an_xml_mesh_like_on_the_generated_one(tmp_path, golden_dir, monkeypatch):
    """The reference's own work flow -- a mesh.xml on disk -- through examples/streamer_discharge.py: identical to the run on
    the equivalent generated mesh (numpy backend; the GPU run of the same script is tests/test_gpu_shim.py)."""
    import sys
    from pathlib import Path
    from fedm_b200.mesh import rectangle
    root = Path(__file__).resolve().parent.parent
    for p in (str(root / "compat"), str(root / "examples")):
        monkeypatch.syspath_prepend(p)
    import dolfin
    from oracle.ufl_backend import NumpyBackend
    dolfin.set_backend(NumpyBackend())
    import streamer_discharge as sd
    n, box = 8, 0.0125
    tri = rectangle((0.0, 0.0), (box, box), n, n, "right")
    _write_dolfin_xml(tmp_path / "mesh.xml", tri.coords, tri.cells)
    try:
        kw = dict(input_dir=golden_dir / "streamer_discharge_input", linear_solver="mumps", max_steps=2)
        u1, _ = sd.main(n=n, output_dir=tmp_path / "a", **kw)
        u2, _ = sd.main(n=n, output_dir=tmp_path / "b", mesh_file=tmp_path / "mesh.xml", **kw)
    finally:
        sys.modules.pop("streamer_discharge", None)
    assert np.array_equal(u1.vector()[:], u2.vector()[:])
    assert (tmp_path / "a" / "relative error.log").read_text() == (tmp_path / "b" / "relative error.log").read_text()
