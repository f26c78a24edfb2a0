"""CPU tests of the reference-named API layer (no GPU): the reference's own unit tests for
modify_approximation_vars (tests/unit_tests/functions/test_modify_approximation_vars.py:6-54, restated), the
controllers, the readers on the reference's input decks, and adaptive_solver's accept/reject logic driven by a
scripted fake solver (log format, retry factors, SystemExit)."""
import numpy as np
import pytest

from fedm_b200 import file_io, functions as F


def test_modify_approximation_vars_lfa_drops_first_species():
    ns, ne, sp, m, c = F.modify_approximation_vars("LFA", 3, ["a", "b", "c"], [1.0, 2.0, 3.0], [0.0, 1.0, -1.0])
    assert (ns, ne, sp, m, c) == (2, 3, ["b", "c"], [2.0, 3.0], [1.0, -1.0])


def test_modify_approximation_vars_lmea_keeps_all():
    ns, ne, sp, m, c = F.modify_approximation_vars("LMEA", 3, ["a", "b", "c"], [1.0, 2.0, 3.0], [0.0, 1.0, -1.0])
    assert (ns, ne, sp) == (3, 4, ["a", "b", "c"]) and m == [1.0, 2.0, 3.0] and c == [0.0, 1.0, -1.0]


def test_modify_approximation_vars_bad_name():
    with pytest.raises(ValueError) as exc:
        F.modify_approximation_vars("hello world", 3, ["a"], [1.0], [0.0])
    assert "hello world" in str(exc.value)


def test_controllers_match_python_formulas():
    e = [2.3e-4, 1.9e-4, 2.8e-4]
    ref = 1e-12 * ((e[1] / e[0]) ** 0.075 * (1e-3 / e[0]) ** 0.175 * (e[1] ** 2 / (e[0] * e[2])) ** 0.01)
    assert F.adaptive_timestep(1e-12, e, 1e-3, 1e-15, 1.0) == ref
    ref = 1e-12 * ((0.8 * 1e-3 / e[0]) ** (0.3 / 3) * (0.8 * e[1] / e[0]) ** (0.4 / 3))
    assert F.adaptive_timestep_PI34(1e-12, e, 1e-3, 1e-15, 1.0) == ref
    ref = 1e-12 * ((0.8 * 1e-3 / e[0]) ** (1 / 12) * (0.8 * 1e-3 / e[1]) ** (1 / 12) * (1e-12 / 2e-12) ** (-1 / 4))
    assert F.adaptive_timestep_H211b(1e-12, 2e-12, e, 1e-3, 1e-15, 1.0) == ref


def test_readers_on_reference_decks(golden_dir):
    file_io.files.file_input = golden_dir / "glow_discharge_input"
    path = file_io.files.file_input / "4_particles"
    n, names, props, stems = file_io.read_speclist(path)
    assert (n, names, stems) == (4, ["Ar[1p0]", "Ar[*]", "Ar[+]", "e"], ["Ar_1p0", "Ar_star", "Ar_plus", "electrons"])
    M, Z = file_io.read_particle_properties(props, "4_particles")
    assert Z == [0.0, 0.0, 1.0, -1.0] and M[3] == 9.10938215e-31
    P, L, G = file_io.reaction_matrices(path, names)
    assert P.tolist() == [[1, 0, 0, 1], [1, 0, 0, 1], [0, 1, 0, 1], [0, 1, 0, 1], [0, 2, 0, 0], [0, 1, 0, 0], [1, 0, 0, 1]]
    assert (G - L).tolist() == [[-1, 1, 0, 0], [-1, 0, 1, 1], [1, -1, 0, 0], [0, -1, 1, 1], [1, -2, 1, 1], [0, -1, 0, 0], [0, 0, 0, 0]]
    assert file_io.read_energy_loss(path) == [11.55, 15.76, -11.55, 4.21, -7.34, 0.0, 1.0]
    kf = file_io.rate_coefficient_file_names(path)
    deps = file_io.read_dependences(kf)
    assert deps == ["Umean"] * 4 + ["const", "const", "Umean"]
    kx, ky = file_io.read_rate_coefficients(kf, deps)
    assert len(kx[0]) == 353 and ky[4] == 0.81e-15 and ky[5] == 1e6
    mx, my, md = file_io.read_transport_coefficients(stems, "mobility", "4_particles")
    assert md == [0, 0, "E/N", "Umean"] and my[0] == 0 and len(mx[2]) == 306
    dx, dy, dd = file_io.read_transport_coefficients(stems, "Diffusion", "4_particles")
    assert dd == ["const", "const", "ESR", "Umean"] and dy[1] == 1.738e20
    file_io.files.file_input = golden_dir / "streamer_discharge_input"
    sx, sy, sd = file_io.read_transport_coefficients(["ions", "e"], "mobility", "benchmark_model")
    assert sd == ["const.", "fun:E"] and sy[1] == "2.3987*E_m**(-0.26)"
    with pytest.raises(RuntimeError):
        file_io.files.file_input = golden_dir / "does_not_exist"


class _FakeFn:
    """Scalar function over the shim's GenericVector with the oracle's numpy store (no mesh needed)."""

    def __init__(self, values):
        from fedm_b200.dolfin import GenericVector
        from oracle.ufl_backend import NumpyStore
        self._v = GenericVector(NumpyStore(len(values)), 1, self)
        self._v[:] = values
        self.assigned = 0

    def vector(self):
        return self._v

    def assign(self, other):
        self.assigned += 1
        self._v[:] = other.vector()[:]


class _FakeSolver:
    """Scripted solver: attempt k fails (raises like dolfin's SNES) or writes ``1 + delta_k`` into the solution."""

    def __init__(self, fail_on, deltas):
        self.calls, self.fail_on, self.deltas = 0, set(fail_on), list(deltas)

    def solve(self, problem, x):
        self.calls += 1
        if self.calls in self.fail_on:
            raise RuntimeError("Newton solver did not converge")
        x[:] = 1.0 + self.deltas.pop(0)
        return 3, True


class _Assigner:
    def assign(self, receiving, assigning):
        for f in receiving:
            f.vector()[:] = assigning.vector()[:]


def _estimate(delta, n=10):
    new, old = np.full(n, 1.0 + delta), np.ones(n)
    return float(np.linalg.norm(new - old + 3.0e-16) / np.linalg.norm(old + 3.0e-16))


def test_adaptive_solver_accept_reject_and_log(tmp_path):
    from fedm_b200.dolfin import Expression
    e1, e2 = _estimate(5e-3), _estimate(4e-4)    # first attempt too inaccurate, third attempt accepted
    solver = _FakeSolver(fail_on={2}, deltas=[5e-3, 4e-4])   # second attempt: Newton failure
    dt, dt_old = Expression("time_step", time_step=4e-12), Expression("time_step", time_step=1e30)
    ones = np.ones(10)
    u_new, u_old = _FakeFn(ones), _FakeFn(ones)
    new, old = [_FakeFn(ones)], [_FakeFn(ones)]
    error, max_error = [0.0, 0.0], [1, 1, 1]
    log = tmp_path / "relative error.log"

    class P:
        t = 0.0
    t = F.adaptive_solver(solver, P(), 1e-9, dt, dt_old, u_new, u_old, new, old, _Assigner(), error, log, max_error, 1e-3, 1e-15,
                          time_dependent_arguments=[], approximation="LMEA")
    dt1 = 4e-12 * (0.5 * 1e-3 / e1)              # error >= ttol: dt *= 0.5*ttol/max_error
    dt2 = dt1 * 0.5                              # exception: dt *= 0.5
    assert dt.time_step == dt2 and t == 1e-9 + dt2 and solver.calls == 3 and u_new.assigned == 2
    rows = log.read_text().splitlines(keepends=True)
    assert rows == [f"{e1:<23}  {1e30:<23}  {4e-12:<23}\n", f"{e2:<23}  {1e30:<23}  {dt2:<23}\n"]
    assert max_error[0] == e2
    # below dt_min the program terminates, as in the reference
    with pytest.raises(SystemExit):
        F.adaptive_solver(_FakeSolver((), [1.0] * 50), P(), 0.0, Expression("time_step", time_step=1e-14), dt_old, u_new, u_old,
                          new, old, _Assigner(), [0.0], log, [1, 1, 1], 1e-3, 1e-15, approximation="LMEA")


@pytest.mark.skipif(not __import__("os").environ.get("FEDM_B200_SLOW"), reason="4 minutes; set FEDM_B200_SLOW=1")
def test_glow_example_on_numpy_backend_matches_fenics_golden(tmp_path, golden_dir, monkeypatch):
    """examples/glow_discharge.py (API mirror + shim) on the oracle's numpy backend against the reference's golden HDF5
    fields -- the CPU check of the script that tests/test_gpu_shim.py runs through the CUDA backend."""
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    for p in (str(root / "compat"), str(root / "examples")):
        monkeypatch.syspath_prepend(p)
    import dolfin
    from oracle.ufl_backend import NumpyBackend
    dolfin.set_backend(NumpyBackend())
    try:
        import glow_discharge as gd
        gd.main(input_dir=golden_dir / "glow_discharge_input", output_dir=tmp_path)
    finally:
        dolfin.set_backend(None)
        dolfin.parameters["form_compiler"]["quadrature_degree"] = None
        sys.modules.pop("glow_discharge", None)
    # the checkpoints are real HDF5 in dolfin's layout and dof numbering: read them the way the reference's own test does
    # (testing_utils.read_h5) and compare entry by entry with the FEniCS goldens -- no re-ordering by the golden's dof map
    from fedm_b200.hdf5_lite import H5Reader, read_h5
    g = np.load(golden_dir / "glow_h5.npz")
    for name in ("electrons", "Ar_plus", "Ar_star"):
        path = tmp_path / "number density" / name / f"{name}.h5"
        vecs = read_h5(path, name)
        assert len(vecs) == 2 and vecs[0].shape == (20201, 1)
        ref = g[name + "_1"]
        assert np.abs((vecs[1].ravel() - ref) / ref).max() <= 1e-11, name
        assert np.array_equal(vecs[0].ravel(), g[name + "_0"])
        cp = H5Reader(path).tree()[name][name + "_1"]
        assert np.array_equal(cp["cell_dofs"].ravel(), g["cell_dofs"])          # dolfin's dof map, bit for bit
        assert np.array_equal(cp["mesh"]["topology"], g["topology"]) and np.array_equal(cp["mesh"]["geometry"], g["geometry"])
        assert (path.with_suffix(".xdmf")).read_text().count("<Grid Name=") == 3


def test_segregated_helpers_on_numpy_backend():
    """Normal_vector / Poisson_solver / BoundaryGradient of the API mirror on the oracle's numpy backend: analytic checks
    (unit normals on straight edges, nodally exact 1-D Poisson profile, consistent boundary flux)."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    from helpers_cases import check_helpers, run_helpers
    from fedm_b200 import dolfin as df
    from oracle.ufl_backend import NumpyBackend
    df.set_backend(NumpyBackend())
    df.parameters["form_compiler"]["quadrature_degree"] = None
    try:
        out = run_helpers(df, F)
    finally:
        df.set_backend(None)
    check_helpers(out)


def test_marking_circle_boundaries():
    """'circle' entries of Marking_boundaries (fedm/functions.py:48-70, 95-98): facets whose vertices and midpoint lie on the
    arc (within the reference's tolerance 1e-8 on the squared distance) and outside the gap are marked; works on the same
    marker function that feeds the device path's `ds(i)` measures."""
    from fedm_b200 import dolfin as df
    from fedm_b200.mesh import TriMesh, rectangle
    m = rectangle((0.0, 0.0), (4e-4, 4e-4), 40, 40, "right")
    R, cz = 1e-3, -1e-3 + 0.0                                   # needle-like electrode: arc through (r=0, z=0), centre below
    xy = m.coords.copy()
    bottom = np.abs(xy[:, 1]) < 1e-15
    # bend the bottom edge onto the circle of radius R around (r=0, z=cz): z = cz + sqrt(R^2 - r^2) <= 0
    xy[bottom, 1] = cz + np.sqrt(R ** 2 - xy[bottom, 0] ** 2)
    mesh = df.Mesh(TriMesh(xy, m.cells))
    markers = F.Marking_boundaries(mesh, [["circle", cz, 0.0, R], ["line", 4e-4, 4e-4, 0.0, 4e-4]], gap_length=4e-4)
    vals = np.asarray(markers.array())
    fv = markers.facets if hasattr(markers, "facets") else mesh.tri.exterior_facets()
    on_arc = [i for i, (a, b) in enumerate(fv) if bottom[a] and bottom[b]]
    assert len(on_arc) == 40 and all(vals[i] == 1 for i in on_arc)
    assert (vals == 2).sum() == 40 and (vals == 1).sum() == 40
