"""The API mirror (fedm_b200/functions.py) against the reference's OWN package (fedm/functions.py, imported unmodified through the
shim): the same calls must produce forms with identical residual vectors (numpy backend, random P1 states).  Covers the branches
the example scripts do not exercise: the non-log balance equation, all equation types, every Boundary_flux type, the uncoupled
Source_term, Energy_Source_term sentinels, Max/Min, semi-implicit coefficients.  Needs /root/reference (build container only)."""
import sys
from pathlib import Path

import numpy as np
import pytest

REF = Path("/root/reference")
ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.skipif(not (REF / "fedm" / "functions.py").exists(), reason="reference checkout not available")


@pytest.fixture()
def both(monkeypatch):
    for p in (str(ROOT / "compat"), str(REF)):          # syspath_prepend: the last call ends up first -> `fedm` = reference
        monkeypatch.syspath_prepend(p)
    for name in [m for m in sys.modules if m == "dolfin" or m == "fedm" or m.startswith("fedm.")]:
        monkeypatch.delitem(sys.modules, name)
    import dolfin
    from oracle.ufl_backend import NumpyBackend
    dolfin.set_backend(NumpyBackend())
    dolfin.parameters["form_compiler"]["quadrature_degree"] = 4
    import fedm.functions as ref_f                       # the reference's package (REF precedes compat/ on sys.path)
    assert str(REF) in ref_f.__file__
    from fedm_b200 import functions as my_f
    yield dolfin, ref_f, my_f
    dolfin.set_backend(None)
    dolfin.parameters["form_compiler"]["quadrature_degree"] = None
    for name in [m for m in sys.modules if m == "dolfin" or m == "fedm" or m.startswith("fedm.")]:
        sys.modules.pop(name, None)


def _residual(df, form, u):
    prob = df.get_backend().compile(form, u, [], 4)
    return prob.residual(u.vector().store.a.copy())


def test_forms_of_mirror_and_reference_agree(both):
    df, R, M = both
    rng = np.random.default_rng(3)
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(1e-2, 2e-2), 5, 6, "crossed")
    markers = M.Marking_boundaries(mesh, [["line", 0.0, 0.0, 0.0, 1e-2], ["line", 2e-2, 2e-2, 0.0, 1e-2], ["line", 0.0, 2e-2, 0.0, 0.0], ["line", 0.0, 2e-2, 1e-2, 1e-2]])
    ds = df.Measure("ds", domain=mesh, subdomain_data=markers)
    normal = df.FacetNormal(mesh)
    P1 = df.FiniteElement("Lagrange", mesh.ufl_cell(), 1)
    ME = df.FunctionSpace(mesh, df.MixedElement([P1] * 3))
    V = df.FunctionSpace(mesh, P1)
    u_new, u_old, u_old1 = df.Function(ME), df.Function(ME), df.Function(ME)
    for f, s in ((u_new, 1.0), (u_old, 0.9), (u_old1, 0.8)):
        f.vector()[:] = s * rng.uniform(0.5, 1.5, ME.dim())
    v = df.TestFunctions(ME)
    coef = [df.Function(V) for _ in range(6)]
    for c in coef:
        c.vector()[:] = rng.uniform(0.5, 2.0, V.dim())
    dt = df.Expression("time_step", time_step=2e-3, degree=0)
    dt_old = df.Expression("time_step", time_step=3e-3, degree=0)
    r = df.Expression("x[0]", degree=1)
    E = -df.grad(u_new[2])

    def build(F):
        forms = []
        # balance equations: log and non-log, the three equation types
        for log_rep in (True, False):
            for k, eq in enumerate(("reaction", "diffusion-reaction", "drift-diffusion-reaction")):
                Gamma = F.Flux(-1.0, u_new[k % 2], coef[0], coef[1], E, grad_diffusion=bool(k % 2), logarithm_representation=log_rep)
                forms.append(F.weak_form_balance_equation(eq, dt, dt_old, df.dx, u_new[k % 2], u_old[k % 2], u_old1[k % 2], v[k % 2],
                                                          coef[2], Gamma, r, coef[3], log_representation=log_rep))
        forms.append(F.weak_form_Poisson_equation(df.dx, u_new[2], v[2], df.exp(u_new[0]) - df.exp(u_new[1]), r))
        # boundary terms
        ion_flux = F.Max(df.dot(F.Flux(1.0, u_new[0], coef[0], coef[1], E), normal), 0) + F.Min(df.dot(E, normal), 0)
        for bc_type in ("zero flux", "flux source", "Neumann"):
            for eq in ("reaction", "diffusion-reaction", "drift-diffusion-reaction"):
                for ptype in ("Heavy", "electrons"):
                    b = F.Boundary_flux(bc_type, eq, ptype, -1.0, coef[1], E, normal, u_new[1], 0.07, v[1], ds(2), r, 0.3 * coef[4], 0.25, ion_flux)
                    if not isinstance(b, (int, float)):
                        forms.append(b)
        # source terms
        p = np.array([[1, 0, 1], [0, 2, 0], [1, 1, 0]])
        l_ = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1]])
        g = np.array([[0, 1, 1], [1, 0, 0], [0, 2, 0]])
        k_si = F.semi_implicit_coefficients(["Umean", "const", "E/N"], coef[4] * df.exp(u_new[0]), coef[5], coef[:3], coef[3:])
        # number of species columns = N0 + the log-densities the variant keeps (fedm/functions.py:831-837)
        for coupling, approx, uu, ns in (("coupled", "LMEA", u_new, 2), ("coupled", "LFA", u_new, 3), ("uncoupled", "LMEA", [u_new[0], u_new[1], u_new[2]], 3)):
            for s_k, f_s in enumerate(F.Source_term(coupling, approx, p[:, :ns], l_[:, :ns], g[:, :ns], k_si, 2.5e1, uu)):
                if not isinstance(f_s, (int, float)):
                    forms.append(f_s * v[s_k % 3] * df.dx)
        losses = [11.5, 8e77 - 1e70, 1e100 - 1e90]                 # plain loss and the two sentinel ranges
        forms.append(F.Energy_Source_term("coupled", p[:, :2], l_[:, :2], g[:, :2], k_si, losses, u_new[0] / u_new[1], 2.5e1, u_new, Ei=15.7) * v[0] * df.dx)
        total = 0
        for f in forms:
            total = total + f
        return total, len(forms)

    form_ref, n_ref = build(R)
    form_my, n_my = build(M)
    assert n_ref == n_my and n_ref > 20
    b_ref, b_my = _residual(df, form_ref, u_new), _residual(df, form_my, u_new)
    scale = np.abs(b_ref).max()
    assert scale > 0 and np.abs(b_my - b_ref).max() <= 1e-13 * scale


def test_controllers_and_helpers_agree(both):
    df, R, M = both
    e = [2.3e-4, 1.9e-4, 2.8e-4]
    assert M.adaptive_timestep(1e-12, e, 1e-3, 1e-15, 1.0) == R.adaptive_timestep(1e-12, e, 1e-3, 1e-15, 1.0)
    assert M.adaptive_timestep_PI34(1e-12, e, 1e-3, 1e-15, 1.0) == R.adaptive_timestep_PI34(1e-12, e, 1e-3, 1e-15, 1.0)
    assert M.adaptive_timestep_H211b(1e-12, 2e-12, e, 1e-3, 1e-15, 1.0) == R.adaptive_timestep_H211b(1e-12, 2e-12, e, 1e-3, 1e-15, 1.0)
    for args in (("LFA", 3, ["a", "b", "c"], [1.0, 2.0, 3.0], [0.0, 1.0, -1.0]), ("LMEA", 2, ["a", "b"], [1.0, 2.0], [0.0, -1.0])):
        a = M.modify_approximation_vars(args[0], args[1], list(args[2]), list(args[3]), list(args[4]))
        b = R.modify_approximation_vars(args[0], args[1], list(args[2]), list(args[3]), list(args[4]))
        assert tuple(a) == tuple(b)
