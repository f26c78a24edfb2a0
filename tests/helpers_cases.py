"""Small problems for the segregated helpers of the API mirror (Normal_vector, Poisson_solver, BoundaryGradient,
fedm/functions.py:1133-1208), shared by the CPU test (numpy backend) and the GPU test (CUDA backend)."""
import numpy as np


def run_helpers(df, F, n=16):
    """Returns a dict of numpy arrays computed through whichever backend is installed in ``df``."""
    out = {}
    mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(1.0, 2.0), n, n)
    out["coords"] = mesh.coordinates().copy()
    # ---- Normal_vector
    nvec = F.Normal_vector(mesh)
    out["normal"] = nvec.vector().get_local().reshape(-1, 2)
    # ---- Poisson_solver: -div grad u = 6 with u = 0 at z = 0 and u = 1 at z = 2 (Cartesian measure: r = 1/(2 pi))
    V = df.FunctionSpace(mesh, "P", 1)
    u, v = df.TrialFunction(V), df.TestFunction(V)
    f = df.Constant(6.0)
    Fp = F.weak_form_Poisson_equation(df.dx, u, v, f)
    a, L = df.lhs(Fp), df.rhs(Fp)
    bcs = [df.DirichletBC(V, df.Constant(0.0), lambda x, on: on and df.near(x[1], 0.0)),
           df.DirichletBC(V, df.Constant(1.0), lambda x, on: on and df.near(x[1], 2.0))]
    A = df.assemble(a)
    for bc in bcs:
        bc.apply(A)
    phi = df.Function(V)
    F.Poisson_solver(A, L, None, bcs, phi)
    out["phi"] = phi.vector().get_local()
    # ---- BoundaryGradient: normal flux of -eps grad(phi) through the electrode z = 0
    markers = F.Marking_boundaries(mesh, [["line", 0.0, 0.0, 0.0, 1.0], ["line", 2.0, 2.0, 0.0, 1.0]])
    ds = df.Measure("ds", domain=mesh, subdomain_data=markers)

    class Rest(df.SubDomain):
        def inside(self, x, on_boundary):
            return x[1] > 1e-12

    En = F.BoundaryGradient(phi, Rest, 6.0 * 1.0, [ds(1), ds(1)], epsilon=1.0)
    out["flux"] = En.vector().get_local()
    return out


def check_helpers(out, n=16):
    xy = out["coords"]
    x, z = xy[:, 0], xy[:, 1]
    nv = out["normal"]
    interior = (x > 1e-9) & (x < 1 - 1e-9) & (z > 1e-9) & (z < 2 - 1e-9)
    assert np.abs(nv[interior]).max() == 0.0                                  # ident_zeros rows
    mid_bottom = (np.abs(z) < 1e-12) & (np.abs(x - 0.5) < 0.07)
    mid_right = (np.abs(x - 1.0) < 1e-12) & (np.abs(z - 1.0) < 0.13)
    assert np.abs(nv[mid_bottom] - [0.0, -1.0]).max() < 1e-3
    assert np.abs(nv[mid_right] - [1.0, 0.0]).max() < 1e-3
    # exact solution of -u'' = 6 on (0, 2), u(0) = 0, u(2) = 1:  u = -3 z^2 + 6.5 z  (P1 is nodally exact in 1-D structure)
    exact = -3.0 * z ** 2 + 6.5 * z
    assert np.abs(out["phi"] - exact).max() < 1e-9
    # the reference's formula adds the consistent (residual-based) flux phi_z(0) = 6.5 and the direct P1 gradient
    # phi(h)/h of the boundary cells; nodes inside zeroDomain are pinned to zero
    fl = out["flux"]
    assert np.abs(fl[z > 1e-12]).max() == 0.0
    h = 2.0 / n
    expected = 6.5 + (-3.0 * h * h + 6.5 * h) / h
    mid = (np.abs(z) < 1e-12) & (np.abs(x - 0.5) < 0.07)
    assert np.abs(fl[mid] - expected).max() < 2e-3 * expected, (fl[mid], expected)
    return fl
