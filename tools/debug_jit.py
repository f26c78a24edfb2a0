"""GPU-box diagnostic: ToF forms through the CUDA backend (JIT row kernel) vs the oracle's numpy backend --
residual, Jacobian, one Newton solve and project(), stage by stage."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "compat")]
import dolfin as df                      # noqa: E402
from dolfin import *                     # noqa: E402,F401,F403
from fedm.functions import *             # noqa: E402,F401,F403


def build(nx=40, ny=40, width=2.5e-4, height=5e-4, t0=2.5e-9, dt_init=1e-12):
    wez, De, alpha_e = 1.7e5, 0.12, 5009.51
    dt = Expression("time_step", time_step=dt_init, degree=0)
    dt_old = Expression("time_step", time_step=1e30, degree=0)
    r = Expression("x[0]", degree=1)
    mesh = RectangleMesh(Point(0, 0), Point(width, height), nx, ny)
    V = FunctionSpace(mesh, "P", 1)
    W = VectorFunctionSpace(mesh, "P", 1)
    u, v = TrialFunction(V), TestFunction(V)
    u_old, u_old1, u_new = Function(V), Function(V), Function(V)
    swarm = "exp(-(pow(x[1]-w*t, 2)+pow(x[0], 2))/(4.0*D*t)+alpha*w*t)"
    kw = dict(D=De, w=wez, alpha=alpha_e, t=t0, pi=pi)
    u_analytical = Expression(f"std::log({swarm}/pow(4*D*t*pi,1.5))", degree=3, **kw)
    u_old.assign(interpolate(u_analytical, V))
    u_old1.assign(interpolate(u_analytical, V))
    w = interpolate(Constant(("0", wez)), W)
    D = interpolate(Constant(De), V)
    Gamma = -grad(D * exp(u)) + w * exp(u)
    f = Expression(f"{swarm}*(w*alpha)/(8*pow(pi,1.5)*pow(D*t, 1.5))", degree=2, **kw)
    F = weak_form_balance_equation_log_representation("drift-diffusion-reaction", dt, dt_old, dx, u, u_old, u_old1, v, f, Gamma, r)
    u_new.assign(interpolate(Expression(f"std::log({swarm}/pow(4.0*D*t*pi,1.5) + DOLFIN_EPS)", degree=2, **kw), V))
    F = action(F, u_new)
    return dict(F=F, u_new=u_new, u_old=u_old, u_old1=u_old1, f=f, ua=u_analytical, dt=dt, dt_old=dt_old, V=V, t0=t0)


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def main():
    from oracle.ufl_backend import NumpyBackend
    df.set_backend(NumpyBackend())
    P = build()
    pn = df.get_backend().compile(P["F"], P["u_new"], [], None)
    df.set_backend(None)
    Q = build()
    pc = df.get_backend().compile(Q["F"], Q["u_new"], [], None)
    print("groups numpy", [(g[0], g[1], g[3]) for g in pn.groups], "cuda", [(g[0], g[1], g[2]) for g in pc.form.groups])
    for step in range(3):
        t = P["t0"] + (step + 1) * 1e-12
        for S in (P, Q):
            S["u_old1"].assign(S["u_old"]); S["u_old"].assign(S["u_new"])
            S["f"].t = t; S["ua"].t = t
        x = P["u_new"].vector().get_local()
        print("state diff", rel(Q["u_new"].vector().get_local(), x))
        rn, rc = pn.residual(x), pc.residual(Q["u_new"])
        print(f"step {step}: residual rel diff {rel(rc, rn):.3e}  |r|={np.abs(rn).max():.3e}")
        Jn = pn.jacobian(x).toarray()
        rowptr, col, val = pc.jacobian_bsr(Q["u_new"])
        import scipy.sparse as sp
        Jc = sp.csr_matrix((np.asarray(val).ravel(), col, rowptr), shape=Jn.shape).toarray()
        print(f"        jacobian rel diff {rel(Jc, Jn):.3e}")
        itn = pn.solve_newton(P["u_new"], rtol=1e-10)
        itc = pc.solve_newton(Q["u_new"], rtol=1e-10)
        print("        newton", itn, itc, "solution diff", rel(Q["u_new"].vector().get_local(), P["u_new"].vector().get_local()))
        if step > 0:
            for S in (P, Q):
                S["dt_old"].time_step = S["dt"].time_step
    df.set_backend(NumpyBackend())
    a = df.project(df.exp(P["u_new"]), P["V"]).vector().get_local()
    a2 = df.project(df.exp(P["ua"]), P["V"]).vector().get_local()
    df.set_backend(None)
    b = df.project(df.exp(Q["u_new"]), Q["V"]).vector().get_local()
    b2 = df.project(df.exp(Q["ua"]), Q["V"]).vector().get_local()
    print("project(exp(u_new)) diff", rel(b, a), " project(exp(u_analytical)) diff", rel(b2, a2))


if __name__ == "__main__":
    main()
