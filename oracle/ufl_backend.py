"""numpy backend for the dolfin shim (oracle -- test infrastructure only).

Evaluates UFL-lite forms (fedm_b200/ufl.py) with the oracle's vectorised assembly and forward-mode AD, so that the
shim + the reference's unmodified ``fedm`` package and test scripts can be executed on the CPU in the build container
and compared with the reference's golden results.  It validates the *shim semantics* (form capture, degree estimation,
Expression handling, boundary conditions, time-stepping glue); the product path uses the CUDA backend instead.
"""
import numpy as np

from fedm_b200 import ufl as U
from . import dual as ad
from .controller import SolveFailed
from .fem import Assembler, apply_dirichlet_matrix, apply_dirichlet_vector
from .mesh import Mesh as OMesh
from .newton import newton


class NumpyStore:
    def __init__(self, n):
        self.a = np.zeros(n)

    def numpy(self):
        return self.a.copy()

    def set_numpy(self, arr):
        self.a[:] = arr

    def size(self):
        return self.a.size

    def norm2(self):
        return float(np.linalg.norm(self.a))

    def clone(self):
        s = NumpyStore(self.a.size)
        s.a[:] = self.a
        return s

    def axpy(self, alpha, other):
        self.a += alpha * other.a

    def add_scalar(self, c):
        self.a += c

    def scale(self, c):
        self.a *= c

    def copy_from(self, other):
        self.a[:] = other.a

    def update_ghosts(self, n_comp):
        pass

    def strided_copy_from(self, src, sstride, soff, dstride, doff, n):
        self.a[doff:doff + n * dstride:dstride] = src.a[soff:soff + n * sstride:sstride]


_P2_NODES = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (0, .5, .5), (.5, 0, .5), (.5, .5, 0)]
_P3_NODES = ([(1, 0, 0), (0, 1, 0), (0, 0, 1)] +
             [tuple((2 / 3 if k == i else 1 / 3 if k == j else 0.0) for k in range(3)) for i in range(3) for j in range(3) if i != j] +
             [(1 / 3, 1 / 3, 1 / 3)])


def _p2_basis(l):
    l0, l1, l2 = l
    return [l0 * (2 * l0 - 1), l1 * (2 * l1 - 1), l2 * (2 * l2 - 1), 4 * l1 * l2, 4 * l0 * l2, 4 * l0 * l1]


def _p3_basis(l):
    out = [0.5 * l[i] * (3 * l[i] - 1) * (3 * l[i] - 2) for i in range(3)]
    out += [4.5 * l[i] * l[j] * (3 * l[i] - 1) for i in range(3) for j in range(3) if i != j]
    out.append(27.0 * l[0] * l[1] * l[2])
    return out


class _Eval:
    """Evaluates UFL-lite nodes at one quadrature point of every cell (or facet)."""

    def __init__(self, ctx, unknown, coefs, normal=None):
        self.ctx, self.unknown, self.coefs, self.normal = ctx, unknown, coefs, normal
        self.cache = {}

    def lam(self):
        l = self.ctx.lam
        return (l[0], l[1], l[2]) if l.ndim == 1 else (l[:, 0], l[:, 1], l[:, 2])

    def __call__(self, e):
        key = id(e)
        if key in self.cache:
            return self.cache[key]
        v = self._eval(e)
        self.cache[key] = v
        return v

    def _eval(self, e):
        c = self.ctx
        if isinstance(e, U.Const):
            return e.value
        if isinstance(e, U.Param):
            return float(e.getter())
        if isinstance(e, U.Coord):
            return c.x[e.i]
        if isinstance(e, U.Normal):
            return self.normal[e.i]
        if isinstance(e, U.Field):
            if e.fn is self.unknown:
                return c.u(e.comp)
            return c.coef((id(e.fn), e.comp))
        if isinstance(e, U.GradComp):
            f = e.arg
            if not isinstance(f, U.Field):
                raise ValueError("unexpected gradient of a test/trial function at evaluation time")
            if f.fn is self.unknown:
                return c.grad_u(f.comp)[e.i]
            nod = c.coef_nodal((id(f.fn), f.comp))
            return (nod * c.geo.g[:, :, e.i]).sum(axis=1)
        if isinstance(e, U.Interp):
            nodes = {1: _P2_NODES[:3], 2: _P2_NODES, 3: _P3_NODES}[e.degree]
            l = self.lam()
            basis = {1: list(l), 2: _p2_basis(l), 3: _p3_basis(l)}[e.degree]
            X = c.geo.x                                           # (n, 3, 2)
            val = 0.0
            for w, b in zip(nodes, basis):
                px = w[0] * X[:, 0, 0] + w[1] * X[:, 1, 0] + w[2] * X[:, 2, 0]
                py = w[0] * X[:, 0, 1] + w[1] * X[:, 1, 1] + w[2] * X[:, 2, 1]
                val = val + b * _eval_plain(e.expr, px, py)
            return val
        if isinstance(e, U.Add):
            return self(e.a) + self(e.b)
        if isinstance(e, U.Mul):
            return self(e.a) * self(e.b)
        if isinstance(e, U.Div):
            return self(e.a) / self(e.b)
        if isinstance(e, U.Neg):
            return -self(e.a)
        if isinstance(e, U.Pow):
            return self(e.a) ** self(e.e)
        if isinstance(e, U.Func):
            a = self(e.a)
            return {"exp": ad.exp, "sqrt": ad.sqrt, "ln": ad.log}[e.name](a)
        if isinstance(e, U.Abs):
            return ad.absolute(self(e.a))
        if isinstance(e, U.Sign):
            return ad.sign(self(e.a))
        raise NotImplementedError(type(e).__name__)


def _eval_plain(e, x0, x1):
    if isinstance(e, U.Const):
        return e.value
    if isinstance(e, U.Param):
        return float(e.getter())
    if isinstance(e, U.Coord):
        return (x0, x1)[e.i]
    if isinstance(e, U.Add):
        return _eval_plain(e.a, x0, x1) + _eval_plain(e.b, x0, x1)
    if isinstance(e, U.Mul):
        return _eval_plain(e.a, x0, x1) * _eval_plain(e.b, x0, x1)
    if isinstance(e, U.Div):
        return _eval_plain(e.a, x0, x1) / _eval_plain(e.b, x0, x1)
    if isinstance(e, U.Neg):
        return -_eval_plain(e.a, x0, x1)
    if isinstance(e, U.Pow):
        return np.power(_eval_plain(e.a, x0, x1), _eval_plain(e.e, x0, x1))
    if isinstance(e, U.Func):
        return {"exp": np.exp, "sqrt": np.sqrt, "ln": np.log}[e.name](_eval_plain(e.a, x0, x1))
    if isinstance(e, U.Abs):
        return np.abs(_eval_plain(e.a, x0, x1))
    raise NotImplementedError(f"Expression strings may not contain {type(e).__name__}")


def _fields(e, out):
    if isinstance(e, U.Vec):
        _fields(e.x, out); _fields(e.y, out)
    elif isinstance(e, U.Field):
        out[(id(e.fn), e.comp)] = e
    else:
        for c in e.children():
            _fields(c, out)


class CompiledProblem:
    def __init__(self, residual, u, bcs, quad_degree, ident_zeros=False):
        self.u, self.bcs, self.ident_zeros = u, bcs, ident_zeros
        space = u.space
        tri = space.mesh().tri
        self.n_eq = space.n_comp
        self.mesh = OMesh(tri.coords, tri.cells)
        self.asm = Assembler(self.mesh, self.n_eq)
        # group integrals: all dx together, ds by subdomain id
        groups = {}
        for it in residual.integrals:
            m = it.measure
            key = ("dx", None) if m.kind == "dx" else ("ds", m.subdomain_id)
            groups.setdefault(key, {"expr": U.Const(0.0), "measure": m})
            groups[key]["expr"] = groups[key]["expr"] + it.integrand
        self.groups = []
        for (kind, sid), g in groups.items():
            if isinstance(g["expr"], U.Const):
                continue
            deg = quad_degree if quad_degree is not None else U.estimate_degree(g["expr"])
            split = U.test_function_split(g["expr"], self.n_eq)
            self.groups.append((kind, sid, g["measure"], deg, split))
        fl = {}
        for _, _, _, _, split in self.groups:
            for s, fx, fy in split:
                for e in (s, fx, fy):
                    _fields(e, fl)
        self.fields = {k: f for k, f in fl.items() if f.fn is not u}
        self.degrees = {(k, sid): d for k, sid, _, d, _ in self.groups}
        self.last_result = None

    def _coefs(self):
        out = {}
        for key, f in self.fields.items():
            a = f.fn.vector().store.a
            out[key] = a.reshape(-1, f.fn.space.n_comp)[:, f.comp]
        return out

    def _terms(self, split, ev):
        terms = []
        for s, fx, fy in split:
            zs = isinstance(s, U.Const) and s.value == 0.0
            zf = isinstance(fx, U.Const) and fx.value == 0.0 and isinstance(fy, U.Const) and fy.value == 0.0
            if zs and zf:
                terms.append(None)
                continue
            terms.append((None if zs else ev(s), None if zf else (ev(fx), ev(fy))))
        return terms

    def element_tensors(self, x, jac):
        Uarr = x.reshape(-1, self.n_eq)
        coefs = self._coefs()
        blocks_R, blocks_J = [], []
        for kind, sid, measure, deg, split in self.groups:
            if kind == "dx":
                def point(ctx, split=split):
                    return self._terms(split, _Eval(ctx, self.u, coefs))
                Re, Je = self.asm.assemble_cells(Uarr, point, deg, coefs, jac=jac)
                blocks_R.append((Re, None)); blocks_J.append((Je, None))
            else:
                mf = measure.subdomain_data
                if mf is None or sid is None:
                    markers = np.ones(self.asm.facet_vertices.shape[0], dtype=np.int32)
                    which = (1,)
                else:
                    lut = {tuple(sorted(f)): int(v) for f, v in zip(mf.facets.tolist(), mf.values.tolist())}
                    markers = np.array([lut.get(tuple(sorted(f)), 0) for f in self.asm.facet_vertices.tolist()], dtype=np.int32)
                    which = (sid,)

                def facet(ctx, nrm, split=split):
                    terms = self._terms(split, _Eval(ctx, self.u, coefs, normal=nrm))
                    return [None if t is None else t[0] for t in terms]
                fc, Rf, Jf = self.asm.assemble_facets(Uarr, facet, deg, markers, which, coefs, jac=jac)
                if Rf is not None:
                    blocks_R.append((Rf, fc)); blocks_J.append((Jf, fc))
        return blocks_R, blocks_J

    def _bc(self):
        dofs = np.concatenate([self.n_eq * bc.nodes.astype(np.int64) + bc.comp for bc in self.bcs]) if self.bcs else np.zeros(0, dtype=np.int64)
        vals = np.concatenate([np.full(bc.nodes.size, bc.current_value()) for bc in self.bcs]) if self.bcs else np.zeros(0)
        return dofs, vals

    def residual(self, x):
        R, _ = self.element_tensors(x, jac=False)
        b = np.zeros(self.asm.ndof)
        for Re, idx in R:
            self.asm.scatter_vector(Re, idx, out=b)
        dofs, vals = self._bc()
        if dofs.size:
            apply_dirichlet_vector(b, x, dofs, vals)
        return b

    def jacobian(self, x):
        _, J = self.element_tensors(x, jac=True)
        A = self.asm.scatter_matrix(J)
        if self.ident_zeros:                       # dolfin's ident_zeros(): unit diagonal on all-zero rows
            import scipy.sparse as sp
            A = A.tocsr()
            empty = np.asarray(abs(A).sum(axis=1)).ravel() == 0.0
            A = (A + sp.diags(empty.astype(float))).tocsr()
        dofs, _ = self._bc()
        if dofs.size:
            A = apply_dirichlet_matrix(A, dofs)
        return A

    def solve_newton(self, u, rtol=1e-9, atol=1e-10, stol=1e-16, max_it=50, linear_solver="mumps", raise_on_failure=True):
        x = u.vector().store.a
        hist = []
        try:
            its, conv = newton(self.residual, self.jacobian, x, rtol=rtol, atol=atol, stol=stol, max_it=max_it, history=hist)
        except SolveFailed as exc:
            self.last_result = (len(hist), False, hist)
            if raise_on_failure:
                raise RuntimeError(str(exc))
            return len(hist), False
        self.last_result = (its, conv, hist)
        return its, conv


class NumpyBackend:
    name = "numpy-oracle"

    def zeros(self, n, space=None):
        return NumpyStore(n)

    def compile(self, residual, u, bcs, quad_degree, ident_zeros=False, reuse=False):
        return CompiledProblem(residual, u, bcs, quad_degree, ident_zeros)
