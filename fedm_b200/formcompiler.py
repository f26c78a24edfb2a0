"""Form compiler: UFL-lite residual form -> CUDA source of a row-assembly model -> plug-in .so (nvcc, sm_100a).

This is the B200 counterpart of the reference's FFC/uflacs + dijitso tool chain (``F = action(F, u_new)``,
``J = derivative(F, u_new, u)``, fedm_streamer.py:296-297): the weak form a script builds through the fedm.functions API
is lowered to a ``struct JitModel`` for the generic row kernel (``fedm_b200/csrc/row_kernel.cuh``).  Per quadrature
point the integrand is written as  sum_k s_k v_k + f_k . grad(v_k)  (``ufl.test_function_split``); s_k and f_k are emitted
as straight-line code over component-sparse dual numbers (``csrc/cdual.cuh``: a sub-expression carries derivative
lanes only for the unknown components it depends on), so the Jacobian row comes out of forward-mode AD in registers.
Cell-constant sub-expressions (everything built from gradients of P1 fields, e.g. mu(|grad Phi|)) are hoisted out of the
quadrature loop.  Quadrature follows FFC: the global ``quadrature_degree`` if set, else UFL's estimated degree of the
summed integrand; FIAT's default triangle schemes, Gauss-Legendre on facets; non-symmetric rules are oriented by
ascending vertex index like FFC's reference cell.
"""
import hashlib
import os
import subprocess
from pathlib import Path

import numpy as np

from . import ufl as U

HERE = Path(__file__).resolve().parent
CACHE = HERE / "jit_cache"


# ---- quadrature tables (FIAT defaults; same constants as fedm_b200/csrc/quadrature_tables.h) ---------------------------
def _bary(pts):
    return [(1.0 - x - y, x, y) for x, y in pts]


def triangle_rule(degree):
    if degree <= 1:
        return _bary([(1 / 3, 1 / 3)]), [0.5]
    if degree == 2:
        return _bary([(1 / 6, 1 / 6), (1 / 6, 2 / 3), (2 / 3, 1 / 6)]), [1 / 6] * 3
    if degree == 3:
        p = [(0.659027622374092, 0.231933368553031), (0.659027622374092, 0.109039009072877),
             (0.231933368553031, 0.659027622374092), (0.231933368553031, 0.109039009072877),
             (0.109039009072877, 0.659027622374092), (0.109039009072877, 0.231933368553031)]
        return _bary(p), [1 / 12] * 6
    if degree == 4:
        a, b, c, d = 0.816847572980459, 0.091576213509771, 0.108103018168070, 0.445948490915965
        return (_bary([(a, b), (b, a), (b, b), (c, d), (d, c), (d, d)]),
                [0.109951743655322 / 2] * 3 + [0.223381589678011 / 2] * 3)
    if degree == 5:
        a1, a2, b1, b2 = 0.101286507323456, 0.797426985353087, 0.470142064105115, 0.059715871789770
        return (_bary([(1 / 3, 1 / 3), (a1, a1), (a2, a1), (a1, a2), (b1, b1), (b2, b1), (b1, b2)]),
                [0.225 / 2] + [0.125939180544827 / 2] * 3 + [0.132394152788506 / 2] * 3)
    if degree == 6:
        p, q, s, t = 0.873821971016996, 0.063089014491502, 0.501426509658179, 0.249286745170910
        e, f, g = 0.636502499121399, 0.310352451033785, 0.053145049844816
        pts = [(p, q), (q, p), (q, q), (s, t), (t, s), (t, t), (e, f), (e, g), (f, e), (f, g), (g, e), (g, f)]
        return _bary(pts), [0.050844906370207 / 2] * 3 + [0.116786275726379 / 2] * 3 + [0.082851075618374 / 2] * 6
    from scipy.special import roots_jacobi
    m = (degree + 2) // 2
    px, wx = roots_jacobi(m, 0.0, 0.0)
    py, wy = roots_jacobi(m, 1.0, 0.0)
    pts, w = [], []
    for x, a in zip(px, wx):
        for y, b in zip(py, wy):
            X = 0.5 * (1 + x) * (1 - y) - 1
            pts.append((0.5 * (X + 1), 0.5 * (y + 1)))
            w.append(0.5 * a * b * 0.25)
    return _bary(pts), w


def facet_rule(degree):
    m = (degree + 2) // 2
    x, w = np.polynomial.legendre.leggauss(m)
    return [0.5 * (v + 1.0) for v in x], [0.5 * v for v in w]


_P_NODES = {1: [(1, 0, 0), (0, 1, 0), (0, 0, 1)],
            2: [(1, 0, 0), (0, 1, 0), (0, 0, 1), (0, .5, .5), (.5, 0, .5), (.5, .5, 0)]}
_P_NODES[3] = (_P_NODES[1] + [tuple((2 / 3 if k == i else 1 / 3 if k == j else 0.0) for k in range(3))
                              for i in range(3) for j in range(3) if i != j] + [(1 / 3, 1 / 3, 1 / 3)])


def _basis_code(degree, l):
    l0, l1, l2 = l
    if degree == 1:
        return [l0, l1, l2]
    if degree == 2:
        return [f"{l0}*(2.0*{l0}-1.0)", f"{l1}*(2.0*{l1}-1.0)", f"{l2}*(2.0*{l2}-1.0)", f"4.0*{l1}*{l2}", f"4.0*{l0}*{l2}", f"4.0*{l0}*{l1}"]
    ls = [l0, l1, l2]
    out = [f"0.5*{ls[i]}*(3.0*{ls[i]}-1.0)*(3.0*{ls[i]}-2.0)" for i in range(3)]
    out += [f"4.5*{ls[i]}*{ls[j]}*(3.0*{ls[i]}-1.0)" for i in range(3) for j in range(3) if i != j]
    out.append(f"27.0*{l0}*{l1}*{l2}")
    return out


def _lit(v):
    return f"({float(v)!r})" if np.isfinite(v) else ("(1.0/0.0)" if v > 0 else "(-1.0/0.0)")


class _Emitter:
    """Emits straight-line code for expression DAGs with common-subexpression reuse, in three scopes: cell-constant,
    quadrature point of the cell rule, quadrature point of a facet rule."""

    def __init__(self, comp):
        self.c = comp
        self.cell_lines, self.cell_names = [], {}
        self.scope_lines, self.scope_names = None, None
        self.n = 0

    def begin_scope(self):
        self.scope_lines, self.scope_names = [], {}

    def end_scope(self):
        lines = self.scope_lines
        self.scope_lines, self.scope_names = None, None
        return lines

    def _new(self, code, dual, cell):
        self.n += 1
        name = f"t{self.n}"
        line = f"const {'auto' if dual else 'double'} {name} = {code};"
        (self.cell_lines if cell else self.scope_lines).append(line)
        return name

    def emit(self, e):
        """Returns (code string usable as an operand, is_dual).  Sub-expressions are shared structurally (hash-consing):
        scripts rebuild e.g. -grad(u[2]) or dt/dt_old many times."""
        key = self.c.skey(e)
        varies = self.c.varies(e)
        names = self.scope_names if varies else self.cell_names
        if key in names:
            return names[key]
        if not varies and self.scope_names is not None and key in self.cell_names:
            return self.cell_names[key]
        res = self._emit(e, cell=not varies)
        names[key] = res
        return res

    def _emit(self, e, cell):
        c = self.c
        if isinstance(e, U.Const):
            return _lit(e.value), False
        if isinstance(e, U.Param):
            return f"mp.p[{c.param_slot(e)}]", False
        if isinstance(e, U.Coord):
            return f"xq{e.i}", False
        if isinstance(e, U.Normal):
            return ("nr", "nz")[e.i], False
        if isinstance(e, U.Field):
            if e.fn is c.unknown:
                c.used_u.add(e.comp)
                return f"u{e.comp}", True
            j = c.field_slot(e)
            c.used_fval.add(j)
            return f"fv{j}", False
        if isinstance(e, U.GradComp):
            f = e.arg
            if f.fn is c.unknown:
                c.used_gu.add(f.comp)
                return f"gu{f.comp}_{e.i}", True
            j = c.field_slot(f)
            c.used_gf.add(j)
            return f"gf{j}_{e.i}", False
        if isinstance(e, U.Interp):
            k = c.interp_slot(e)
            return f"ip{k}", False
        if isinstance(e, (U.Add, U.Mul, U.Div)):
            a, da = self.emit(e.a)
            b, db = self.emit(e.b)
            op = {U.Add: "+", U.Mul: "*", U.Div: "/"}[type(e)]
            return self._new(f"{a} {op} {b}", da or db, cell), da or db
        if isinstance(e, U.Neg):
            a, da = self.emit(e.a)
            return self._new(f"-{a}", da, cell), da
        if isinstance(e, U.Pow):
            a, da = self.emit(e.a)
            if not isinstance(e.e, U.Const):
                b, db = self.emit(e.e)
                if da or db:
                    raise NotImplementedError("power with a non-constant exponent that depends on the unknowns")
                return self._new(f"pow({a}, {b})", False, cell), False
            fn = "dpow" if da else "pow"
            return self._new(f"{fn}({a}, {_lit(e.e.value)})", da, cell), da
        if isinstance(e, U.Func):
            a, da = self.emit(e.a)
            fn = {"exp": ("exp", "dexp"), "sqrt": ("sqrt", "dsqrt"), "ln": ("log", "dlog")}[e.name][1 if da else 0]
            return self._new(f"{fn}({a})", da, cell), da
        if isinstance(e, U.Abs):
            a, da = self.emit(e.a)
            return self._new(f"{'dabs' if da else 'fabs'}({a})", da, cell), da
        if isinstance(e, U.Sign):
            a, da = self.emit(e.a)
            return self._new(f"dsign({a})", False, cell), False
        raise NotImplementedError(type(e).__name__)


def _plain_code(e, comp):
    """C expression in X0, X1 (doubles) for the inside of an interpolated Expression."""
    if isinstance(e, U.Const):
        return _lit(e.value)
    if isinstance(e, U.Param):
        return f"mp.p[{comp.param_slot(e)}]"
    if isinstance(e, U.Coord):
        return f"X{e.i}"
    if isinstance(e, (U.Add, U.Mul, U.Div)):
        op = {U.Add: "+", U.Mul: "*", U.Div: "/"}[type(e)]
        return f"({_plain_code(e.a, comp)} {op} {_plain_code(e.b, comp)})"
    if isinstance(e, U.Neg):
        return f"(-{_plain_code(e.a, comp)})"
    if isinstance(e, U.Pow):
        return f"pow({_plain_code(e.a, comp)}, {_plain_code(e.e, comp)})"
    if isinstance(e, U.Func):
        return f"{ {'exp': 'exp', 'sqrt': 'sqrt', 'ln': 'log'}[e.name] }({_plain_code(e.a, comp)})"
    if isinstance(e, U.Abs):
        return f"fabs({_plain_code(e.a, comp)})"
    raise NotImplementedError(f"Expression strings may not contain {type(e).__name__}")


class CompiledForm:
    """Source + metadata of one compiled residual form."""

    def __init__(self, residual, unknown, quad_degree=None):
        self.unknown = unknown
        self.n_eq = unknown.space.n_comp
        self.fields, self._field_ids = [], {}          # [(fn, comp)]
        self.params, self._param_ids = [], {}          # [Param]
        self.interps, self._interp_ids = [], {}
        self._varies = {}
        self._skeys = {}
        self._deps = {}
        self.used_u, self.used_gu, self.used_fval, self.used_gf = set(), set(), set(), set()
        groups = {}
        self.facet_data = None
        for it in residual.integrals:
            m = it.measure
            key = ("dx", None) if m.kind == "dx" else ("ds", m.subdomain_id)
            g = groups.setdefault(key, {"expr": U.Const(0.0), "measure": m})
            g["expr"] = g["expr"] + it.integrand
            if m.kind == "ds" and m.subdomain_data is not None:
                self.facet_data = m.subdomain_data
        self.groups = []
        for (kind, sid), g in groups.items():
            if isinstance(g["expr"], U.Const):
                continue
            deg = quad_degree if quad_degree is not None else U.estimate_degree(g["expr"])
            self.groups.append((kind, sid, deg, U.test_function_split(g["expr"], self.n_eq)))
        self.source = self._generate()
        # the cache key covers the library headers the plug-in is compiled against (AsmArgs / ModelParams layout)
        h = hashlib.sha1(self.source.encode())
        for name in ("row_kernel.cuh", "fedmx_internal.h", "cdual.cuh"):
            h.update((HERE / "csrc" / name).read_bytes())
        h.update((HERE.parent / "include" / "fedmx.h").read_bytes())
        self.key = h.hexdigest()[:20]

    # ---- symbol tables
    def field_slot(self, f):
        k = (id(f.fn), f.comp)
        if k not in self._field_ids:
            self._field_ids[k] = len(self.fields)
            self.fields.append((f.fn, f.comp))
        return self._field_ids[k]

    def param_slot(self, p):
        if id(p) not in self._param_ids:
            self._param_ids[id(p)] = len(self.params)
            self.params.append(p)
        return self._param_ids[id(p)]

    def interp_slot(self, e):
        if id(e) not in self._interp_ids:
            self._interp_ids[id(e)] = len(self.interps)
            self.interps.append(e)
        return self._interp_ids[id(e)]

    def skey(self, e):
        """Structural key of an expression (memoised per node)."""
        k = id(e)
        if k not in self._skeys:
            if isinstance(e, U.Const):
                v = ("c", e.value)
            elif isinstance(e, U.Param):
                v = ("p", id(e))
            elif isinstance(e, (U.Coord, U.Normal)):
                v = (type(e).__name__, e.i)
            elif isinstance(e, U.Field):
                v = ("f", id(e.fn), e.comp)
            elif isinstance(e, U.GradComp):
                v = ("g", self.skey(e.arg), e.i)
            elif isinstance(e, U.Interp):
                v = ("i", id(e))
            elif isinstance(e, U.Func):
                v = ("fn", e.name, self.skey(e.a))
            else:
                v = (type(e).__name__,) + tuple(self.skey(ch) for ch in e.children())
            self._skeys[k] = v
        return self._skeys[k]

    def varies(self, e):
        """True if the value changes between quadrature points of a cell."""
        k = id(e)
        if k not in self._varies:
            if isinstance(e, (U.Field, U.Coord, U.Interp, U.Normal)):
                v = True
            elif isinstance(e, (U.Const, U.Param, U.GradComp)):
                v = False
            else:
                v = any(self.varies(ch) for ch in e.children())
            self._varies[k] = v
        return self._varies[k]

    def deps(self, e):
        """Components of the unknown an expression depends on (a superset of what the emitted C++ type carries)."""
        k = id(e)
        if k not in self._deps:
            if isinstance(e, U.Field):
                v = frozenset([e.comp]) if e.fn is self.unknown else frozenset()
            elif isinstance(e, U.Interp):
                v = frozenset()
            else:
                v = frozenset().union(*[self.deps(ch) for ch in e.children()]) if e.children() else frozenset()
            self._deps[k] = v
        return self._deps[k]

    def param_values(self):
        return np.array([float(p.getter()) for p in self.params] + [0.0], dtype=np.float64)

    # ---- code generation
    def _generate(self):
        neq = self.n_eq
        em = _Emitter(self)
        cell_blocks, facet_blocks, tables = [], [], []

        def zero(x):
            return isinstance(x, U.Const) and x.value == 0.0

        for gi, (kind, sid, deg, split) in enumerate(self.groups):
            em.begin_scope()
            acc = []
            for k, (s, fx, fy) in enumerate(split):
                if zero(s) and zero(fx) and zero(fy):
                    continue
                parts = []
                if not zero(s):
                    parts.append(f"({em.emit(s)[0]}) * LA")
                if not zero(fx):
                    parts.append(f"({em.emit(fx)[0]}) * ga_r")
                if not zero(fy):
                    parts.append(f"({em.emit(fy)[0]}) * ga_z")
                acc.append(f"R[{k}] = R[{k}] + ({' + '.join(parts)}) * W;")
            scope = em.end_scope()
            # cell-wise variant: one quadrature loop per equation k (only 5 dual accumulators live at a time: the integrand
            # against the three test functions, S_i += s W l_i, and its flux part F += f W, which meets grad v_i after the loop)
            per_k = []
            if kind == "dx":
                for k, (s, fx, fy) in enumerate(split):
                    if zero(s) and zero(fx) and zero(fy):
                        continue
                    em.begin_scope()
                    lines = []
                    if not zero(s):
                        code = em.emit(s)[0]
                        lines += [f"const auto sW = ({code}) * W;", "S0 = S0 + sW * l0;", "S1 = S1 + sW * l1;", "S2 = S2 + sW * l2;"]
                    if not zero(fx):
                        lines.append(f"FX = FX + ({em.emit(fx)[0]}) * W;")
                    if not zero(fy):
                        lines.append(f"FY = FY + ({em.emit(fy)[0]}) * W;")
                    mask = 0
                    for e in (s, fx, fy):
                        for c in self.deps(e):
                            mask |= 1 << c
                    per_k.append((k, mask, not zero(fx), not zero(fy), em.end_scope() + lines))
            if kind == "dx":
                pts, w = triangle_rule(deg)
                tables.append(f"static __device__ __constant__ double QC{gi}[{len(w)}][4] = {{" +
                              ", ".join("{" + ", ".join(repr(float(v)) for v in (*p, ww)) + "}" for p, ww in zip(pts, w)) + "};")
                cell_blocks.append((gi, len(w), scope + acc, per_k))
            else:
                t, w = facet_rule(deg)
                tables.append(f"static __device__ __constant__ double QF{gi}[{len(w)}][2] = {{" +
                              ", ".join("{" + f"{float(a)!r}, {float(b)!r}" + "}" for a, b in zip(t, w)) + "};")
                facet_blocks.append((gi, sid, len(w), scope + acc))

        nf = len(self.fields)
        rec_jac = (neq + 3 * neq * neq + 1) // 2 * 2          # doubles per (cell, test vertex) record, 16-byte aligned
        rec_res = (neq + 1) // 2 * 2
        # cell-wise evaluation (row_kernel.cuh) pays when the integrand is expensive and the 3 n_eq (1 + 3 n_eq) accumulators
        # of a cell still fit the register file
        cost = sum(nq * len(body) for _, nq, body, _ in cell_blocks)
        self.cellwise = len(cell_blocks) == 1 and neq <= 3 and cost >= 60
        L = []
        L.append("// Generated by fedm_b200/formcompiler.py -- do not edit.")
        L.append('#include "row_kernel.cuh"')
        L.append('#include "cdual.cuh"')
        L.append("namespace {")
        L += tables
        for k, ip in enumerate(self.interps):
            L.append(f"static __device__ __forceinline__ double interp_fn{k}(double X0, double X1, const ModelParams &mp) {{ return {_plain_code(ip.expr, self)}; }}")
        L.append("struct JitModel {")
        L.append("    static constexpr int MINB = 2, MINB_RES = 2;")
        L.append("    static constexpr unsigned ZERO_MASK = 0;")
        L.append(f"    static constexpr int NEQ = {neq};\n    static constexpr int NF = {nf};\n    static constexpr int NC = 0;")
        L.append(f"    static constexpr bool CELLS = {'true' if self.cellwise else 'false'};")
        cells_jac = os.environ.get("FEDMX_JIT_CELLS_JAC", "1") != "0"       # tuning knob: Jacobian launches row-wise
        L.append(f"    static constexpr bool CELLS_JAC = {'true' if cells_jac else 'false'};")
        L.append(f"    static constexpr int REC_JAC = {rec_jac}, REC_RES = {rec_res}, CELL_MINB = {int(os.environ.get('FEDMX_JIT_CELL_MINB', '2'))};")
        L.append(f"    typedef Vtx<{neq}, {nf}, 0> V;")
        if neq > 16:
            raise NotImplementedError("more than 16 unknown components")

        def cell_prelude():
            """Geometry, gradients of the unknowns (seeded duals) and of the coefficient fields, cell-constant sub-expressions:
            shared by row() (a = the row's vertex) and cell() (a, p, n = the cell's vertices)."""
            P = []
            P.append(f"        constexpr unsigned ALL = JAC ? {(1 << neq) - 1}u : 0u;   // residual-only: the same code on value-only duals")
            P.append("        const double e1r = p.r - a.r, e1z = p.z - a.z, e2r = n.r - a.r, e2z = n.z - a.z;")
            P.append("        const double det = e1r * e2z - e1z * e2r, inv = 1.0 / det, adet = fabs(det);")
            P.append("        const double g_r[3] = {-(e2z - e1z) * inv, e2z * inv, -e1z * inv};")
            P.append("        const double g_z[3] = {-(e1r - e2r) * inv, -e2r * inv, e1r * inv};")
            P.append("        // FFC's reference cell lists the vertices by ascending index: ranks of (a,p,n) in that order")
            P.append("        const int ra = (a.id > p.id) + (a.id > n.id), rp = (p.id > a.id) + (p.id > n.id), rn = 3 - ra - rp;")
            for k in sorted(self.used_gu):
                for i, g in enumerate(("g_r", "g_z")):
                    P.append(f"        const auto gu{k}_{i} = cd_seed<JAC ? {1 << k}u : 0u>(a.u[{k}] * {g}[0] + p.u[{k}] * {g}[1] + "
                             f"n.u[{k}] * {g}[2], {g}[0], {g}[1], {g}[2]);")
            for j in sorted(self.used_gf):
                for i, g in enumerate(("g_r", "g_z")):
                    P.append(f"        const double gf{j}_{i} = a.f[{j}] * {g}[0] + p.f[{j}] * {g}[1] + n.f[{j}] * {g}[2];")
            # nodal values of interpolated Expressions (reference vertex order = ascending index, harmless for symmetric node sets)
            for k, ip in enumerate(self.interps):
                nodes = _P_NODES[ip.degree]
                P.append(f"        double nv{k}[{len(nodes)}];")
                for b, w in enumerate(nodes):
                    P.append(f"        nv{k}[{b}] = interp_fn{k}({w[0]!r} * a.r + {w[1]!r} * p.r + {w[2]!r} * n.r, {w[0]!r} * a.z + {w[1]!r} * p.z + {w[2]!r} * n.z, mp);")
            P += ["        " + ln for ln in em.cell_lines]
            return P

        def point_prelude(l0, l1, l2, indent):
            P = []
            P.append(f"const double xq0 = {l0} * a.r + {l1} * p.r + {l2} * n.r, xq1 = {l0} * a.z + {l1} * p.z + {l2} * n.z;")
            for k in sorted(self.used_u):
                P.append(f"const auto u{k} = cd_seed<JAC ? {1 << k}u : 0u>({l0} * a.u[{k}] + {l1} * p.u[{k}] + {l2} * n.u[{k}], "
                         f"{l0}, {l1}, {l2});")
            for j in sorted(self.used_fval):
                P.append(f"const double fv{j} = {l0} * a.f[{j}] + {l1} * p.f[{j}] + {l2} * n.f[{j}];")
            for k, ip in enumerate(self.interps):
                basis = _basis_code(ip.degree, (l0, l1, l2))
                P.append(f"const double ip{k} = " + " + ".join(f"({b}) * nv{k}[{i}]" for i, b in enumerate(basis)) + ";")
            return [indent + ln for ln in P]

        # ---- row(): row `a` of the element tensor of cell (a, p, n), cell and exterior-facet integrals
        L.append("    template <bool JAC>")
        L.append(f"    static __device__ __noinline__ void row(const ModelParams &mp, const V &a, const V &p, const V &n, uint32_t mk_ap, uint32_t mk_an, RowOut<{neq}> &o) {{")
        L += cell_prelude()
        L.append("        const double ga_r = g_r[0], ga_z = g_z[0];")
        L.append(f"        CD<ALL> R[{neq}];")
        L.append(f"        for (int k = 0; k < {neq}; ++k) R[k] = cd_zero<ALL>();")
        for gi, nq, body, _ in cell_blocks:
            L.append(f"        for (int q = 0; q < {nq}; ++q) {{")
            L.append(f"            const double l0 = QC{gi}[q][ra], l1 = QC{gi}[q][rp], l2 = QC{gi}[q][rn];")
            L.append(f"            const double W = QC{gi}[q][3] * adet, LA = l0;")
            L += point_prelude("l0", "l1", "l2", "            ")
            L += ["            " + ln for ln in body]
            L.append("        }")
        if facet_blocks:
            L.append("        for (int side = 0; side < 2; ++side) {")
            L.append("            const uint32_t mk = side == 0 ? mk_ap : mk_an;")
            L.append("            if (mk == 0) continue;")
            L.append("            const V &b = side == 0 ? p : n;")
            L.append("            const V &c = side == 0 ? n : p;")
            L.append("            const double tr_ = b.r - a.r, tz_ = b.z - a.z, FL = sqrt(tr_ * tr_ + tz_ * tz_);")
            L.append("            double nr = tz_ / FL, nz = -tr_ / FL;")
            L.append("            if (nr * (c.r - a.r) + nz * (c.z - a.z) > 0.0) { nr = -nr; nz = -nz; }")
            for gi, sid, nq, body in facet_blocks:
                cond = "true" if sid is None else f"mk == {int(sid)}u"
                L.append(f"            if ({cond}) for (int q = 0; q < {nq}; ++q) {{")
                L.append(f"                const double tq = QF{gi}[q][0];")
                L.append("                const double l0 = 1.0 - tq, l1 = side == 0 ? tq : 0.0, l2 = side == 0 ? 0.0 : tq;")
                L.append(f"                const double W = QF{gi}[q][1] * FL, LA = l0;")
                L += point_prelude("l0", "l1", "l2", "                ")
                L += ["                " + ln for ln in body]
                L.append("            }")
            L.append("        }")
        L.append(f"        for (int k = 0; k < {neq}; ++k) {{")
        L.append("            o.R[k] = R[k].v;")
        L.append(f"            if constexpr (JAC) for (int bb = 0; bb < 3; ++bb) for (int l = 0; l < {neq}; ++l) o.J[bb][k * {neq} + l] = R[k].d[3 * l + bb];")
        L.append("        }")
        L.append("    }")
        # ---- cell(): the cell integrals of all three rows at once, written as three records (row_kernel.cuh, cell-wise path)
        L.append("    template <bool JAC>")
        L.append("    static __device__ __forceinline__ void cell(const ModelParams &mp, const V &a, const V &p, const V &n, double *rec) {")
        if self.cellwise:
            L += cell_prelude()
            L.append("        constexpr int REC = JAC ? REC_JAC : REC_RES, RS = FEDMX_SLICE;   // record element stride (see row_kernel.cuh)")
            done = set()
            for gi, nq, _, per_k in cell_blocks:
                for k, mask, has_fx, has_fy, body in per_k:
                    if k in done:
                        raise NotImplementedError("cell-wise path: several cell quadrature rules for one equation")
                    done.add(k)
                    L.append("        {")
                    L.append(f"            constexpr unsigned MK = JAC ? {mask}u : 0u;       // components equation {k} depends on")
                    L.append("            CD<MK> S0 = cd_zero<MK>(), S1 = cd_zero<MK>(), S2 = cd_zero<MK>(), FX = cd_zero<MK>(), FY = cd_zero<MK>();")
                    L.append(f"            for (int q = 0; q < {nq}; ++q) {{")
                    L.append(f"                const double l0 = QC{gi}[q][ra], l1 = QC{gi}[q][rp], l2 = QC{gi}[q][rn];")
                    L.append(f"                const double W = QC{gi}[q][3] * adet;")
                    L += point_prelude("l0", "l1", "l2", "                ")
                    L += ["                " + ln for ln in body]
                    L.append("            }")
                    for i in range(3):
                        terms = [f"S{i}"] + ([f"FX * g_r[{i}]"] if has_fx else []) + ([f"FY * g_z[{i}]"] if has_fy else [])
                        L.append(f"            const CD<MK> T{i} = {' + '.join(terms)};")
                        L.append(f"            rec[({i} * REC + {k}) * RS] = T{i}.v;")
                        L.append("            if constexpr (JAC) {")
                        for bb in range(3):
                            for l in range(neq):
                                if mask >> l & 1:
                                    rank = bin(mask & ((1 << l) - 1)).count("1")
                                    val = f"T{i}.d[{3 * rank + bb}]"
                                else:
                                    val = "0.0"
                                L.append(f"                rec[({i} * REC + {neq + bb * neq * neq + k * neq + l}) * RS] = {val};")
                        L.append("            }")
                    L.append("        }")
            for k in range(neq):
                if k not in done:                       # equation without a cell integral: zero records
                    for i in range(3):
                        L.append(f"        rec[({i} * REC + {k}) * RS] = 0.0;")
                        L.append("        if constexpr (JAC) {")
                        for bb in range(3):
                            for l in range(neq):
                                L.append(f"            rec[({i} * REC + {neq + bb * neq * neq + k * neq + l}) * RS] = 0.0;")
                        L.append("        }")
        L.append("    }")
        L.append("};")
        L.append("}  // namespace")
        L.append('extern "C" int fedmx_plugin_neq() { return JitModel::NEQ; }')
        L.append('extern "C" int fedmx_plugin_nf() { return JitModel::NF; }')
        L.append('extern "C" int fedmx_plugin_args_size() { return int(sizeof(AsmArgs)); }')
        L.append('extern "C" int fedmx_plugin_cell_doubles(int jac) { return JitModel::CELLS ? 3 * (jac ? JitModel::REC_JAC : JitModel::REC_RES) : 0; }')
        L.append('extern "C" int fedmx_plugin_launch(const void *args, int jac, int minb, void *stream) {')
        L.append("    launch_rows<JitModel>(*(const AsmArgs *)args, jac != 0, minb, (cudaStream_t)stream);")
        L.append("    return cudaGetLastError() == cudaSuccess ? 0 : 1;")
        L.append("}")
        return "\n".join(L) + "\n"

    # ---- build
    def build(self, verbose=False):
        """nvcc -> fedm_b200/jit_cache/<hash>.so (cached by source hash).  Fails loudly if nvcc is unavailable."""
        CACHE.mkdir(exist_ok=True)
        cu, so = CACHE / f"form_{self.key}.cu", CACHE / f"form_{self.key}.so"
        if so.exists():
            return so
        tmp = f"{so}.{os.getpid()}.tmp"                     # per process: ranks compiling the same form must not share it
        if not cu.exists() or cu.read_text() != self.source:
            cu_tmp = f"{cu}.{os.getpid()}.tmp"
            Path(cu_tmp).write_text(self.source)
            os.replace(cu_tmp, cu)
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
               "-shared", "-I", str(HERE / "csrc"), "-o", tmp, str(cu)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"form compiler: nvcc failed for {cu}\n{res.stderr[-4000:]}")
        os.replace(tmp, so)
        return so
