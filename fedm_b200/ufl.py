"""UFL-lite: the symbolic layer under the problem-definition API.

Captures exactly what the reference's scripts and ``fedm/functions.py`` build with UFL:
scalar and 2-vector expressions over P1 unknowns / coefficient functions / mutable scalar parameters / spatial
coordinates / facet normals, with ``+ - * / **``, ``exp sqrt ln abs``, ``grad dot inner``, indexing, measures
``dx`` / ``ds(i)``, forms as sums of integrals, ``action``, ``derivative``, ``lhs`` / ``rhs``.  Two consumers:

* the JIT form compiler ``fedm_b200/formcompiler.py`` lowers a form to a CUDA row kernel (product path);
* degree estimation (``estimate_degree``) reproduces UFL's ``SumDegreeEstimator`` rules so that scripts which do not set
  ``parameters["form_compiler"]["quadrature_degree"]`` get FFC's quadrature (time of flight: degree 8).

Spatial gradients are expanded symbolically through every P1 field (``grad(D*exp(u))`` -> chain rule), which is
what UFL's ``apply_derivatives`` does before FFC sees the form.
"""
import math
import numbers


# ---------------------------------------------------------------------------------------------------------------------
# expression nodes
# ---------------------------------------------------------------------------------------------------------------------
class Expr:
    """Scalar expression node.  ``shape`` is () for scalars and (2,) for vectors (see Vec)."""
    shape = ()

    # arithmetic -------------------------------------------------------------------------------------------------------
    def __add__(self, o):
        return _add(self, as_expr(o))

    def __radd__(self, o):
        return _add(as_expr(o), self)

    def __sub__(self, o):
        return _add(self, _neg(as_expr(o)))

    def __rsub__(self, o):
        return _add(as_expr(o), _neg(self))

    def __mul__(self, o):
        if isinstance(o, Measure):
            return Form([Integral(self, o)])
        o = as_expr(o)
        if isinstance(o, Vec):
            return Vec(self * o.x, self * o.y)
        return _mul(self, o)

    def __rmul__(self, o):
        return _mul(as_expr(o), self)

    def __truediv__(self, o):
        return _div(self, as_expr(o))

    def __rtruediv__(self, o):
        return _div(as_expr(o), self)

    def __neg__(self):
        return _neg(self)

    def __pos__(self):
        return self

    def __pow__(self, e):
        e = as_expr(e)
        if isinstance(e, Const):             # UFL's Power constructor folds these
            if e.value == 0.0:
                return Const(1.0, is_int=True)
            if e.value == 1.0:
                return self
        return Pow(self, e)

    def __rpow__(self, b):
        return Pow(as_expr(b), self)

    def __abs__(self):
        return Abs(self)

    def children(self):
        return ()


class Const(Expr):
    def __init__(self, value, is_int=False):
        self.value = float(value)
        self.is_int = is_int            # UFL distinguishes x**2 (IntValue) from x**2.0 (FloatValue) when estimating degrees

    def __repr__(self):
        return repr(self.value)


class Param(Expr):
    """Scalar read from a Python object at assembly time: ``Expression("time_step", time_step=..)``, ``Constant``, or a
    degree-0 Expression whose C string is evaluated on the host (``getter`` returns the current float)."""

    def __init__(self, getter, name="p"):
        self.getter, self.name = getter, name

    def __repr__(self):
        return f"Param({self.name})"


class Coord(Expr):
    """x[i] (``Expression('x[0]', degree=1)`` / SpatialCoordinate)."""

    def __init__(self, i):
        self.i = i

    def __repr__(self):
        return f"x[{self.i}]"


class Normal(Expr):
    def __init__(self, i):
        self.i = i

    def __repr__(self):
        return f"n[{self.i}]"


class Field(Expr):
    """Component ``comp`` of a P1 function living in a device vector (``fn.vector()`` with ``n_comp`` interleaved
    components).  ``unknown`` marks the components of the function the form is differentiated with respect to."""

    def __init__(self, fn, comp=0):
        self.fn, self.comp = fn, comp

    def __repr__(self):
        return f"{getattr(self.fn, '_name', 'f')}[{self.comp}]"


class Arg(Expr):
    """Test (number 0) or trial (number 1) function component."""

    def __init__(self, space, comp, number):
        self.space, self.comp, self.number = space, comp, number

    def __repr__(self):
        return ("v" if self.number == 0 else "du") + f"[{self.comp}]"


class GradComp(Expr):
    """d(arg)/dx_i for a Field or Arg (cell-wise constant for P1)."""

    def __init__(self, arg, i):
        self.arg, self.i = arg, i

    def children(self):
        return (self.arg,)

    def __repr__(self):
        return f"d{self.arg!r}/dx{self.i}"


class Interp(Expr):
    """Nodal interpolant of an expression in x (and parameters) into P_degree: how FFC treats ``Expression(cpp, degree=k)``
    inside a form."""

    def __init__(self, expr, degree):
        self.expr, self.degree = expr, degree

    def children(self):
        return (self.expr,)


class Add(Expr):
    def __init__(self, a, b):
        self.a, self.b = a, b

    def children(self):
        return (self.a, self.b)


class Mul(Expr):
    def __init__(self, a, b):
        self.a, self.b = a, b

    def children(self):
        return (self.a, self.b)


class Div(Expr):
    def __init__(self, a, b):
        self.a, self.b = a, b

    def children(self):
        return (self.a, self.b)


class Neg(Expr):
    def __init__(self, a):
        self.a = a

    def children(self):
        return (self.a,)


class Pow(Expr):
    def __init__(self, a, e):
        self.a, self.e = a, e

    def children(self):
        return (self.a, self.e)


class Func(Expr):
    """exp | sqrt | ln."""

    def __init__(self, name, a):
        self.name, self.a = name, a

    def children(self):
        return (self.a,)


class Abs(Expr):
    def __init__(self, a):
        self.a = a

    def children(self):
        return (self.a,)


class Vec:
    """2-vector of scalar expressions (grad, E, FacetNormal, vector constants)."""
    shape = (2,)

    def __init__(self, x, y):
        self.x, self.y = as_expr(x), as_expr(y)

    def __getitem__(self, i):
        return (self.x, self.y)[i]

    def __len__(self):
        return 2

    def __neg__(self):
        return Vec(-self.x, -self.y)

    def __add__(self, o):
        if _is_zero_scalar(o):
            return self
        return Vec(self.x + o.x, self.y + o.y)

    __radd__ = __add__

    def __sub__(self, o):
        return Vec(self.x - o.x, self.y - o.y)

    def __mul__(self, o):
        o = as_expr(o)
        return Vec(self.x * o, self.y * o)

    __rmul__ = __mul__

    def __truediv__(self, o):
        o = as_expr(o)
        return Vec(self.x / o, self.y / o)


def _is_zero_scalar(o):
    return isinstance(o, numbers.Real) and o == 0 or isinstance(o, Const) and o.value == 0.0


def as_expr(o):
    if isinstance(o, (Expr, Vec)):
        return o
    if isinstance(o, numbers.Real):
        return Const(o, is_int=isinstance(o, numbers.Integral) and not isinstance(o, bool))
    if hasattr(o, "ufl_expr"):
        return o.ufl_expr()
    raise TypeError(f"cannot use {type(o).__name__} in a form expression")


def _add(a, b):
    if isinstance(a, Vec) or isinstance(b, Vec):
        if _is_zero_scalar(a):
            return b
        if _is_zero_scalar(b):
            return a
        return Vec(a.x + b.x, a.y + b.y)
    if isinstance(a, Const) and a.value == 0.0:
        return b
    if isinstance(b, Const) and b.value == 0.0:
        return a
    if isinstance(a, Const) and isinstance(b, Const):
        return Const(a.value + b.value)
    return Add(a, b)


def _neg(a):
    if isinstance(a, Vec):
        return Vec(_neg(a.x), _neg(a.y))
    if isinstance(a, Const):
        return Const(-a.value)
    if isinstance(a, Neg):
        return a.a
    return Neg(a)


def _mul(a, b):
    if isinstance(b, Vec):
        return Vec(_mul(a, b.x), _mul(a, b.y))
    if isinstance(a, Vec):
        return Vec(_mul(a.x, b), _mul(a.y, b))
    for p, q in ((a, b), (b, a)):
        if isinstance(p, Const):
            if p.value == 0.0:
                return Const(0.0)          # UFL folds 0*expr to Zero (this is why sign=0 species drop out of forms)
            if p.value == 1.0:
                return q
    if isinstance(a, Const) and isinstance(b, Const):
        return Const(a.value * b.value)
    return Mul(a, b)


def _div(a, b):
    if isinstance(a, Vec):
        return Vec(_div(a.x, b), _div(a.y, b))
    if isinstance(a, Const) and a.value == 0.0:
        return Const(0.0)
    if isinstance(b, Const) and b.value == 1.0:
        return a
    if isinstance(a, Const) and isinstance(b, Const):
        return Const(a.value / b.value)
    return Div(a, b)


# public functions with UFL's names -------------------------------------------------------------------------------------
def exp(a):
    a = as_expr(a)
    return Const(math.exp(a.value)) if isinstance(a, Const) else Func("exp", a)


def sqrt(a):
    a = as_expr(a)
    return Const(math.sqrt(a.value)) if isinstance(a, Const) else Func("sqrt", a)


def ln(a):
    a = as_expr(a)
    return Const(math.log(a.value)) if isinstance(a, Const) else Func("ln", a)


def dot(a, b):
    a, b = as_expr(a), as_expr(b)
    if isinstance(a, Vec) and isinstance(b, Vec):
        return a.x * b.x + a.y * b.y
    return a * b


inner = dot


def grad(e):
    """Spatial gradient, expanded through every P1 field / coordinate the expression contains."""
    e = as_expr(e)
    if isinstance(e, Vec):
        raise NotImplementedError("grad of a vector is not needed by the FEDM form family")
    return Vec(_ddx(e, 0), _ddx(e, 1))


def _ddx(e, i):
    if isinstance(e, (Const, Param, Normal)):
        return Const(0.0)
    if isinstance(e, Coord):
        return Const(1.0 if e.i == i else 0.0)
    if isinstance(e, (Field, Arg)):
        return GradComp(e, i)
    if isinstance(e, GradComp):
        return Const(0.0)                    # second derivatives of P1 functions vanish cell-wise
    if isinstance(e, Interp):
        raise NotImplementedError("grad of an interpolated Expression is not needed by the FEDM form family")
    if isinstance(e, Add):
        return _ddx(e.a, i) + _ddx(e.b, i)
    if isinstance(e, Neg):
        return -_ddx(e.a, i)
    if isinstance(e, Mul):
        return _ddx(e.a, i) * e.b + e.a * _ddx(e.b, i)
    if isinstance(e, Div):
        return _ddx(e.a, i) / e.b - (e * _ddx(e.b, i)) / e.b
    if isinstance(e, Pow):
        if not isinstance(e.e, Const):
            raise NotImplementedError("grad of a power with non-constant exponent")
        return Const(e.e.value) * Pow(e.a, Const(e.e.value - 1.0)) * _ddx(e.a, i)
    if isinstance(e, Func):
        d = _ddx(e.a, i)
        if e.name == "exp":
            return e * d
        if e.name == "sqrt":
            return d / (Const(2.0) * e)
        if e.name == "ln":
            return d / e.a
    if isinstance(e, Abs):
        return Sign(e.a) * _ddx(e.a, i)
    raise NotImplementedError(type(e).__name__)


class Sign(Expr):
    """sign(a), derivative of |a| (0 at 0), never differentiated further."""

    def __init__(self, a):
        self.a = a

    def children(self):
        return (self.a,)


# ---------------------------------------------------------------------------------------------------------------------
# measures, integrals, forms
# ---------------------------------------------------------------------------------------------------------------------
class Measure:
    def __init__(self, kind, domain=None, subdomain_data=None, subdomain_id=None):
        self.kind, self.domain, self.subdomain_data, self.subdomain_id = kind, domain, subdomain_data, subdomain_id

    def __call__(self, subdomain_id):
        return Measure(self.kind, self.domain, self.subdomain_data, subdomain_id)

    def __rmul__(self, integrand):
        return Form([Integral(as_expr(integrand), self)])


class Integral:
    def __init__(self, integrand, measure):
        self.integrand, self.measure = integrand, measure


class Form:
    def __init__(self, integrals):
        self.integrals = list(integrals)

    def __add__(self, o):
        if _is_zero_scalar(o):
            return self
        return Form(self.integrals + o.integrals)

    __radd__ = __add__

    def __sub__(self, o):
        if _is_zero_scalar(o):
            return self
        return Form(self.integrals + [Integral(-i.integrand, i.measure) for i in o.integrals])

    def __rsub__(self, o):
        return (-self) + o

    def __neg__(self):
        return Form([Integral(-i.integrand, i.measure) for i in self.integrals])

    def __mul__(self, c):
        c = as_expr(c)
        return Form([Integral(c * i.integrand, i.measure) for i in self.integrals])

    __rmul__ = __mul__


def substitute(e, fn):
    """Rebuild an expression bottom-up; fn(node) may return a replacement for leaves (or None)."""
    r = fn(e)
    if r is not None:
        return r
    if isinstance(e, Vec):
        return Vec(substitute(e.x, fn), substitute(e.y, fn))
    if isinstance(e, Add):
        return substitute(e.a, fn) + substitute(e.b, fn)
    if isinstance(e, Mul):
        return substitute(e.a, fn) * substitute(e.b, fn)
    if isinstance(e, Div):
        return substitute(e.a, fn) / substitute(e.b, fn)
    if isinstance(e, Neg):
        return -substitute(e.a, fn)
    if isinstance(e, Pow):
        return Pow(substitute(e.a, fn), substitute(e.e, fn))
    if isinstance(e, Func):
        return Func(e.name, substitute(e.a, fn))
    if isinstance(e, Abs):
        return Abs(substitute(e.a, fn))
    if isinstance(e, Sign):
        return Sign(substitute(e.a, fn))
    if isinstance(e, GradComp):
        inner_ = substitute(e.arg, fn)
        return _ddx(inner_, e.i)
    if isinstance(e, Interp):
        return Interp(substitute(e.expr, fn), e.degree)
    return e


def action(form, fn):
    """Replace the trial function of ``form`` by the Function ``fn`` (component-wise)."""
    def rep(e):
        if isinstance(e, Arg) and e.number == 1:
            return Field(fn, e.comp)
        return None
    return Form([Integral(substitute(i.integrand, rep), i.measure) for i in form.integrals])


class Jacobian:
    """``derivative(F, u, du)``: the Gateaux derivative is taken by forward-mode AD inside the generated kernel, so the
    object only records what is differentiated with respect to what."""

    def __init__(self, form, fn):
        self.form, self.fn = form, fn


def derivative(form, fn, du=None):
    return Jacobian(form, fn)


def _has_trial(e):
    if isinstance(e, Arg):
        return e.number == 1
    if isinstance(e, Vec):
        return _has_trial(e.x) or _has_trial(e.y)
    return any(_has_trial(c) for c in e.children())


def _split_linear(e):
    """(part containing the trial function, part without) of an expression that is affine in the trial function."""
    if not _has_trial(e):
        return Const(0.0), e
    if isinstance(e, Add):
        a1, a0 = _split_linear(e.a)
        b1, b0 = _split_linear(e.b)
        return a1 + b1, a0 + b0
    if isinstance(e, Neg):
        a1, a0 = _split_linear(e.a)
        return -a1, -a0
    if isinstance(e, Mul):
        if _has_trial(e.a) and _has_trial(e.b):
            raise ValueError("lhs/rhs: the form is not linear in the trial function")
        if _has_trial(e.a):
            a1, a0 = _split_linear(e.a)
            return a1 * e.b, a0 * e.b
        b1, b0 = _split_linear(e.b)
        return e.a * b1, e.a * b0
    if isinstance(e, Div):
        a1, a0 = _split_linear(e.a)
        return a1 / e.b, a0 / e.b
    return e, Const(0.0)


def lhs(form):
    return Form([Integral(_split_linear(i.integrand)[0], i.measure) for i in form.integrals])


def rhs(form):
    return Form([Integral(-_split_linear(i.integrand)[1], i.measure) for i in form.integrals])


# ---------------------------------------------------------------------------------------------------------------------
# degree estimation (UFL SumDegreeEstimator rules, the ones the FEDM forms exercise)
# ---------------------------------------------------------------------------------------------------------------------
def estimate_degree(e):
    if isinstance(e, Vec):
        return max(estimate_degree(e.x), estimate_degree(e.y))
    if isinstance(e, (Const, Param, Normal)):
        return 0
    if isinstance(e, Coord):
        return 1
    if isinstance(e, (Field, Arg)):
        return 1
    if isinstance(e, GradComp):
        return 0
    if isinstance(e, Interp):
        return e.degree
    if isinstance(e, Add):
        return max(estimate_degree(e.a), estimate_degree(e.b))
    if isinstance(e, (Mul, Div)):
        return estimate_degree(e.a) + estimate_degree(e.b)
    if isinstance(e, (Neg, Abs, Sign)):
        return estimate_degree(e.a)
    if isinstance(e, Pow):
        p = estimate_degree(e.a)
        if isinstance(e.e, Const) and getattr(e.e, "is_int", False) and e.e.value >= 0:
            return int(p * e.e.value)
        return p + 2
    if isinstance(e, Func):
        return estimate_degree(e.a) + 2
    raise NotImplementedError(type(e).__name__)


# ---------------------------------------------------------------------------------------------------------------------
# test-function coefficients: every integrand is linear and homogeneous in the test function
# ---------------------------------------------------------------------------------------------------------------------
def linear_coefficient(e, is_target):
    """d e / d leaf for the leaves selected by is_target(node) (an Arg or GradComp(Arg) of the test function); since the
    integrand is linear in those leaves this is exactly their coefficient."""
    if is_target(e):
        return Const(1.0)
    if isinstance(e, (Const, Param, Coord, Normal, Field, Arg, GradComp, Interp, Sign)):
        return Const(0.0)
    if isinstance(e, Add):
        return linear_coefficient(e.a, is_target) + linear_coefficient(e.b, is_target)
    if isinstance(e, Neg):
        return -linear_coefficient(e.a, is_target)
    if isinstance(e, Mul):
        return linear_coefficient(e.a, is_target) * e.b + e.a * linear_coefficient(e.b, is_target)
    if isinstance(e, Div):
        return linear_coefficient(e.a, is_target) / e.b
    if isinstance(e, (Pow, Func, Abs)):
        return Const(0.0)                  # test functions never sit inside nonlinear operators
    raise NotImplementedError(type(e).__name__)


def test_function_split(integrand, n_eq):
    """[(s_k, fx_k, fy_k)] with integrand = sum_k s_k v_k + fx_k dv_k/dx + fy_k dv_k/dy."""
    out = []
    for k in range(n_eq):
        s = linear_coefficient(integrand, lambda n, k=k: isinstance(n, Arg) and n.number == 0 and n.comp == k)
        fx = linear_coefficient(integrand, lambda n, k=k: isinstance(n, GradComp) and isinstance(n.arg, Arg)
                                and n.arg.number == 0 and n.arg.comp == k and n.i == 0)
        fy = linear_coefficient(integrand, lambda n, k=k: isinstance(n, GradComp) and isinstance(n.arg, Arg)
                                and n.arg.number == 0 and n.arg.comp == k and n.i == 1)
        out.append((s, fx, fy))
    return out
