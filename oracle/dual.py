"""Forward-mode automatic differentiation over numpy arrays (oracle -- test infrastructure only).

``Dual`` carries a value array ``v`` of shape (n,) and a dense derivative array ``d`` of shape
(n, nd).  ``Nest`` is a generic dual whose components can themselves be ``Dual``s; it is used
for *spatial* derivatives of composites of P1 fields (grad c = sum_a dc/dlambda_a grad phi_a),
which UFL's ``grad(D*exp(u))`` requires (fedm/functions.py:234 in the reference).  The Jacobians
produced this way are independent of the hand-derived Jacobians inside the CUDA kernels.
"""
import numpy as np


class Dual:
    __slots__ = ("v", "d")
    __array_priority__ = 1000
    __array_ufunc__ = None

    def __init__(self, v, d):
        self.v = v
        self.d = d

    @staticmethod
    def const(v, nd):
        v = np.asarray(v, dtype=np.float64)
        return Dual(v, np.zeros(v.shape + (nd,)))

    @staticmethod
    def seed(v, idx, nd):
        v = np.asarray(v, dtype=np.float64)
        d = np.zeros(v.shape + (nd,))
        d[..., idx] = 1.0
        return Dual(v, d)

    def _lift(self, o):
        return o if isinstance(o, Dual) else None

    def __add__(self, o):
        if isinstance(o, Nest):
            return NotImplemented
        if isinstance(o, Dual):
            return Dual(self.v + o.v, self.d + o.d)
        return Dual(self.v + o, self.d)

    __radd__ = __add__

    def __neg__(self):
        return Dual(-self.v, -self.d)

    def __sub__(self, o):
        if isinstance(o, Nest):
            return NotImplemented
        if isinstance(o, Dual):
            return Dual(self.v - o.v, self.d - o.d)
        return Dual(self.v - o, self.d)

    def __rsub__(self, o):
        return Dual(o - self.v, -self.d)

    def __mul__(self, o):
        if isinstance(o, Nest):
            return NotImplemented
        if isinstance(o, Dual):
            return Dual(self.v * o.v, self.d * o.v[..., None] + o.d * self.v[..., None])
        o = np.asarray(o, dtype=np.float64)
        return Dual(self.v * o, self.d * (o[..., None] if o.ndim else o))

    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, Nest):
            return NotImplemented
        if isinstance(o, Dual):
            q = self.v / o.v
            return Dual(q, (self.d - o.d * q[..., None]) / o.v[..., None])
        o = np.asarray(o, dtype=np.float64)
        return Dual(self.v / o, self.d / (o[..., None] if o.ndim else o))

    def __rtruediv__(self, o):
        q = o / self.v
        return Dual(q, -self.d * (q / self.v)[..., None])

    def __pow__(self, p):
        # real (float) exponent only, as in the reference's coefficient strings
        vp = self.v ** p
        return Dual(vp, self.d * (p * self.v ** (p - 1.0))[..., None])


class Nest:
    """Generic dual number: value ``v`` plus a list ``d`` of derivative components of the same type."""
    __slots__ = ("v", "d")
    __array_priority__ = 2000
    __array_ufunc__ = None

    def __init__(self, v, d):
        self.v = v
        self.d = list(d)

    def __add__(self, o):
        if isinstance(o, Nest):
            return Nest(self.v + o.v, [a + b for a, b in zip(self.d, o.d)])
        return Nest(self.v + o, self.d)

    __radd__ = __add__

    def __neg__(self):
        return Nest(-self.v, [-a for a in self.d])

    def __sub__(self, o):
        if isinstance(o, Nest):
            return Nest(self.v - o.v, [a - b for a, b in zip(self.d, o.d)])
        return Nest(self.v - o, self.d)

    def __rsub__(self, o):
        return Nest(o - self.v, [-a for a in self.d])

    def __mul__(self, o):
        if isinstance(o, Nest):
            return Nest(self.v * o.v, [a * o.v + self.v * b for a, b in zip(self.d, o.d)])
        return Nest(self.v * o, [a * o for a in self.d])

    __rmul__ = __mul__

    def __truediv__(self, o):
        if isinstance(o, Nest):
            q = self.v / o.v
            return Nest(q, [(a - q * b) / o.v for a, b in zip(self.d, o.d)])
        return Nest(self.v / o, [a / o for a in self.d])

    def __rtruediv__(self, o):
        q = o / self.v
        return Nest(q, [-(q / self.v) * a for a in self.d])

    def __pow__(self, p):
        vp = self.v ** p
        f = p * self.v ** (p - 1.0)
        return Nest(vp, [f * a for a in self.d])


def val(x):
    """Plain numpy value of a (possibly nested) dual."""
    while isinstance(x, (Dual, Nest)):
        x = x.v
    return x


def exp(x):
    if isinstance(x, Nest):
        e = exp(x.v)
        return Nest(e, [e * a for a in x.d])
    if isinstance(x, Dual):
        e = np.exp(x.v)
        return Dual(e, x.d * e[..., None])
    return np.exp(x)


def log(x):
    if isinstance(x, Nest):
        return Nest(log(x.v), [a / x.v for a in x.d])
    if isinstance(x, Dual):
        return Dual(np.log(x.v), x.d / x.v[..., None])
    return np.log(x)


def sqrt(x):
    if isinstance(x, Nest):
        s = sqrt(x.v)
        return Nest(s, [a / (2.0 * s) for a in x.d])
    if isinstance(x, Dual):
        s = np.sqrt(x.v)
        with np.errstate(divide="ignore", invalid="ignore"):
            return Dual(s, x.d / (2.0 * s)[..., None])
    return np.sqrt(x)


def sign(x):
    return np.sign(val(x))


def absolute(x):
    """|x| with d|x|/dx = sign(x) (0 at 0), as UFL differentiates abs()."""
    if isinstance(x, (Dual, Nest)):
        return x * sign(x)
    return np.abs(x)
