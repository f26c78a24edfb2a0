"""Tiny parser for the C++ expression strings of ``dolfin.Expression`` (SURVEY.md Appendix E lists every string the
reference uses): numbers, identifiers (user parameters, ``pi``, ``DOLFIN_EPS``), ``x[0]``/``x[1]``, ``+ - * /``, unary
minus, parentheses and the calls ``pow exp log sqrt fabs`` with an optional ``std::`` prefix.  The result is a small AST
that can be evaluated with numpy at points (``interpolate``, Dirichlet values) or lowered to UFL-lite (coefficients
that appear inside forms)."""
import math
import re

import numpy as np

_TOKEN = re.compile(r"\s*(?:(\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?)|((?:std::)?[A-Za-z_]\w*)|(.))")


def _tokens(text):
    out = []
    for num, name, op in _TOKEN.findall(text):
        if num:
            out.append(("num", float(num)))
        elif name:
            out.append(("id", name[5:] if name.startswith("std::") else name))
        elif op.strip():
            out.append(("op", op))
    return out


class _Parser:
    def __init__(self, text):
        self.t, self.i = _tokens(text), 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else ("end", None)

    def take(self, kind=None, val=None):
        k, v = self.peek()
        if (kind and k != kind) or (val is not None and v != val):
            raise SyntaxError(f"Expression string: unexpected token {v!r}")
        self.i += 1
        return v

    def expr(self):
        node = self.term()
        while self.peek() in (("op", "+"), ("op", "-")):
            op = self.take()
            node = (op, node, self.term())
        return node

    def term(self):
        node = self.unary()
        while self.peek() in (("op", "*"), ("op", "/")):
            op = self.take()
            node = (op, node, self.unary())
        return node

    def unary(self):
        if self.peek() == ("op", "-"):
            self.take()
            return ("neg", self.unary())
        if self.peek() == ("op", "+"):
            self.take()
            return self.unary()
        return self.atom()

    def atom(self):
        k, v = self.peek()
        if k == "num":
            self.take()
            return ("num", v)
        if k == "id":
            self.take()
            if self.peek() == ("op", "("):
                self.take()
                args = [self.expr()]
                while self.peek() == ("op", ","):
                    self.take()
                    args.append(self.expr())
                self.take("op", ")")
                return ("call", v, args)
            if self.peek() == ("op", "["):
                self.take()
                idx = int(self.take("num"))
                self.take("op", "]")
                return ("x", idx)
            return ("id", v)
        if (k, v) == ("op", "("):
            self.take()
            node = self.expr()
            self.take("op", ")")
            return node
        raise SyntaxError(f"Expression string: unexpected token {v!r}")


def parse(text):
    p = _Parser(text)
    node = p.expr()
    if p.peek()[0] != "end":
        raise SyntaxError(f"Expression string: trailing input in {text!r}")
    return node


_BUILTIN = {"pi": math.pi, "DOLFIN_EPS": 3.0e-16}


def evaluate(node, x, params):
    """numpy evaluation; x = (x0, x1) arrays, params: name -> float or array."""
    kind = node[0]
    if kind == "num":
        return node[1]
    if kind == "x":
        return x[node[1]]
    if kind == "id":
        if node[1] in params:
            return params[node[1]]
        if node[1] in _BUILTIN:
            return _BUILTIN[node[1]]
        raise NameError(f"Expression string: unknown name '{node[1]}'")
    if kind == "neg":
        return -evaluate(node[1], x, params)
    if kind in "+-*/":
        a, b = evaluate(node[1], x, params), evaluate(node[2], x, params)
        return a + b if kind == "+" else a - b if kind == "-" else a * b if kind == "*" else a / b
    if kind == "call":
        args = [evaluate(a, x, params) for a in node[2]]
        fn = {"pow": np.power, "exp": np.exp, "log": np.log, "sqrt": np.sqrt, "fabs": np.abs, "abs": np.abs}[node[1]]
        return fn(*args)
    raise ValueError(kind)


def to_ufl(node, params, ufl):
    """Lower to UFL-lite; params: name -> Expr (Param nodes for mutable scalars)."""
    kind = node[0]
    if kind == "num":
        return ufl.Const(node[1])
    if kind == "x":
        return ufl.Coord(node[1])
    if kind == "id":
        if node[1] in params:
            return params[node[1]]
        if node[1] in _BUILTIN:
            return ufl.Const(_BUILTIN[node[1]])
        raise NameError(f"Expression string: unknown name '{node[1]}'")
    if kind == "neg":
        return -to_ufl(node[1], params, ufl)
    if kind in "+-*/":
        a, b = to_ufl(node[1], params, ufl), to_ufl(node[2], params, ufl)
        return a + b if kind == "+" else a - b if kind == "-" else a * b if kind == "*" else a / b
    if kind == "call":
        args = [to_ufl(a, params, ufl) for a in node[2]]
        if node[1] == "pow":
            return ufl.Pow(args[0], args[1])
        if node[1] == "exp":
            return ufl.exp(args[0])
        if node[1] == "log":
            return ufl.ln(args[0])
        if node[1] == "sqrt":
            return ufl.sqrt(args[0])
        return abs(args[0])
    raise ValueError(kind)
