"""Dry-run backend of the dolfin shim: executes a FEDM script WITHOUT solving anything, only to collect the forms it hands to the
backend and to pre-compile their CUDA row kernels (``__graft_entry__.build()`` -> ``tools/prebuild_jit.py``; nvcc cross-compiles
sm_100a without a GPU).  It is not a compute path: every solve returns immediately and leaves the unknowns at their initial
values, vectors are plain host arrays that only carry the scripts' own initialisation arithmetic."""
import numpy as np

from ..formcompiler import CompiledForm


class HostStore:
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
        s = HostStore(self.a.size)
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


class _Recorded:
    def __init__(self, so):
        self.so, self.last_result = so, None

    def solve_newton(self, u, **kw):
        # nothing is solved in a dry run; fresh (all-zero) targets, e.g. of project(), are filled with ones so that the
        # scripts' own post-processing (norm ratios) stays finite
        a = u.vector().store.a
        if not a.any():
            a[:] = 1.0
        return 1, True


class DryRunBackend:
    name = "dry-run (form collection + kernel pre-compilation)"

    def __init__(self):
        self.built = []

    def zeros(self, n, space=None):
        return HostStore(n)

    def compile(self, residual, u, bcs, quad_degree, ident_zeros=False, reuse=False):
        so = CompiledForm(residual, u, quad_degree).build()
        self.built.append(so.name)
        return _Recorded(so)
