"""Device-backed stand-ins for the dolfin objects the hot path touches (hot-path subset).

These are the objects ``fedm.functions.adaptive_solver`` manipulates: time-step ``Expression``s with a mutable
``time_step`` / ``t`` attribute, ``Function``s with ``.vector()`` / ``.assign()``, a ``FunctionAssigner`` between the
mixed space and per-field functions, ``norm``, and the ``NonlinearProblem`` / ``PETScSNESSolver`` pair.  Vectors are
torch CUDA tensors used as buffers; every operation is a libfedmx kernel.  UFL form *construction* is not emulated:
a problem is defined by choosing one of the compiled form families (``fedm_b200.hotpath.MODEL_*``), which is the role
FFC's JIT plays for the reference.
"""
import numpy as np
import torch

DOLFIN_EPS = 3.0e-16


class Expression:
    """``Expression("time_step", time_step=..., degree=0)`` and friends: a bag of mutable scalar parameters."""

    def __init__(self, cpp_code="", degree=0, **params):
        self.cpp_code, self.degree = cpp_code, degree
        for k, v in params.items():
            setattr(self, k, v)


class DeviceVector:
    """GenericVector look-alike over a contiguous float64 CUDA tensor."""

    def __init__(self, hot, tensor, n_comp):
        self.hot, self.t, self.n_comp = hot, tensor, n_comp

    def __getitem__(self, key):
        return self.t.cpu().numpy()[key]

    def __setitem__(self, key, value):
        if isinstance(key, slice) and key == slice(None):
            self.t.copy_(torch.as_tensor(np.broadcast_to(np.asarray(value, dtype=np.float64), self.t.shape).copy()))
        else:
            a = self.t.cpu().numpy()
            a[key] = value
            self.t.copy_(torch.from_numpy(a))

    def get_local(self):
        return self.t.cpu().numpy()

    def set_local(self, values):
        self[:] = values

    def update_ghost_values(self):
        self.hot.halo_exchange(self.t, self.n_comp)

    def apply(self, mode):
        pass

    def size(self):
        return self.t.numel()


class Function:
    def __init__(self, space, name=None):
        self.space = space
        self._vec = DeviceVector(space.hot, torch.zeros(space.dim, dtype=torch.float64, device=space.hot.device), space.n_comp)
        self._name = name or "f"

    def vector(self):
        return self._vec

    def assign(self, other):
        src = other.vector().t if isinstance(other, Function) else torch.as_tensor(other, dtype=torch.float64)
        self._vec.t.copy_(src if src.numel() == self._vec.t.numel() else src.expand_as(self._vec.t))

    def rename(self, name, label):
        self._name = name

    def function_space(self):
        return self.space


class FunctionSpace:
    """P1 space with n_comp interleaved components on the mesh of a HotPath (dof = n_comp*node + comp)."""

    def __init__(self, hot, n_comp):
        self.hot, self.n_comp = hot, n_comp
        self.dim = hot.n_vertices * n_comp

    def sub(self, i):
        return (self, i)


class FunctionAssigner:
    """``FunctionAssigner(receiving, assigning)`` between a mixed space and a list of scalar spaces (either direction)."""

    def __init__(self, receiving, assigning):
        self.to_list = isinstance(receiving, (list, tuple))

    def assign(self, receiving, assigning):
        if self.to_list:                                  # per-field <- mixed
            mixed = assigning
            hot, n = mixed.space.hot, mixed.space.n_comp
            for k, f in enumerate(receiving):
                hot.strided_copy(mixed.vector().t, n, k, f.vector().t, 1, 0, hot.n_vertices)
        else:                                             # mixed <- per-field
            mixed = receiving
            hot, n = mixed.space.hot, mixed.space.n_comp
            for k, f in enumerate(assigning):
                hot.strided_copy(f.vector().t, 1, 0, mixed.vector().t, n, k, hot.n_vertices)


def norm(vec, norm_type="l2"):
    return vec.hot.norm2(vec.t, vec.hot.n_owned * vec.n_comp)


class NonlinearProblem:
    pass


class PETScSNESSolver:
    """``PETScSNESSolver()``: ``parameters[...]`` + ``solve(problem, x)`` -> (iterations, converged); raises RuntimeError on
    non-convergence (dolfin's error_on_nonconvergence=True), which is what adaptive_solver's retry relies on."""

    def __init__(self):
        self.parameters = {"relative_tolerance": 1e-9, "absolute_tolerance": 1e-10, "solution_tolerance": 1e-16,
                           "maximum_iterations": 50, "linear_solver": "default", "preconditioner": "default",
                           "error_on_nonconvergence": True, "report": True, "method": "default", "line_search": "basic",
                           "krylov_solver": {"relative_tolerance": 1e-5, "absolute_tolerance": 1e-50, "maximum_iterations": 10000,
                                             "restart": 30}}
        self.last_result = None

    def solve(self, problem, x):
        p = self.parameters
        direct = p["linear_solver"] in ("mumps", "lu", "default", "superlu", "petsc")
        ks = p["krylov_solver"]
        res = problem.solve_newton(x, rtol=p["relative_tolerance"], atol=p["absolute_tolerance"], stol=p["solution_tolerance"],
                                   max_it=p["maximum_iterations"],
                                   ksp_rtol=1e-13 if direct else ks["relative_tolerance"],      # direct solvers: MUMPS-grade accuracy
                                   ksp_atol=ks["absolute_tolerance"], ksp_max_it=50000 if direct else ks["maximum_iterations"],
                                   ksp_restart=ks["restart"], raise_on_failure=p["error_on_nonconvergence"])
        self.last_result = res
        return res.iterations, res.reason > 0
