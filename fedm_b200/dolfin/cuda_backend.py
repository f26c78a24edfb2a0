"""CUDA backend of the dolfin shim: vectors are torch CUDA tensors (buffers only), forms are compiled by
``fedm_b200.formcompiler`` into row kernels and solved by libfedmx's device Newton-Krylov (``fedm_b200.hotpath``)."""
import ctypes as C

import numpy as np
import torch

from .. import hotpath as hp
from ..formcompiler import CompiledForm


class TorchStore:
    def __init__(self, n, device):
        self.t = torch.zeros(n, dtype=torch.float64, device=device)

    def numpy(self):
        return self.t.cpu().numpy()

    def set_numpy(self, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        self.t.copy_(torch.from_numpy(a if a.flags.writeable else a.copy()))

    def size(self):
        return self.t.numel()

    def norm2(self):
        return float(torch.linalg.vector_norm(self.t))

    def clone(self):
        s = TorchStore(0, self.t.device)
        s.t = self.t.clone()
        return s

    def axpy(self, alpha, other):
        self.t.add_(other.t, alpha=alpha)

    def add_scalar(self, c):
        self.t.add_(c)

    def scale(self, c):
        self.t.mul_(c)

    def copy_from(self, other):
        self.t.copy_(other.t)

    def update_ghosts(self, n_comp):
        pass

    def strided_copy_from(self, src, sstride, soff, dstride, doff, n):
        self.t[doff:doff + n * dstride:dstride] = src.t[soff:soff + n * sstride:sstride]


class CudaProblem:
    """One compiled residual form bound to a HotPath context."""

    def __init__(self, backend, residual, u, bcs, quad_degree, ident_zeros=False):
        self.u, self.bcs = u, list(bcs)
        self.form = CompiledForm(residual, u, quad_degree)
        so = self.form.build()
        space = u.space
        tri = space.mesh().tri
        self.hot = hot = hp.HotPath(backend.device_index)
        fd = self.form.facet_data
        if fd is not None:
            # the fan tables carry a 4-bit marker per exterior facet with 0 = interior edge: unmarked exterior facets (still
            # part of an un-indexed `ds`) travel as 15
            vals = np.asarray(fd.values, dtype=np.int32)
            if vals.size and vals.max() >= 15:
                raise NotImplementedError("facet markers must be < 15 on the device path")
            hot.set_mesh(space.n_comp, tri.coords, tri.cells, fd.facets, np.where(vals == 0, 15, vals).astype(np.int32))
        else:
            fv = tri.exterior_facets()
            hot.set_mesh(space.n_comp, tri.coords, tri.cells, fv, np.ones(fv.shape[0], dtype=np.int32))
        comps = {bc.comp for bc in self.bcs}
        if len(comps) > 1:
            raise NotImplementedError("Dirichlet conditions on more than one component")
        self.bc_comp = comps.pop() if comps else None
        if self.bcs:
            nodes = np.concatenate([bc.nodes for bc in self.bcs]).astype(np.int32)
            index = np.concatenate([np.full(bc.nodes.size, i) for i, bc in enumerate(self.bcs)]).astype(np.int32)
            hot.set_dirichlet(self.bc_comp, nodes, index)
            self.bc_nodes = nodes
        hot._ck(hot.lib.fedmx_plugin_load(hot.h, str(so).encode()))
        hot._ck(hot.lib.fedmx_ident_zeros_set(hot.h, 1 if ident_zeros else 0))
        self._keep = []
        for k, (fn, comp) in enumerate(self.form.fields):
            t = fn.vector().store.t
            self._keep.append(t)
            hot._ck(hot.lib.fedmx_plugin_set_field(hot.h, k, C.c_void_p(t.data_ptr()), fn.space.n_comp, comp))
        self.precond_ready = False
        self.last_result = None

    def _refresh(self):
        p = self.form.param_values()
        self.hot._ck(self.hot.lib.fedmx_model_set(self.hot.h, 100, p.size, p.ctypes.data_as(C.c_void_p)))
        self.hot.bc_values = np.array([bc.current_value() for bc in self.bcs], dtype=np.float64)

    def _setup_preconditioner(self, x):
        """Two-stage AMG on the Dirichlet-constrained (potential) component: its diagonal block is the state-independent
        stiffness operator in every FEDM form; scalar problems without Dirichlet rows (projections, ToF) use Jacobi."""
        self.precond_ready = True
        # multigrid needs an elliptic block: skip it when most rows are constrained (BoundaryGradient's anti-conditions)
        if self.bc_comp is None or 2 * len(set(self.bc_nodes.tolist())) > self.hot.n_owned:
            return
        self.hot.jacobian(x, 1.0, 1.0)
        self.hot.setup_two_stage_preconditioner(self.bc_comp, self.bc_nodes)

    def solve_newton(self, u, rtol=1e-9, atol=1e-10, stol=1e-16, max_it=50, linear_solver="mumps", raise_on_failure=True):
        x = u.vector().store.t
        self._refresh()
        if not self.precond_ready:
            self._setup_preconditioner(x)
        direct = linear_solver in ("mumps", "lu", "default", "superlu", "petsc", "umfpack")
        res = self.hot.newton(x, 1.0, 1.0, 0.0, rtol=rtol, atol=atol, stol=stol, max_it=max_it, ksp_restart=30,
                              ksp_rtol=1e-13 if direct else 1e-5, ksp_max_it=50000 if direct else 10000,
                              raise_on_failure=raise_on_failure)
        self.last_result = res
        return res.iterations, res.reason > 0

    # test hooks
    def residual(self, u):
        self._refresh()
        return self.hot.residual(u.vector().store.t, 1.0, 1.0).cpu().numpy()

    def jacobian_bsr(self, u):
        self._refresh()
        self.hot.jacobian(u.vector().store.t, 1.0, 1.0)
        return self.hot.matrix_bsr()


class CudaBackend:
    name = "cuda"

    def __init__(self, device_index=0):
        if not torch.cuda.is_available():
            raise RuntimeError("fedm_b200.dolfin: the CUDA backend needs a GPU (there is no CPU fallback)")
        self.device_index = device_index
        self.device = torch.device("cuda", device_index)

    def zeros(self, n, space=None):
        return TorchStore(n, self.device)

    def compile(self, residual, u, bcs, quad_degree, ident_zeros=False):
        return CudaProblem(self, residual, u, bcs, quad_degree, ident_zeros)
