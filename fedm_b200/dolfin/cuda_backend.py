"""CUDA backend of the dolfin shim: vectors are torch CUDA tensors (buffers only), forms are compiled by
``fedm_b200.formcompiler`` into row kernels and solved by libfedmx's device Newton-Krylov (``fedm_b200.hotpath``)."""
import ctypes as C
import os

import numpy as np
import torch

from .. import hotpath as hp
from ..formcompiler import CompiledForm


class TorchStore:
    """Device vector of the shim.  torch only owns the buffer and does the BLAS-1 style updates (axpy, scale); everything the
    reference's hot path does on dof vectors -- df.norm (fedm/functions.py:1062-1064), FunctionAssigner.assign (:1050) and
    np.interp (:627-636) -- goes through the library's kernels (fixed-order reductions, bit-exact interpolation)."""
    util = None                 # utility HotPath context of the backend (set by CudaBackend)

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
        return TorchStore.util.norm2(self.t, self.t.numel())           # fedmx_norm2: two-stage fixed-order reduction

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
        TorchStore.util.strided_copy(src.t, sstride, soff, self.t, dstride, doff, n)       # fedmx_strided_copy

    def interp_from(self, src, kx, ky, divisor=1.0):
        """self = np.interp(src, kx, ky) / divisor on the device (fedmx_interp_table, bit-exact with numpy)."""
        key = (id(kx), id(ky), len(kx))
        tabs = TorchStore._tables.get(key)
        if tabs is None or tabs[0] is not kx:
            dev = self.t.device
            tabs = (kx, torch.from_numpy(np.ascontiguousarray(kx, dtype=np.float64)).to(dev),
                    torch.from_numpy(np.ascontiguousarray(ky, dtype=np.float64)).to(dev))
            if len(TorchStore._tables) > 256:
                TorchStore._tables.clear()
            TorchStore._tables[key] = tabs
        TorchStore.util.interp_table(src.t, tabs[1], tabs[2], self.t, divisor=float(divisor))

    _tables = {}


class CudaProblem:
    """One compiled residual form bound to a HotPath context."""

    def __init__(self, backend, residual, u, bcs, quad_degree, ident_zeros=False, form=None):
        self.u, self.bcs = u, list(bcs)
        self.form = form if form is not None else CompiledForm(residual, u, quad_degree)
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
            fv, fm = fd.facets, np.where(vals == 0, 15, vals).astype(np.int32)
        else:
            fv = tri.exterior_facets()
            fm = np.ones(fv.shape[0], dtype=np.int32)
        # One process per GPU (`torchrun` = the reference's `mpirun -np N`, README.md:63-67): the shim's vectors stay
        # replicated in global vertex order on every rank, the SOLVE is partitioned like dolfin's MPI decomposition -- each
        # rank assembles and solves the rows of its own vertices (+ one ghost layer, halo exchange / allreduce in libfedmx)
        # and the result is gathered back into the replicated vector.
        self.local = None
        world, rank = backend.world, backend.rank
        if world > 1:
            from ..partition import build_local_mesh, partition_owner
            owner = partition_owner(tri.coords, world, os.environ.get("FEDM_B200_PARTITION", "auto"))
            self.local = lm = build_local_mesh(tri.coords, tri.cells, fv, fm, owner, rank, world)
            hot.set_mesh(space.n_comp, lm.coords, lm.cells, lm.facet_vertices, lm.facet_markers, n_owned=lm.n_owned)
            backend.comm_setup(hot, lm)
            self.l2g = torch.from_numpy(lm.local_to_global).to(backend.device)
            g2l = lm.global_to_local()
        else:
            hot.set_mesh(space.n_comp, tri.coords, tri.cells, fv, fm)
        comps = {bc.comp for bc in self.bcs}
        if len(comps) > 1:
            raise NotImplementedError("Dirichlet conditions on more than one component")
        self.bc_comp = comps.pop() if comps else None
        if self.bcs:
            nodes = np.concatenate([bc.nodes for bc in self.bcs]).astype(np.int64)
            index = np.concatenate([np.full(bc.nodes.size, i) for i, bc in enumerate(self.bcs)]).astype(np.int32)
            if self.local is not None:                     # constrained nodes this rank sees (owned or ghost), local ids
                keep = g2l[nodes] >= 0
                nodes, index = g2l[nodes[keep]], index[keep]
            nodes = nodes.astype(np.int32)
            hot.set_dirichlet(self.bc_comp, nodes, index)
            self.bc_nodes = nodes
        hot._ck(hot.lib.fedmx_plugin_load(hot.h, str(so).encode()))
        hot._ck(hot.lib.fedmx_ident_zeros_set(hot.h, 1 if ident_zeros else 0))
        self._bind_fields()
        self.precond_ready = False
        self.last_result = None

    def _bind_fields(self):
        hot = self.hot
        self._keep = []
        self._local_fields = {}                    # id(fn) -> (global tensor, local tensor): partitioned runs read local copies
        for k, (fn, comp) in enumerate(self.form.fields):
            t = fn.vector().store.t
            if self.local is not None:
                if id(fn) not in self._local_fields:
                    loc = torch.empty(self.local.num_vertices * fn.space.n_comp, dtype=torch.float64, device=t.device)
                    self._local_fields[id(fn)] = (t, loc, fn.space.n_comp)
                t = self._local_fields[id(fn)][1]
            self._keep.append(t)
            hot._ck(hot.lib.fedmx_plugin_set_field(hot.h, k, C.c_void_p(t.data_ptr()), fn.space.n_comp, comp))

    def _to_local(self, glob, loc, n_comp):
        loc.view(-1, n_comp).copy_(glob.view(-1, n_comp)[self.l2g])

    def _refresh_local_fields(self):
        for glob, loc, nc in self._local_fields.values():
            self._to_local(glob, loc, nc)

    def _gather_solution(self, x_glob, x_loc, n_comp):
        """owned rows of every rank -> the replicated global vector (what `Function.vector()` holds on all ranks)."""
        import torch.distributed as dist
        no = self.local.n_owned
        x_glob.zero_()
        x_glob.view(-1, n_comp)[self.l2g[:no]] = x_loc.view(-1, n_comp)[:no]
        dist.all_reduce(x_glob)

    def rebind(self, form, u, bcs):
        """Re-use of a compiled problem for a structurally identical form (same generated kernel, same mesh, same constrained
        nodes) built from other Function / Constant objects: only the coefficient pointers and parameter sources change."""
        self.form, self.u, self.bcs = form, u, list(bcs)
        self._bind_fields()

    def close(self):
        self.hot.close()

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
        self.hot.setup_two_stage_preconditioner(self.bc_comp, self.bc_nodes, local_mesh=self.local)

    def solve_newton(self, u, rtol=1e-9, atol=1e-10, stol=1e-16, max_it=50, linear_solver="mumps", raise_on_failure=True):
        x = xg = u.vector().store.t
        nc = u.space.n_comp
        self._refresh()
        if self.local is not None:
            if getattr(self, "_x_local", None) is None:
                self._x_local = torch.empty(self.local.num_vertices * nc, dtype=torch.float64, device=xg.device)
            x = self._x_local
            self._to_local(xg, x, nc)
            self._refresh_local_fields()
        if not self.precond_ready:
            self._setup_preconditioner(x)
        direct = linear_solver in ("mumps", "lu", "default", "superlu", "petsc", "umfpack")
        try:
            res = self.hot.newton(x, 1.0, 1.0, 0.0, rtol=rtol, atol=atol, stol=stol, max_it=max_it, ksp_restart=30,
                                  ksp_rtol=1e-13 if direct else 1e-5, ksp_max_it=50000 if direct else 10000,
                                  raise_on_failure=raise_on_failure)
        finally:
            if self.local is not None:                    # also after a failed attempt: every rank takes the same path
                self._gather_solution(xg, x, nc)
        self.last_result = res
        return res.iterations, res.reason > 0

    # test hooks
    def residual(self, u):
        self._refresh()
        if self.local is not None:
            raise NotImplementedError("test hook: un-partitioned runs only")
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
        # one process per GPU under torchrun: the device is the local rank's
        self.world, self.rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
        if self.world > 1:
            device_index = int(os.environ.get("LOCAL_RANK", str(device_index)))
            from . import MPI
            MPI._dist()                                    # creates the process group (NCCL)
        self.device_index = device_index
        self.device = torch.device("cuda", device_index)
        torch.cuda.set_device(self.device)
        # utility context for vector kernels that need no mesh of their own (norms, strided copies, table interpolation)
        util = hp.HotPath(device_index)
        util.set_mesh(1, np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]), np.array([[0, 1, 2]], dtype=np.int32))
        TorchStore.util = self.util = util
        self._cache = {}                 # structurally identical linear problems (project / Poisson_solver in a time loop)

    def zeros(self, n, space=None):
        return TorchStore(n, self.device)

    def comm_setup(self, hot, lm):
        """One NCCL communicator per context (rank 0 creates the id, torch.distributed carries its 128 bytes) + halo plan."""
        import torch.distributed as dist
        buf = C.create_string_buffer(128)
        if self.rank == 0:
            hot._ck(hot.lib.fedmx_comm_unique_id(buf))
        t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=self.device)
        dist.broadcast(t, 0)
        hot.comm_init(self.world, self.rank, bytes(t.cpu().tolist()))
        hot.set_halo(lm.neighbors, lm.send_ptr, lm.send_nodes, lm.recv_ptr, lm.recv_nodes)

    def compile(self, residual, u, bcs, quad_degree, ident_zeros=False, reuse=False):
        """reuse=True (solve / project / Poisson_solver ...): look the problem up by the hash of the generated kernel, the
        mesh, the facet markers and the constrained nodes, so that calls inside a time loop do not pay code generation, mesh
        upload, pattern and multigrid set-up again (and do not churn device memory)."""
        if not reuse:
            return CudaProblem(self, residual, u, bcs, quad_degree, ident_zeros)
        import hashlib
        form = CompiledForm(residual, u, quad_degree)
        h = hashlib.sha1(form.key.encode())
        tri = u.space.mesh().tri
        h.update(str((id(tri), tri.num_vertices, tri.num_cells, bool(ident_zeros))).encode())
        if form.facet_data is not None:
            h.update(np.asarray(form.facet_data.values, dtype=np.int32).tobytes())
        for bc in bcs:
            h.update(str(bc.comp).encode())
            h.update(np.asarray(bc.nodes, dtype=np.int32).tobytes())
        key = h.hexdigest()
        prob = self._cache.get(key)
        if prob is None:
            if len(self._cache) >= 16:                      # bounded: drop the oldest entry and its device context
                self._cache.pop(next(iter(self._cache))).close()
            prob = self._cache[key] = CudaProblem(self, residual, u, bcs, quad_degree, ident_zeros, form=form)
        else:
            prob.rebind(form, u, bcs)
        return prob
