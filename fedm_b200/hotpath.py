"""Device-resident nonlinear problem: the Python face of libfedmx (ctypes; torch tensors as buffers only).

Plays the role of dolfin's ``NonlinearProblem`` + ``PETScSNESSolver`` pair for the FEDM form family:
``HotPath.residual`` / ``HotPath.jacobian`` are ``Problem.F`` / ``Problem.J`` (fedm/functions.py:188-202),
``HotPath.newton`` is ``nonlinear_solver.solve(problem, x)`` (fedm/functions.py:1047) including its
"raise on non-convergence" behaviour, ``HotPath.error_norm`` the estimator of fedm/functions.py:1062-1064.
"""
import ctypes as C

import numpy as np
import torch

from . import lib as _l

MODEL_STREAMER, MODEL_TOF, MODEL_GLOW, MODEL_POISSON, MODEL_PROJECTION = 1, 2, 3, 4, 5

SNES_REASONS = {2: "CONVERGED_FNORM_ABS", 3: "CONVERGED_FNORM_RELATIVE", 4: "CONVERGED_SNORM_RELATIVE",
                -3: "DIVERGED_LINEAR_SOLVE", -4: "DIVERGED_FNORM_NAN", -5: "DIVERGED_MAX_IT"}


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        if not t.is_cuda or t.dtype != torch.float64 or not t.is_contiguous():
            raise ValueError("expected a contiguous float64 CUDA tensor")
        return C.c_void_p(t.data_ptr())
    raise TypeError(type(t))


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class HotPath:
    def __init__(self, device=0):
        if not torch.cuda.is_available():
            raise RuntimeError("fedm_b200.HotPath needs a CUDA device (there is no CPU fallback)")
        self.lib = _l.load()
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        h = C.c_void_p()
        rc = self.lib.fedmx_ctx_create(device, C.byref(h))
        if rc != 0:
            raise _l.FedmxError(f"fedmx_ctx_create failed ({rc})")
        self.h = h
        self.n_eq = 0
        self.n_vertices = self.n_owned = 0
        self.bc_values = np.zeros(0)
        self.fields = None
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.lib.fedmx_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise _l.FedmxError(f"libfedmx error {rc}: {self.lib.fedmx_last_error(self.h).decode()}")

    # ---- set-up -------------------------------------------------------------------------------------
    def set_mesh(self, n_eq, coords, cells, facet_vertices=None, facet_markers=None, n_owned=None):
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        nf = 0 if facet_vertices is None else len(facet_vertices)
        fv = np.ascontiguousarray(facet_vertices if nf else np.zeros((0, 2)), dtype=np.int32)
        fm = np.ascontiguousarray(facet_markers if nf else np.zeros(0), dtype=np.int32)
        n_owned = coords.shape[0] if n_owned is None else int(n_owned)
        self._ck(self.lib.fedmx_mesh_set(self.h, n_eq, coords.shape[0], _np_ptr(coords), cells.shape[0], _np_ptr(cells),
                                         nf, _np_ptr(fv), _np_ptr(fm), n_owned))
        self.n_eq, self.n_vertices, self.n_owned = n_eq, coords.shape[0], n_owned

    def set_dirichlet(self, comp, nodes, value_index):
        nodes = np.ascontiguousarray(nodes, dtype=np.int32)
        vi = np.ascontiguousarray(value_index, dtype=np.int32)
        self._ck(self.lib.fedmx_dirichlet_set(self.h, comp, nodes.size, _np_ptr(nodes), _np_ptr(vi)))

    def set_model(self, model_id, params, n_fields):
        p = np.ascontiguousarray(params, dtype=np.float64)
        self._ck(self.lib.fedmx_model_set(self.h, model_id, p.size, _np_ptr(p)))
        self.fields = torch.zeros((self.n_vertices, max(n_fields, 1)), dtype=torch.float64, device=self.device)
        self.n_fields = n_fields
        self._ck(self.lib.fedmx_fields_bind(self.h, n_fields, _ptr(self.fields)))

    def pattern(self):
        nr, nz = C.c_int64(), C.c_int64()
        self._ck(self.lib.fedmx_pattern_get(self.h, C.byref(nr), C.byref(nz), None, None))
        rowptr = np.empty(nr.value + 1, dtype=np.int32)
        colidx = np.empty(nz.value, dtype=np.int32)
        self._ck(self.lib.fedmx_pattern_get(self.h, C.byref(nr), C.byref(nz), _np_ptr(rowptr), _np_ptr(colidx)))
        return rowptr, colidx

    def dofmap_fenics(self):
        """dof -> vertex map of dolfin's serial P1 numbering for this context's mesh (fedmx_dofmap_fenics)."""
        d2v = np.empty(self.n_vertices, dtype=np.int32)
        self._ck(self.lib.fedmx_dofmap_fenics(self.h, _np_ptr(d2v)))
        return d2v

    def pattern_fenics(self):
        """dolfin's scalar SparsityPattern in dolfin's own dof numbering (rows/columns n_eq*dof + comp): (indptr, indices)."""
        nr, nz = C.c_int64(), C.c_int64()
        self._ck(self.lib.fedmx_pattern_get_fenics(self.h, C.byref(nr), C.byref(nz), None, None))
        indptr = np.empty(nr.value + 1, dtype=np.int64)
        indices = np.empty(nz.value, dtype=np.int32)
        self._ck(self.lib.fedmx_pattern_get_fenics(self.h, C.byref(nr), C.byref(nz), _np_ptr(indptr), _np_ptr(indices)))
        return indptr, indices

    def matrix_bsr(self):
        rowptr, colidx = self.pattern()
        vals = np.empty((colidx.size, self.n_eq, self.n_eq))
        self._ck(self.lib.fedmx_matrix_get_bsr(self.h, _np_ptr(vals)))
        return rowptr, colidx, vals

    def matrix_component(self, i, j):
        """Entry (i,j) of every stored block as a scipy CSR matrix (owned rows x local columns)."""
        import scipy.sparse as sp
        rowptr, colidx = self.pattern()
        vals = np.empty(colidx.size)
        self._ck(self.lib.fedmx_matrix_get_component(self.h, i, j, _np_ptr(vals)))
        return sp.csr_matrix((vals, colidx, rowptr), shape=(self.n_owned, self.n_vertices))

    def setup_two_stage_preconditioner(self, comp, bc_nodes, nu=1, theta=0.0, local_mesh=None, hierarchy=None):
        """Builds the smoothed-aggregation hierarchy of the (state-independent) comp-comp block of the Jacobian that
        is currently assembled on the device, uploads it and switches GMRES to the two-stage preconditioner.
        Partitioned runs (local_mesh given): the owned rows of all ranks are gathered into the global block, every
        rank builds the identical global hierarchy, keeps its own slice of the finest level (halo exchange per
        smoothing step) and a replica of all coarser levels (the restricted residual is all-reduced)."""
        import os
        import scipy.sparse as sp
        from . import amg
        nu = int(os.environ.get("FEDMX_AMG_NU", nu))
        # hierarchy=(levels, coarse_inv): re-use the host hierarchy of another context with the SAME operator on the same mesh
        # (e.g. the set-up Poisson problem and the potential block of the coupled system: both are 2 pi r grad u . grad v)
        Kloc = self.matrix_component(comp, comp) if hierarchy is None else None
        bc = np.asarray(bc_nodes, dtype=np.int64)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        if local_mesh is None or local_mesh.world == 1:
            if hierarchy is None:
                K = amg.symmetric_dirichlet(Kloc[:, :self.n_owned], bc[bc < self.n_owned])
                levels, coarse_inv = amg.build_hierarchy(K, theta=theta, coarse_size=int(os.environ.get('FEDMX_AMG_COARSE', 600)))
            else:
                levels, coarse_inv = hierarchy
            fine_ops = None
        elif hierarchy is not None:
            lm = local_mesh
            l2g = lm.local_to_global
            levels, coarse_inv = hierarchy
            owned_g = l2g[:self.n_owned]
            g2l = lm.global_to_local()
            A0 = levels[0]["A"][owned_g].tocsr()
            A0 = sp.csr_matrix((A0.data, g2l[A0.indices], A0.indptr), shape=(self.n_owned, self.n_vertices))
            fine_ops = (A0, levels[0]["P"][owned_g].tocsr(), levels[0]["R"][:, owned_g].tocsr(), levels[0]["dinv"][owned_g],
                        levels[0]["A"].diagonal()[owned_g], np.diff(levels[0]["A"].indptr)[owned_g] == 1)
        else:
            import torch.distributed as dist
            lm = local_mesh
            l2g = lm.local_to_global
            coo = Kloc.tocoo()
            trip = torch.from_numpy(np.stack([l2g[coo.row].astype(np.float64), l2g[coo.col].astype(np.float64), coo.data])).to(self.device)
            sizes = [torch.zeros(1, dtype=torch.int64, device=self.device) for _ in range(lm.world)]
            dist.all_gather(sizes, torch.tensor([trip.shape[1]], dtype=torch.int64, device=self.device))
            mx = int(max(int(t.item()) for t in sizes))
            pad = torch.zeros((3, mx), dtype=torch.float64, device=self.device)
            pad[:, :trip.shape[1]] = trip
            bufs = [torch.empty_like(pad) for _ in range(lm.world)]
            dist.all_gather(bufs, pad)
            parts = [b[:, :int(sz.item())].cpu().numpy() for b, sz in zip(bufs, sizes)]
            allt = np.concatenate(parts, axis=1)
            ng = lm.n_global_vertices
            Kg = sp.csr_matrix((allt[2], (allt[0].astype(np.int64), allt[1].astype(np.int64))), shape=(ng, ng))
            bcg = torch.zeros(ng, dtype=torch.float64, device=self.device)
            bcg[torch.from_numpy(l2g[bc[bc < self.n_owned]]).to(self.device)] = 1.0
            dist.all_reduce(bcg)
            Kg = amg.symmetric_dirichlet(Kg, np.nonzero(bcg.cpu().numpy() > 0)[0])
            levels, coarse_inv = amg.build_hierarchy(Kg, theta=theta, coarse_size=int(os.environ.get('FEDMX_AMG_COARSE', 600)))
            owned_g = l2g[:self.n_owned]
            g2l = lm.global_to_local()
            A0 = levels[0]["A"][owned_g].tocsr()
            A0 = sp.csr_matrix((A0.data, g2l[A0.indices], A0.indptr), shape=(self.n_owned, self.n_vertices))
            assert (A0.indices >= 0).all()
            P0 = levels[0]["P"][owned_g].tocsr()
            R0 = levels[0]["R"][:, owned_g].tocsr()          # coarse rows x this rank's fine columns (local ids = 0..n_owned)
            fine_ops = (A0, P0, R0, levels[0]["dinv"][owned_g], levels[0]["A"].diagonal()[owned_g],
                        np.diff(levels[0]["A"].indptr)[owned_g] == 1)
        self._ck(self.lib.fedmx_amg_reset(self.h))
        for l, lv in enumerate(levels):
            A = lv["A"]
            dinv = lv["dinv"].copy()
            isolated = np.diff(A.indptr) == 1            # Dirichlet rows: make the damped sweep exact
            diagA = A.diagonal()
            if l == 0 and fine_ops is not None:
                A, P, R, dinv, diagA, isolated = fine_ops
                dinv = dinv.copy()
            dinv[isolated] = 1.0 / (lv["omega"] * diagA[isolated])
            if "P" in lv:
                if not (l == 0 and fine_ops is not None):
                    P, R = lv["P"], lv["R"]
                arrs = [i32(A.indptr), i32(A.indices), f64(A.data), i32(P.indptr), i32(P.indices), f64(P.data),
                        i32(R.indptr), i32(R.indices), f64(R.data)]
                self._ck(self.lib.fedmx_amg_add_level(self.h, A.shape[0], _np_ptr(f64(dinv)), float(lv["omega"]),
                                                      _np_ptr(arrs[0]), _np_ptr(arrs[1]), _np_ptr(arrs[2]), P.shape[1],
                                                      _np_ptr(arrs[3]), _np_ptr(arrs[4]), _np_ptr(arrs[5]),
                                                      _np_ptr(arrs[6]), _np_ptr(arrs[7]), _np_ptr(arrs[8])))
            else:
                self._ck(self.lib.fedmx_amg_add_level(self.h, A.shape[0], _np_ptr(f64(dinv)), float(lv["omega"]),
                                                      None, None, None, 0, None, None, None, None, None, None))
        ci = f64(coarse_inv)
        self._ck(self.lib.fedmx_amg_set_coarse(self.h, ci.shape[0], _np_ptr(ci)))
        # finest-level cycle: V(1,1) on one GPU; partitioned runs default to V(0,1) with a prolongation that covers the ghost
        # rows, which removes both halo exchanges of the cycle (FEDMX_MG_V01=0/1 overrides)
        v01 = os.environ.get("FEDMX_MG_V01")
        v01 = (fine_ops is not None) if v01 is None else v01 == "1"
        if v01 and len(levels) > 1:
            if fine_ops is not None:
                Pg = levels[0]["P"][local_mesh.local_to_global].tocsr()            # owned + ghost rows, coarse ids are global
                self._ck(self.lib.fedmx_amg_set_fine_prolongation(self.h, Pg.shape[0], _np_ptr(i32(Pg.indptr)), _np_ptr(i32(Pg.indices)),
                                                                  _np_ptr(f64(Pg.data)), 1))
            else:
                self._ck(self.lib.fedmx_amg_set_fine_prolongation(self.h, 0, None, None, None, 1))
        self._ck(self.lib.fedmx_precond_set(self.h, 1, comp, nu))
        self.amg_levels = [lv["A"].shape[0] for lv in levels]
        self.amg_host = (levels, coarse_inv)             # kept for tests (numpy reference cycle)
        return self.amg_levels

    def pc_apply(self, r, z=None):
        """z = M^-1 r with the current preconditioner (PCApply)."""
        z = self.new_vector() if z is None else z
        self._ck(self.lib.fedmx_pc_apply(self.h, _ptr(r), _ptr(z)))
        return z

    def use_block_jacobi(self):
        self._ck(self.lib.fedmx_precond_set(self.h, 0, 0, 1))

    def use_ilu0(self):
        """PETSc's default PC of the reference test's KSP: ILU(0) on this rank's diagonal block (multicolour ordering)."""
        self._ck(self.lib.fedmx_precond_set(self.h, 2, 0, 1))

    def ilu0_info(self):
        """(number of colours, stored blocks, elimination order: node of every permuted row)."""
        nc, nb = C.c_int(), C.c_int64()
        order = np.empty(self.n_owned, dtype=np.int32)
        self._ck(self.lib.fedmx_ilu0_info(self.h, C.byref(nc), C.byref(nb), _np_ptr(order)))
        return nc.value, nb.value, order

    def ilu0_solve(self, r, z=None):
        z = self.new_vector() if z is None else z
        self._ck(self.lib.fedmx_ilu0_solve(self.h, _ptr(r), _ptr(z)))
        return z

    def new_vector(self):
        return torch.zeros(self.n_vertices * self.n_eq, dtype=torch.float64, device=self.device)

    # ---- per-step calls ------------------------------------------------------------------------------
    def _bc(self):
        b = np.ascontiguousarray(self.bc_values, dtype=np.float64)
        self._keep = [b]
        return _np_ptr(b) if b.size else None, int(b.size)

    def bdf2_history(self, n_hist, x_old, x_old1, dt, dt_old, field_offset=0):
        self._ck(self.lib.fedmx_bdf2_history(self.h, n_hist, _ptr(x_old), _ptr(x_old1), dt, dt_old, _ptr(self.fields),
                                             self.fields.shape[1], field_offset))

    def residual(self, x, dt, dt_old, t=0.0, out=None):
        out = self.new_vector() if out is None else out
        p, n = self._bc()
        self._ck(self.lib.fedmx_assemble_residual(self.h, _ptr(x), dt, dt_old, t, p, n, _ptr(out)))
        return out

    def jacobian(self, x, dt, dt_old, t=0.0, residual_out=None):
        p, n = self._bc()
        self._ck(self.lib.fedmx_assemble_jacobian(self.h, _ptr(x), dt, dt_old, t, p, n, _ptr(residual_out)))

    def spmv(self, x, y=None):
        y = self.new_vector() if y is None else y
        self._ck(self.lib.fedmx_spmv(self.h, _ptr(x), _ptr(y)))
        return y

    def norm2(self, x, n=None):
        out = C.c_double()
        self._ck(self.lib.fedmx_norm2(self.h, _ptr(x), self.n_owned * self.n_eq if n is None else n, C.byref(out)))
        return out.value

    def gmres(self, b, x=None, restart=30, rtol=1e-5, atol=1e-50, max_it=10000):
        x = self.new_vector() if x is None else x
        its, conv, rn = C.c_int(), C.c_int(), C.c_double()
        self._ck(self.lib.fedmx_gmres(self.h, _ptr(b), _ptr(x), restart, rtol, atol, max_it, C.byref(its), C.byref(rn),
                                      C.byref(conv)))
        return x, its.value, bool(conv.value), rn.value

    def newton(self, x, dt, dt_old, t=0.0, rtol=1e-9, atol=1e-10, stol=1e-16, max_it=50, ksp_restart=30,
               ksp_rtol=1e-5, ksp_atol=1e-50, ksp_max_it=10000, raise_on_failure=True):
        """x is updated in place.  Returns the NewtonResult; raises RuntimeError on divergence like dolfin's
        PETScSNESSolver with error_on_nonconvergence=True (that is what fedm's adaptive_solver catches)."""
        prm = _l.NewtonParams(rtol, atol, stol, max_it, ksp_restart, ksp_rtol, ksp_atol, ksp_max_it)
        res = _l.NewtonResult()
        p, n = self._bc()
        self._ck(self.lib.fedmx_newton_solve(self.h, _ptr(x), dt, dt_old, t, p, n, C.byref(prm), C.byref(res)))
        if res.reason < 0 and raise_on_failure:
            raise RuntimeError(f"Newton solver did not converge: {SNES_REASONS.get(res.reason, res.reason)} after "
                               f"{res.iterations} iterations (|F| = {res.fnorm:g})")
        return res

    def error_norm(self, x_new, x_old, comp, eps=3.0e-16):
        out = C.c_double()
        self._ck(self.lib.fedmx_error_norm(self.h, _ptr(x_new), _ptr(x_old), comp, eps, C.byref(out)))
        return out.value

    def error_norm_strided(self, a, b, stride=1, offset=0, n=None, eps=3.0e-16):
        out = C.c_double()
        n = self.n_owned if n is None else n
        self._ck(self.lib.fedmx_error_norm_strided(self.h, _ptr(a), _ptr(b), stride, offset, n, eps, C.byref(out)))
        return out.value

    def interp_table(self, x, kx, ky, out, divisor=1.0, x_stride=1, out_stride=1, n=None):
        n = (x.numel() // x_stride) if n is None else n
        self._ck(self.lib.fedmx_interp_table(self.h, _ptr(x), x_stride, n, _ptr(kx), _ptr(ky), kx.numel(), divisor,
                                             _ptr(out), out_stride))
        return out

    def strided_copy(self, src, src_stride, src_off, dst, dst_stride, dst_off, n):
        self._ck(self.lib.fedmx_strided_copy(self.h, _ptr(src), src_stride, src_off, _ptr(dst), dst_stride, dst_off, n))

    def strided_scale(self, src, src_stride, src_off, dst, dst_stride, dst_off, n, mul, div=1.0):
        self._ck(self.lib.fedmx_strided_scale(self.h, _ptr(src), src_stride, src_off, _ptr(dst), dst_stride, dst_off, n, mul, div))

    def exp_diff(self, x, n_eq, comp_a, comp_b, dst, dst_stride, dst_off, n):
        self._ck(self.lib.fedmx_exp_diff(self.h, _ptr(x), n_eq, comp_a, comp_b, _ptr(dst), dst_stride, dst_off, n))

    def launch_count(self, reset=False):
        return int(self.lib.fedmx_launch_count(self.h, 1 if reset else 0))

    def profile_enable(self, on=True):
        """on: False/0 off, True/1 kernel classes, 2 = plus the nested multigrid-level / collective break-down."""
        self._ck(self.lib.fedmx_profile_enable(self.h, int(on)))

    def profile_get(self):
        """{kind: (launches, total_ms)} since the last call, per kernel class (CUDA events on the stream)."""
        out = {}
        for kind, name in enumerate(("assemble_jacobian", "assemble_residual", "spmv", "orthogonalisation", "preconditioner",
                                     "mg_fine", "mg_mid", "mg_coarse", "comm")):
            n, ms = C.c_int64(), C.c_double()
            self._ck(self.lib.fedmx_profile_get(self.h, kind, C.byref(n), C.byref(ms)))
            out[name] = (n.value, ms.value)
        return out

    # ---- multi-GPU ------------------------------------------------------------------------------------
    def comm_init(self, n_ranks, rank, unique_id):
        self._ck(self.lib.fedmx_comm_init(self.h, n_ranks, rank, unique_id))

    def set_halo(self, neighbor_ranks, send_ptr, send_nodes, recv_ptr, recv_nodes):
        a = [np.ascontiguousarray(v, dtype=np.int32) for v in (neighbor_ranks, send_ptr, send_nodes, recv_ptr, recv_nodes)]
        self._ck(self.lib.fedmx_halo_set(self.h, a[0].size, *[_np_ptr(v) for v in a]))

    def halo_exchange(self, vec, n_comp=None):
        self._ck(self.lib.fedmx_halo_exchange(self.h, _ptr(vec), self.n_eq if n_comp is None else n_comp))

    def allreduce_sum(self, buf, n=None):
        self._ck(self.lib.fedmx_allreduce_sum(self.h, _ptr(buf), buf.numel() if n is None else n))

    def peer_init(self, world, rank, neighbor_ranks, recv_ptr, n_global_vertices):
        """NVLink peer-memory transport (fedmx_peer.cu): allocate the symmetric heap, exchange the CUDA IPC handles and the
        per-neighbour staging offsets over torch.distributed, map every peer.  Call after comm_init + set_halo.
        Every stage that can fail locally (allocation, IPC open) is followed by a collective agreement, so either ALL ranks end
        up on the peer transport or all stay on NCCL -- no rank is left waiting in a collective the others skipped.
        Returns True when the peer transport is active."""
        import torch.distributed as dist

        def agree(ok):
            t = torch.tensor([1 if ok else 0], device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return bool(t.item())

        def stage(fn):
            try:
                fn()
                return True
            except Exception as exc:                           # e.g. CUDA IPC not permitted in this container
                print(f"[fedm_b200] rank {rank}: peer-memory transport unavailable ({exc})", flush=True)
                return False

        nb = [int(q) for q in neighbor_ranks]
        rp = [int(v) for v in recv_ptr]
        plans = [None] * world
        dist.all_gather_object(plans, (nb, rp))
        halo_cap = 5 * max(p[1][-1] for p in plans) + 64          # FEDMX_MAX_EQ components per ghost node
        big_cap = int(n_global_vertices) // 2 + 1024              # restricted multigrid residual (level 1 of the hierarchy)
        handle = C.create_string_buffer(64)
        created = stage(lambda: self._ck(self.lib.fedmx_peer_create(self.h, halo_cap, big_cap, handle)))
        if not agree(created):
            if created:
                self.peer_enable(False)
            return False
        handles = [None] * world
        dist.all_gather_object(handles, handle.raw)
        opened = stage(lambda: self._ck(self.lib.fedmx_peer_open(self.h, b"".join(handles))))
        if not agree(opened):
            self.peer_enable(False)
            return False
        from .partition import peer_staging_offsets
        remote = np.array(peer_staging_offsets(rank, plans), dtype=np.int32)
        ready = stage(lambda: self._ck(self.lib.fedmx_peer_halo_set(self.h, len(nb), _np_ptr(remote))))
        if not agree(ready):
            self.peer_enable(False)
            return False
        dist.barrier()
        return True

    def peer_enable(self, on=True):
        self._ck(self.lib.fedmx_peer_enable(self.h, 1 if on else 0))

    def peer_error(self):
        return int(self.lib.fedmx_peer_error(self.h))


def pi34_step(dt, err, tol, dt_min, dt_max):
    e = (C.c_double * 3)(*[float(v) for v in list(err)[:3]] + [1.0] * max(0, 3 - len(err)))
    return float(_l.load().fedmx_pi34_step(dt, e, tol, dt_min, dt_max))


def h211b_step(dt, dt_old, err, tol, dt_min, dt_max):
    e = (C.c_double * 3)(*[float(v) for v in list(err)[:3]] + [1.0] * max(0, 3 - len(err)))
    return float(_l.load().fedmx_h211b_step(dt, dt_old, e, tol, dt_min, dt_max))


def pid_step(dt, err, tol, dt_min, dt_max):
    """``adaptive_timestep`` (fedm/functions.py:915-927) through the library's host libm implementation."""
    e = (C.c_double * 3)(*[float(v) for v in err])
    if e[0] == 0.0 or e[2] == 0.0:
        raise ZeroDivisionError("float division by zero")          # what the reference's Python expression raises
    return float(_l.load().fedmx_pid_step(dt, e, tol, dt_min, dt_max))
