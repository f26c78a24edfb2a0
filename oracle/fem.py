"""P1 finite-element assembly as dolfin/FFC do it (oracle -- test infrastructure only).

Restates, for node-interleaved mixed P1 spaces (dof = N_eq*node + comp):
* ``assemble(F, tensor=b)`` / ``assemble(J, tensor=A)``  (reference call sites fedm/functions.py:192,200),
* ``DirichletBC.apply(b, x)`` / ``.apply(A)``            (fedm/functions.py:193-194, 201-202).

A weak form is given in *pointwise flux form*: for every equation k the integrand is
``s_k * v_k + f_k . grad(v_k)`` with s_k, f_k functions of the unknowns, their (cell-constant)
gradients, P1 coefficient fields and x.  The element vector is
``R[k,a] = sum_q w_q |detJ| (s_k lambda_a + f_k . g_a)``; the element matrix is obtained by
forward-mode AD with respect to the 3*N_eq local dofs (oracle/dual.py), i.e. it does not share
any hand-derived Jacobian with the CUDA kernels.
"""
import numpy as np
import scipy.sparse as sp

from .dual import Dual, Nest
from .quadrature import triangle_rule, facet_rule


class Geometry:
    """Per-cell affine data: |detJ| and the constant gradients g_a of the P1 basis."""

    def __init__(self, mesh):
        x = mesh.coords[mesh.cells]                       # (nc, 3, 2)
        j00 = x[:, 1, 0] - x[:, 0, 0]
        j01 = x[:, 2, 0] - x[:, 0, 0]
        j10 = x[:, 1, 1] - x[:, 0, 1]
        j11 = x[:, 2, 1] - x[:, 0, 1]
        det = j00 * j11 - j01 * j10
        self.detJ = np.abs(det)
        g = np.empty((x.shape[0], 3, 2))
        g[:, 1, 0] = j11 / det
        g[:, 1, 1] = -j01 / det
        g[:, 2, 0] = -j10 / det
        g[:, 2, 1] = j00 / det
        g[:, 0] = -g[:, 1] - g[:, 2]
        self.g = g
        self.x = x


class CellContext:
    """What a pointwise integrand may query at one quadrature point of every cell."""

    def __init__(self, asm, lam, Uloc, coefs, gradU):
        self.asm, self.lam, self.Uloc, self.coefs, self._gradU = asm, lam, Uloc, coefs, gradU
        self.geo = asm.geo
        xq = np.einsum("a,cad->cd", lam, self.geo.x)
        self.x = (xq[:, 0], xq[:, 1])

    def u(self, k):
        l = self.lam
        return self.Uloc[k][0] * l[0] + self.Uloc[k][1] * l[1] + self.Uloc[k][2] * l[2]

    def grad_u(self, k):
        return self._gradU[k]

    def coef(self, name):
        c = self.coefs[name][self.asm.mesh.cells]          # (nc,3)
        return c @ self.lam

    def coef_nodal(self, name):
        return self.coefs[name][self.asm.mesh.cells]

    def p2(self, nodal6):
        """Value of a P2 interpolant given its 6 nodal values per cell (3 vertices, then the
        midpoints of the edges opposite to vertex 0,1,2) -- FFC's treatment of Expression(degree=2)."""
        l0, l1, l2 = self.lam
        phi = np.array([l0 * (2 * l0 - 1), l1 * (2 * l1 - 1), l2 * (2 * l2 - 1),
                        4 * l1 * l2, 4 * l0 * l2, 4 * l0 * l1])
        return nodal6 @ phi

    def spatial_grad(self, fn, unknowns=(), coef_names=()):
        """Gradient of a composite of P1 fields: fn receives dicts of fields lifted to Nest duals in
        the barycentric coordinates and returns a Nest; grad = sum_a d/dlambda_a * g_a."""
        l = self.lam
        U, C = {}, {}
        for k in unknowns:
            a0, a1, a2 = self.Uloc[k]
            U[k] = Nest(a0 * l[0] + a1 * l[1] + a2 * l[2], [a0, a1, a2])
        for n in coef_names:
            c = self.coefs[n][self.asm.mesh.cells]
            C[n] = Nest(c @ l, [c[:, 0], c[:, 1], c[:, 2]])
        out = fn(U, C)
        g = self.geo.g
        gx = out.d[0] * g[:, 0, 0] + out.d[1] * g[:, 1, 0] + out.d[2] * g[:, 2, 0]
        gy = out.d[0] * g[:, 0, 1] + out.d[1] * g[:, 1, 1] + out.d[2] * g[:, 2, 1]
        return out.v, (gx, gy)


class Assembler:
    """Assembles residual and Jacobian of a pointwise-flux-form P1 problem on one mesh."""

    def __init__(self, mesh, n_eq):
        self.mesh, self.n_eq = mesh, n_eq
        self.geo = Geometry(mesh)
        self.ndof = n_eq * mesh.num_vertices
        c = mesh.cells.astype(np.int64)
        # local dof (k,a) -> index k*3+a ; global dof = n_eq*vertex + k
        self.ldofs = (n_eq * c[:, None, :] + np.arange(n_eq)[None, :, None]).reshape(c.shape[0], -1)
        fv, fc, fo = mesh.boundary_facets()
        self.facet_vertices, self.facet_cell, self.facet_opp = fv, fc, fo
        self._pattern = None

    # ---- pattern: every (cell dof, cell dof) pair, as dolfin's SparsityPatternBuilder inserts them
    def pattern(self):
        if self._pattern is None:
            ld = self.ldofs
            rows = np.repeat(ld, ld.shape[1], axis=1).ravel()
            cols = np.tile(ld, (1, ld.shape[1])).ravel()
            A = sp.csr_matrix((np.ones(rows.size, dtype=np.int8), (rows, cols)), shape=(self.ndof, self.ndof))
            A.sum_duplicates()
            A.sort_indices()
            self._pattern = (A.indptr.astype(np.int32), A.indices.astype(np.int32))
        return self._pattern

    def _seed(self, U, jac, cells_idx=None):
        n_eq = self.n_eq
        cells = self.mesh.cells if cells_idx is None else self.mesh.cells[cells_idx]
        nd = 3 * n_eq
        Uloc = []
        for k in range(n_eq):
            row = []
            for a in range(3):
                v = U[cells[:, a], k]
                row.append(Dual.seed(v, 3 * k + a, nd) if jac else v)
            Uloc.append(row)
        return Uloc

    def assemble_cells(self, U, point_fn, degree, coefs=None, jac=True):
        """Returns (Re (nc, 3*n_eq), Je (nc, 3*n_eq, 3*n_eq) or None) -- element tensors."""
        n_eq, geo = self.n_eq, self.geo
        lam_all, w_all = triangle_rule(degree)
        Uloc = self._seed(U, jac)
        g = geo.g
        gradU = []
        for k in range(n_eq):
            gx = Uloc[k][0] * g[:, 0, 0] + Uloc[k][1] * g[:, 1, 0] + Uloc[k][2] * g[:, 2, 0]
            gy = Uloc[k][0] * g[:, 0, 1] + Uloc[k][1] * g[:, 1, 1] + Uloc[k][2] * g[:, 2, 1]
            gradU.append((gx, gy))
        nc = self.mesh.num_cells
        nd = 3 * n_eq
        Re = np.zeros((nc, nd))
        Je = np.zeros((nc, nd, nd)) if jac else None
        for lam, w in zip(lam_all, w_all):
            ctx = CellContext(self, lam, Uloc, coefs or {}, gradU)
            terms = point_fn(ctx)                          # list over k of (s_k, (fx_k, fy_k)) or None
            W = w * geo.detJ
            for k, term in enumerate(terms):
                if term is None:
                    continue
                s, f = term
                for a in range(3):
                    t = 0.0
                    if s is not None:
                        t = s * lam[a]
                    if f is not None:
                        t = t + f[0] * g[:, a, 0] + f[1] * g[:, a, 1]
                    t = t * W
                    if isinstance(t, Dual):
                        Re[:, 3 * k + a] += t.v
                        if jac:
                            Je[:, 3 * k + a, :] += t.d
                    else:
                        Re[:, 3 * k + a] += t
        return Re, Je

    def assemble_facets(self, U, facet_fn, degree, markers, which, coefs=None, jac=True):
        """Exterior-facet integrals on facets whose marker is in ``which``.
        facet_fn(ctx, normal) -> list over k of s_k (scalar integrand multiplying v_k) or None.
        Returns (cell indices, Re, Je) for the owning cells."""
        sel = np.nonzero(np.isin(markers, list(which)))[0]
        if sel.size == 0:
            return sel, None, None
        n_eq, geo = self.n_eq, self.geo
        fcell, fopp, fv = self.facet_cell[sel], self.facet_opp[sel], self.facet_vertices[sel]
        cells = self.mesh.cells[fcell]
        xa, xb = self.mesh.coords[fv[:, 0]], self.mesh.coords[fv[:, 1]]
        t = xb - xa
        L = np.linalg.norm(t, axis=1)
        xo = self.mesh.coords[cells[np.arange(sel.size), fopp]]
        nrm = np.stack([t[:, 1], -t[:, 0]], axis=1) / L[:, None]
        flip = np.einsum("fd,fd->f", nrm, xo - xa) > 0.0
        nrm[flip] *= -1.0
        # local indices of the two facet vertices inside the owning cell
        la = np.argmax(cells == fv[:, [0]], axis=1)
        lb = np.argmax(cells == fv[:, [1]], axis=1)
        sub = _SubAssembler(self, fcell)
        Uloc = self._seed(U, jac, fcell)
        g = sub.geo.g
        gradU = []
        for k in range(n_eq):
            gx = Uloc[k][0] * g[:, 0, 0] + Uloc[k][1] * g[:, 1, 0] + Uloc[k][2] * g[:, 2, 0]
            gy = Uloc[k][0] * g[:, 0, 1] + Uloc[k][1] * g[:, 1, 1] + Uloc[k][2] * g[:, 2, 1]
            gradU.append((gx, gy))
        tq, wq = facet_rule(degree)
        nf, nd = sel.size, 3 * n_eq
        Re = np.zeros((nf, nd))
        Je = np.zeros((nf, nd, nd)) if jac else None
        ar = np.arange(nf)
        for tt, w in zip(tq, wq):
            lam = np.zeros((nf, 3))
            lam[ar, la] = 1.0 - tt
            lam[ar, lb] = tt
            ctx = FacetContext(sub, lam, Uloc, coefs or {}, gradU)
            terms = facet_fn(ctx, (nrm[:, 0], nrm[:, 1]))
            W = w * L
            for k, s in enumerate(terms):
                if s is None:
                    continue
                for a in range(3):
                    tv = s * lam[:, a] * W
                    if isinstance(tv, Dual):
                        Re[:, 3 * k + a] += tv.v
                        if jac:
                            Je[:, 3 * k + a, :] += tv.d
                    else:
                        Re[:, 3 * k + a] += tv
        return fcell, Re, Je

    # ---- global tensors
    def scatter_vector(self, Re, cell_idx=None, out=None):
        ld = self.ldofs if cell_idx is None else self.ldofs[cell_idx]
        b = np.zeros(self.ndof) if out is None else out
        np.add.at(b, ld.ravel(), Re.ravel())
        return b

    def scatter_matrix(self, blocks):
        """blocks: list of (Je, cell_idx or None).  Returns CSR with the full dolfin pattern
        (explicit zeros kept: the pattern must not be pruned by value)."""
        indptr, indices = self.pattern()
        rows, cols, vals = [], [], []
        for Je, cell_idx in blocks:
            ld = self.ldofs if cell_idx is None else self.ldofs[cell_idx]
            nd = ld.shape[1]
            rows.append(np.repeat(ld, nd, axis=1).ravel())
            cols.append(np.tile(ld, (1, nd)).ravel())
            vals.append(Je.ravel())
        # add the pattern itself with zeros so that structurally-present entries survive
        prow = np.repeat(np.arange(self.ndof), np.diff(indptr))
        rows.append(prow)
        cols.append(indices.astype(np.int64))
        vals.append(np.zeros(indices.size))
        A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(self.ndof, self.ndof)).tocsr()
        A.sort_indices()
        assert A.nnz == indices.size
        return A


class _SubAssembler:
    """View of an assembler restricted to a list of cells (used for facet integrals)."""

    def __init__(self, asm, cell_idx):
        self.mesh = _SubMesh(asm.mesh, cell_idx)
        self.geo = _SubGeo(asm.geo, cell_idx)
        self.n_eq = asm.n_eq


class _SubMesh:
    def __init__(self, mesh, idx):
        self.cells = mesh.cells[idx]
        self.coords = mesh.coords


class _SubGeo:
    def __init__(self, geo, idx):
        self.g, self.x, self.detJ = geo.g[idx], geo.x[idx], geo.detJ[idx]


class FacetContext(CellContext):
    """Same queries as CellContext but with per-facet barycentric coordinates lam (nf,3)."""

    def __init__(self, asm, lam, Uloc, coefs, gradU):
        self.asm, self.lam, self.Uloc, self.coefs, self._gradU = asm, lam, Uloc, coefs, gradU
        self.geo = asm.geo
        xq = np.einsum("ca,cad->cd", lam, self.geo.x)
        self.x = (xq[:, 0], xq[:, 1])

    def u(self, k):
        l = self.lam
        return self.Uloc[k][0] * l[:, 0] + self.Uloc[k][1] * l[:, 1] + self.Uloc[k][2] * l[:, 2]

    def coef(self, name):
        c = self.coefs[name][self.asm.mesh.cells]
        return np.einsum("ca,ca->c", c, self.lam)


def apply_dirichlet_vector(b, x, bc_dofs, bc_vals):
    """``bc.apply(b, x)``: b[i] = x[i] - g_i on constrained dofs (fedm/functions.py:193-194)."""
    b[bc_dofs] = x[bc_dofs] - bc_vals
    return b


def apply_dirichlet_matrix(A, bc_dofs):
    """``bc.apply(A)``: zero the constrained rows, put 1 on the diagonal, keep the pattern."""
    A = A.tocsr()
    mask = np.zeros(A.shape[0], dtype=bool)
    mask[bc_dofs] = True
    rowid = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    kill = mask[rowid]
    A.data[kill] = 0.0
    diag = kill & (A.indices == rowid)
    A.data[diag] = 1.0
    return A


def mass_matrix_cartesian(mesh):
    """Cartesian P1 mass matrix (dolfin ``project``: fedm_gd.py:309,432; fedm_tof.py:164-165)."""
    geo = Geometry(mesh)
    c = mesh.cells.astype(np.int64)
    Me = np.array([[2.0, 1.0, 1.0], [1.0, 2.0, 1.0], [1.0, 1.0, 2.0]]) / 24.0
    vals = (geo.detJ[:, None, None] * Me[None]).ravel()
    rows = np.repeat(c, 3, axis=1).ravel()
    cols = np.tile(c, (1, 3)).ravel()
    n = mesh.num_vertices
    return sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
