"""Meshes as dolfin 2019.1.0 builds them (oracle -- test infrastructure only).

``rectangle_mesh`` restates dolfin's ``RectangleMesh`` vertex/cell numbering for the "right" and
"crossed" diagonals; coordinates are computed as ``a + i*(b-a)/n`` (multiply first, then divide),
which reproduces the golden ``mesh/geometry`` of the glow-discharge HDF5 files bit-exactly
(tests/test_oracle_dofmap.py).  Used by fedm_tof.py:95, fedm_gd.py:163 of the
reference tests; the streamer mesh.xml is missing, so streamer runs use a synthetic "right" grid.
"""
import numpy as np


class Mesh:
    """Triangle mesh: coords (N_v,2) f8, cells (N_c,3) i4 (dolfin sorts cell vertices ascending)."""

    def __init__(self, coords, cells):
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        self._facets = None

    @property
    def num_vertices(self):
        return self.coords.shape[0]

    @property
    def num_cells(self):
        return self.cells.shape[0]

    def boundary_facets(self):
        """Exterior facets: (N_f,2) vertex ids, owning cell, and the cell-local opposite vertex."""
        if self._facets is None:
            c = self.cells
            e = np.concatenate([c[:, [1, 2]], c[:, [0, 2]], c[:, [0, 1]]], axis=0).astype(np.int64)
            opp = np.repeat(np.arange(3), c.shape[0])
            cell = np.tile(np.arange(c.shape[0]), 3)
            lo, hi = e.min(axis=1), e.max(axis=1)
            key = lo * self.num_vertices + hi
            order = np.argsort(key, kind="stable")
            ks = key[order]
            first = np.ones(len(ks), dtype=bool)
            first[1:] = ks[1:] != ks[:-1]
            last = np.ones(len(ks), dtype=bool)
            last[:-1] = ks[1:] != ks[:-1]
            single = order[first & last]
            single.sort()
            self._facets = (np.stack([lo[single], hi[single]], axis=1).astype(np.int32),
                            cell[single].astype(np.int32), opp[single].astype(np.int32))
        return self._facets

    def mark_boundaries(self, boundaries, eps=3.0e-16):
        """Restates ``Marking_boundaries`` for 'line' entries (fedm/functions.py:86-124):
        a facet gets marker idx+1 if *all* its vertices and its midpoint lie in the eps-inflated
        box [r1,r2]x[z1,z2] (dolfin's SubDomain.mark with check_midpoint=True); later entries win."""
        fv, _, _ = self.boundary_facets()
        markers = np.zeros(fv.shape[0], dtype=np.int32)
        x = self.coords
        for idx, b in enumerate(boundaries):
            if b[0] != "line":
                raise ValueError("only 'line' boundaries are restated in the oracle")
            z1, z2, r1, r2 = b[1] - eps, b[2] + eps, b[3] - eps, b[4] + eps

            def inside(p):
                return (p[:, 0] >= r1) & (p[:, 0] <= r2) & (p[:, 1] >= z1) & (p[:, 1] <= z2)

            pa, pb = x[fv[:, 0]], x[fv[:, 1]]
            ok = inside(pa) & inside(pb) & inside(0.5 * (pa + pb))
            markers[ok] = idx + 1
        return markers

    def hmax(self):
        p = self.coords[self.cells]
        a = np.linalg.norm(p[:, 1] - p[:, 2], axis=1)
        b = np.linalg.norm(p[:, 0] - p[:, 2], axis=1)
        c = np.linalg.norm(p[:, 0] - p[:, 1], axis=1)
        return float(np.max(np.maximum(np.maximum(a, b), c)))   # greatest vertex distance (dolfin Cell::h)


def rectangle_mesh(x0, y0, x1, y1, nx, ny, diagonal="right", grading=None):
    """dolfin ``RectangleMesh(Point(x0,y0), Point(x1,y1), nx, ny, diagonal)``.

    ``grading`` (optional pair of callables s->s') warps the uniform parameter in x and y for the
    synthetic refined streamer meshes; ``None`` keeps dolfin's exact coordinates.
    """
    ix = np.arange(nx + 1, dtype=np.float64)
    iy = np.arange(ny + 1, dtype=np.float64)
    if grading is None:
        xs = x0 + ix * (x1 - x0) / nx
        ys = y0 + iy * (y1 - y0) / ny
    else:
        xs = x0 + (x1 - x0) * grading[0](ix / nx)
        ys = y0 + (y1 - y0) * grading[1](iy / ny)
    X, Y = np.meshgrid(xs, ys, indexing="xy")          # row-major, x fastest
    coords = np.stack([X.ravel(), Y.ravel()], axis=1)
    jj, ii = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    v0 = (jj * (nx + 1) + ii).ravel()
    v1 = v0 + 1
    v2 = v0 + (nx + 1)
    v3 = v2 + 1
    if diagonal == "right":
        cells = np.empty((nx * ny, 2, 3), dtype=np.int64)
        cells[:, 0] = np.stack([v0, v1, v3], axis=1)
        cells[:, 1] = np.stack([v0, v2, v3], axis=1)
        cells = cells.reshape(-1, 3)
    elif diagonal == "crossed":
        if grading is None:
            xm = x0 + (np.arange(nx) + 0.5) * (x1 - x0) / nx
            ym = y0 + (np.arange(ny) + 0.5) * (y1 - y0) / ny
        else:
            xm = 0.5 * (xs[:-1] + xs[1:])
            ym = 0.5 * (ys[:-1] + ys[1:])
        XM, YM = np.meshgrid(xm, ym, indexing="xy")
        coords = np.concatenate([coords, np.stack([XM.ravel(), YM.ravel()], axis=1)], axis=0)
        c = (nx + 1) * (ny + 1) + (jj * nx + ii).ravel()
        cells = np.empty((nx * ny, 4, 3), dtype=np.int64)
        cells[:, 0] = np.stack([v0, v1, c], axis=1)
        cells[:, 1] = np.stack([v0, v2, c], axis=1)
        cells[:, 2] = np.stack([v1, v3, c], axis=1)
        cells[:, 3] = np.stack([v2, v3, c], axis=1)
        cells = cells.reshape(-1, 3)
    else:
        raise ValueError(diagonal)
    return Mesh(coords, cells)
