"""Host-side triangle meshes for the hot path (numpy; set-up only).

Mirrors what the reference obtains from dolfin: ``RectangleMesh(Point, Point, nx, ny[, "crossed"])``
(fedm_tof.py:95, fedm_gd.py:163), exterior-facet extraction and ``Marking_boundaries`` for straight
'line' boundaries (fedm/functions.py:86-124).  Vertex/cell numbering and coordinates follow dolfin
2019.1.0 so that results can be compared dof by dof with reference output files.
"""
import numpy as np

DOLFIN_EPS = 3.0e-16


class TriMesh:
    def __init__(self, coords, cells):
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        if self.coords.ndim != 2 or self.coords.shape[1] != 2 or self.cells.ndim != 2 or self.cells.shape[1] != 3:
            raise ValueError("TriMesh: coords must be (N,2) and cells (M,3)")
        self._ext = None

    num_vertices = property(lambda self: self.coords.shape[0])
    num_cells = property(lambda self: self.cells.shape[0])

    def exterior_facets(self):
        """Edges that belong to exactly one cell, as (F,2) vertex pairs (ascending within a pair)."""
        if self._ext is None:
            c = self.cells.astype(np.int64)
            edges = np.vstack((c[:, (0, 1)], c[:, (1, 2)], c[:, (0, 2)]))
            edges.sort(axis=1)
            key = edges[:, 0] * self.num_vertices + edges[:, 1]
            uniq, counts = np.unique(key, return_counts=True)
            bkey = uniq[counts == 1]
            self._ext = np.stack((bkey // self.num_vertices, bkey % self.num_vertices), axis=1).astype(np.int32)
        return self._ext

    def mark_lines(self, boundaries):
        """Facet markers 1..len(boundaries) for ['line', z1, z2, r1, r2] entries (later entries overwrite)."""
        fv = self.exterior_facets()
        pa, pb = self.coords[fv[:, 0]], self.coords[fv[:, 1]]
        pm = 0.5 * (pa + pb)
        markers = np.zeros(fv.shape[0], dtype=np.int32)
        for number, entry in enumerate(boundaries, start=1):
            kind = entry[0]
            if kind != "line":
                raise ValueError(f"fedm_b200.mesh: boundary type '{kind}' is not supported on the device path")
            z_lo, z_hi = entry[1] - DOLFIN_EPS, entry[2] + DOLFIN_EPS
            r_lo, r_hi = entry[3] - DOLFIN_EPS, entry[4] + DOLFIN_EPS
            hit = np.ones(fv.shape[0], dtype=bool)
            for pts in (pa, pb, pm):
                hit &= (pts[:, 0] >= r_lo) & (pts[:, 0] <= r_hi) & (pts[:, 1] >= z_lo) & (pts[:, 1] <= z_hi)
            markers[hit] = number
        return markers

    def hmax(self):
        p = self.coords[self.cells]
        d = [np.hypot(*(p[:, i] - p[:, j]).T) for i, j in ((0, 1), (1, 2), (0, 2))]
        return float(np.max(np.maximum.reduce(d)))

    def hmin(self):
        p = self.coords[self.cells]
        d = [np.hypot(*(p[:, i] - p[:, j]).T) for i, j in ((0, 1), (1, 2), (0, 2))]
        return float(np.min(np.maximum.reduce(d)))


def rectangle(p0, p1, nx, ny, diagonal="right", map_x=None, map_y=None):
    """dolfin-numbered structured triangulation of [p0, p1]; map_x/map_y warp the unit parameter (graded meshes)."""
    (x0, y0), (x1, y1) = p0, p1
    tx = np.arange(nx + 1, dtype=np.float64)
    ty = np.arange(ny + 1, dtype=np.float64)
    xs = x0 + tx * (x1 - x0) / nx if map_x is None else x0 + (x1 - x0) * map_x(tx / nx)
    ys = y0 + ty * (y1 - y0) / ny if map_y is None else y0 + (y1 - y0) * map_y(ty / ny)
    coords = np.empty(((nx + 1) * (ny + 1), 2))
    coords[:, 0] = np.tile(xs, ny + 1)
    coords[:, 1] = np.repeat(ys, nx + 1)
    quad = (np.arange(ny)[:, None] * (nx + 1) + np.arange(nx)[None, :]).ravel()
    sw, se, nw, ne = quad, quad + 1, quad + nx + 1, quad + nx + 2
    if diagonal == "right":
        cells = np.stack((sw, se, ne, sw, nw, ne), axis=1).reshape(-1, 3)
    elif diagonal == "crossed":
        xm = x0 + (np.arange(nx) + 0.5) * (x1 - x0) / nx if map_x is None else 0.5 * (xs[1:] + xs[:-1])
        ym = y0 + (np.arange(ny) + 0.5) * (y1 - y0) / ny if map_y is None else 0.5 * (ys[1:] + ys[:-1])
        mid = np.empty((nx * ny, 2))
        mid[:, 0] = np.tile(xm, ny)
        mid[:, 1] = np.repeat(ym, nx)
        coords = np.vstack((coords, mid))
        ctr = (nx + 1) * (ny + 1) + np.arange(nx * ny)
        cells = np.stack((sw, se, ctr, sw, nw, ctr, se, ne, ctr, nw, ne, ctr), axis=1).reshape(-1, 3)
    else:
        raise ValueError(f"unknown diagonal '{diagonal}'")
    return TriMesh(coords, cells)
