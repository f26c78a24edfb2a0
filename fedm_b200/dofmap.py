"""FEniCS 2019.1.0 dof numbering of P1 spaces as an export permutation (host side of fedm_b200/csrc/fedmx_dofmap.cu).

libfedmx works in mesh-vertex order; dolfin numbers the dofs of ``FunctionSpace(mesh, 'P', 1)`` with SCOTCH's
Gibbs-Poole-Stockmeyer ordering (fedm_gd.py:116-123), and everything the reference stores in dof order -- the
``vector`` / ``cell_dofs`` data sets of ``write_checkpoint`` (fedm/file_io.py:594-600) -- depends on it.
"""
import numpy as np

from . import lib as _l


def vertex_graph(cells, n_vertices):
    """CSR adjacency (ascending neighbours, no self entries) of the mesh's vertex graph."""
    c = np.asarray(cells, dtype=np.int64)
    e = np.concatenate([c[:, [0, 1]], c[:, [1, 2]], c[:, [2, 0]]])
    key = np.unique(np.concatenate([e[:, 0] * n_vertices + e[:, 1], e[:, 1] * n_vertices + e[:, 0]]))
    src = key // n_vertices
    rowptr = np.zeros(n_vertices + 1, dtype=np.int32)
    rowptr[1:] = np.cumsum(np.bincount(src, minlength=n_vertices))
    return rowptr, (key % n_vertices).astype(np.int32)


def fenics_dof_to_vertex(cells, n_vertices, start=0):
    """d2v[dof] = vertex for dolfin's serial P1 numbering of this mesh."""
    rowptr, col = vertex_graph(cells, n_vertices)
    d2v = np.empty(n_vertices, dtype=np.int32)
    rc = _l.load().fedmx_gps_numbering(n_vertices, rowptr.ctypes.data, col.ctypes.data, start, d2v.ctypes.data)
    if rc != 0:
        raise _l.FedmxError(f"fedmx_gps_numbering failed ({rc}): the mesh graph must be connected")
    return d2v


def fenics_vertex_to_dof(cells, n_vertices):
    d2v = fenics_dof_to_vertex(cells, n_vertices)
    v2d = np.empty_like(d2v)
    v2d[d2v] = np.arange(n_vertices, dtype=np.int32)
    return v2d


def checkpoint_datasets(coords, cells, vertex_values, v2d=None):
    """The data sets dolfin writes for one P1 checkpoint: mesh, per-cell dofs, and the vector in dolfin dof order."""
    nv, nc = coords.shape[0], cells.shape[0]
    v2d = fenics_vertex_to_dof(cells, nv) if v2d is None else v2d
    vec = np.empty(nv)
    vec[v2d] = np.asarray(vertex_values, dtype=np.float64)
    return {"mesh/topology": np.ascontiguousarray(cells, dtype=np.int32),
            "mesh/geometry": np.ascontiguousarray(coords, dtype=np.float64),
            "cell_dofs": v2d[np.asarray(cells)].astype(np.int32).reshape(-1, 1),
            "x_cell_dofs": (3 * np.arange(nc + 1, dtype=np.uint64)).reshape(-1, 1),
            "cells": np.arange(nc, dtype=np.uint64).reshape(-1, 1),
            "vector": vec.reshape(-1, 1)}
