"""north_star: "the sparsity pattern and dof maps are bit-exact to FEniCS 2019.1.0".  Through the C ABI on the device context:
fedmx_dofmap_fenics on the glow-discharge mesh == the dof map stored in the golden HDF5 file, and fedmx_pattern_get_fenics ==
dolfin's pattern (every cell-dof x cell-dof pair) expressed in that numbering, for a mixed space (dof = n_eq*d + comp)."""
import numpy as np
import pytest
import scipy.sparse as sp

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from fedm_b200 import hotpath as hp


def test_fenics_dofmap_and_pattern_through_the_abi(golden_dir):
    g = np.load(golden_dir / "glow_h5.npz")
    topo, geom = g["topology"], g["geometry"]
    n, neq = geom.shape[0], 2
    hot = hp.HotPath(0)
    hot.set_mesh(neq, geom, topo)
    d2v_gold = np.empty(n, dtype=np.int64)
    d2v_gold[g["cell_dofs"]] = topo.ravel()
    d2v = hot.dofmap_fenics()
    assert np.array_equal(d2v, d2v_gold)
    v2d = np.empty(n, dtype=np.int64)
    v2d[d2v] = np.arange(n)
    # dolfin's pattern: all pairs of dofs sharing a cell, in dolfin's numbering
    cd = v2d[topo]                                                            # (cells, 3) dofs of the scalar space
    dofs = (neq * cd[:, :, None] + np.arange(neq)[None, None, :]).reshape(-1, 3 * neq)
    rows = np.repeat(dofs, 3 * neq, axis=1).ravel()
    cols = np.tile(dofs, (1, 3 * neq)).ravel()
    ref = sp.csr_matrix((np.ones(rows.size), (rows, cols)), shape=(neq * n, neq * n))
    ref.sum_duplicates()
    ref.sort_indices()
    indptr, indices = hot.pattern_fenics()
    assert np.array_equal(indptr, ref.indptr) and np.array_equal(indices, ref.indices)
    hot.close()
