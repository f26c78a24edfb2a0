"""Pins oracle.mesh (dolfin 'crossed' RectangleMesh numbering, bit-exact coordinates) and
oracle.dofmap (SCOTCH-GPS serial P1 numbering) on the golden glow-discharge HDF5 data sets
(tests/integrated_tests/glow_discharge/20220707_results/electrons.h5: mesh/topology, mesh/geometry, cell_dofs)."""
import numpy as np

from oracle.dofmap import gps_numbering
from oracle.mesh import rectangle_mesh


def test_crossed_mesh_is_dolfin_bit_exact(golden_dir):
    g = np.load(golden_dir / "glow_h5.npz")
    m = rectangle_mesh(0.0, 0.0, 0.01, 0.01, 100, 100, "crossed")
    assert np.array_equal(m.cells, g["topology"])
    assert np.array_equal(m.coords, g["geometry"])          # bit-exact, not allclose


def test_gps_numbering_reproduces_golden_cell_dofs(golden_dir):
    g = np.load(golden_dir / "glow_h5.npz")
    topo, cell_dofs = g["topology"], g["cell_dofs"].reshape(-1, 3)
    n = 20201
    d2v_gold = np.empty(n, dtype=np.int64)
    d2v_gold[cell_dofs.ravel()] = topo.ravel()
    for start in (0, n - 1):
        d2v = gps_numbering(topo, n, start=start)
        assert np.array_equal(d2v, d2v_gold)
