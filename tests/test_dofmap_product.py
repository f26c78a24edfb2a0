"""The PRODUCT's FEniCS numbering (fedm_b200/dofmap.py -> native fedmx_gps_numbering, fedm_b200/csrc/fedmx_dofmap.cu) against the
golden glow-discharge HDF5 data written by dolfin 2019.1.0 (tests/integrated_tests/glow_discharge/20220707_results/electrons.h5,
extract in tests/golden/glow_h5.npz): `cell_dofs` bit for bit, and a checkpoint built from vertex values reproduces the golden
`vector` data set entry by entry.  Host-only code: runs without a GPU."""
import numpy as np
import pytest

from fedm_b200.dofmap import checkpoint_datasets, fenics_dof_to_vertex, fenics_vertex_to_dof, vertex_graph


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(golden_dir / "glow_h5.npz")


def test_dof_to_vertex_map_is_dolfins(gold):
    topo, cell_dofs = gold["topology"], gold["cell_dofs"].reshape(-1, 3)
    n = 20201
    d2v_gold = np.empty(n, dtype=np.int64)
    d2v_gold[cell_dofs.ravel()] = topo.ravel()
    for start in (0, n - 1):                                 # dolfin's numbering = start vertex 0; the opposite corner reaches the same root
        assert np.array_equal(fenics_dof_to_vertex(topo, n, start), d2v_gold)
    v2d = fenics_vertex_to_dof(topo, n)
    assert np.array_equal(v2d[topo].ravel(), gold["cell_dofs"])


def test_checkpoint_datasets_reproduce_golden_vector(gold):
    topo, geom = gold["topology"], gold["geometry"]
    n = geom.shape[0]
    d2v = np.empty(n, dtype=np.int64)
    d2v[gold["cell_dofs"]] = topo.ravel()
    vertex_values = np.empty(n)
    vertex_values[d2v] = gold["electrons_1"]                 # what a vertex-ordered solver would hold
    ds = checkpoint_datasets(geom, topo, vertex_values)
    assert np.array_equal(ds["vector"].ravel(), gold["electrons_1"])
    assert np.array_equal(ds["cell_dofs"].ravel(), gold["cell_dofs"]) and ds["cell_dofs"].dtype == np.int32
    assert np.array_equal(ds["mesh/topology"], topo) and np.array_equal(ds["mesh/geometry"], geom)
    assert ds["x_cell_dofs"].dtype == np.uint64 and ds["x_cell_dofs"][-1, 0] == 3 * topo.shape[0]


def test_vertex_graph_and_disconnected_mesh():
    cells = np.array([[0, 1, 2], [1, 2, 3]], dtype=np.int32)
    rowptr, col = vertex_graph(cells, 4)
    assert rowptr.tolist() == [0, 2, 5, 8, 10] and col.tolist() == [1, 2, 0, 2, 3, 0, 1, 3, 1, 2]
    from fedm_b200.lib import FedmxError
    with pytest.raises(FedmxError):
        fenics_dof_to_vertex(np.array([[0, 1, 2], [3, 4, 5]], dtype=np.int32), 6)       # two components: no serial numbering
