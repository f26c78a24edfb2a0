"""HDF5 checkpoint container (fedm_b200/hdf5_lite.py): the strict reader is pinned on the reference's golden files written by
libhdf5 (tests/integrated_tests/glow_discharge/20220707_results/*.h5 -- read here when /root/reference exists, and through the
committed extract tests/golden/glow_h5.npz), the writer is checked through the reader and, message by message, against the
structure libhdf5 produced for the same data (reference call: fedm/file_io.py:594-600)."""
import struct
from pathlib import Path

import numpy as np
import pytest

from fedm_b200.hdf5_lite import H5CheckpointFile, H5Reader, read_h5

REF = Path("/root/reference/tests/integrated_tests/glow_discharge/20220707_results")


def checkpoint_arrays(g, vector):
    nc = g["topology"].shape[0]
    return {"mesh/topology": g["topology"].astype(np.int32), "mesh/geometry": g["geometry"],
            "cell_dofs": g["cell_dofs"].astype(np.int32).reshape(-1, 1),
            "x_cell_dofs": (3 * np.arange(nc + 1, dtype=np.uint64)).reshape(-1, 1),
            "cells": np.arange(nc, dtype=np.uint64).reshape(-1, 1), "vector": vector.reshape(-1, 1)}


@pytest.mark.skipif(not REF.exists(), reason="reference checkout not present")
def test_reader_on_libhdf5_files(golden_dir):
    g = np.load(golden_dir / "glow_h5.npz")
    for name in ("electrons", "Ar_plus", "Ar_star", "Ar_1p0"):
        t = H5Reader(REF / f"{name}.h5").tree()
        subs = list(t[name])
        assert subs == [f"{name}_0", f"{name}_1"][:len(subs)]
        for k, sub in enumerate(subs):
            cp = t[name][sub]
            assert sorted(cp) == ["cell_dofs", "cells", "mesh", "vector", "x_cell_dofs"]
            assert np.array_equal(cp["vector"].ravel(), g[f"{name}_{k}"])
            assert cp["vector"].shape == (20201, 1) and cp["cell_dofs"].dtype == np.int32
            assert np.array_equal(cp["x_cell_dofs"].ravel(), 3 * np.arange(40001)) and cp["cells"].dtype == np.uint64
        assert np.array_equal(t[name][f"{name}_0"]["mesh"]["topology"], g["topology"])
        assert np.array_equal(t[name][f"{name}_0"]["mesh"]["geometry"], g["geometry"])
        assert np.array_equal(t[name][f"{name}_0"]["cell_dofs"].ravel(), g["cell_dofs"])
    vecs = read_h5(REF / "electrons.h5", "electrons")
    assert len(vecs) == 2 and np.array_equal(vecs[1].ravel(), g["electrons_1"])


def structure(path):
    """Canonical listing of every object: path, message types/flags/bodies with addresses and time stamps masked."""
    rd = H5Reader(path)
    out = []

    def walk(addr, prefix):
        for nm, (child, is_group) in rd.group_members(addr).items():
            msgs = []
            for mtype, flags, body in rd.messages(child):
                if mtype == 0x11:
                    body = b"<tables>"
                elif mtype == 0x08:
                    body = body[:2] + b"<addr>" + body[10:]
                elif mtype == 0x12:
                    body = body[:4] + b"<time>"
                msgs.append((mtype, flags, bytes(body)))
            out.append((prefix + "/" + nm, msgs))
            if is_group:
                walk(child, prefix + "/" + nm)

    walk(rd.root_header, "")
    return out, (rd.leaf_k, rd.internal_k)


def test_writer_roundtrip_and_append(tmp_path, golden_dir):
    g = np.load(golden_dir / "glow_h5.npz")
    p = tmp_path / "electrons.h5"
    f = H5CheckpointFile(p)
    assert f.add_checkpoint("electrons", checkpoint_arrays(g, g["electrons_0"]), append=False) == 0
    size0 = p.stat().st_size
    f2 = H5CheckpointFile(p)                                  # a new writer object appends to the existing file
    assert f2.add_checkpoint("electrons", checkpoint_arrays(g, g["electrons_1"]), append=True) == 1
    assert p.stat().st_size < 2 * size0 + 4096                # data is never rewritten
    vecs = read_h5(p, "electrons")
    assert len(vecs) == 2 and np.array_equal(vecs[0].ravel(), g["electrons_0"]) and np.array_equal(vecs[1].ravel(), g["electrons_1"])
    t = H5Reader(p).tree()
    assert np.array_equal(t["electrons"]["electrons_1"]["mesh"]["topology"], g["topology"])
    assert t["electrons"]["electrons_1"]["mesh"]["geometry"].dtype == np.float64
    # append=False truncates
    f2.add_checkpoint("electrons", checkpoint_arrays(g, g["electrons_0"]), append=False)
    assert len(read_h5(p, "electrons")) == 1


def test_many_appends_split_symbol_tables(tmp_path):
    """> 8 members need several symbol-table nodes, > 256 a second B-tree level."""
    p = tmp_path / "many.h5"
    f = H5CheckpointFile(p)
    n = 300
    for k in range(n):
        f.add_checkpoint("u", {"vector": np.full((5, 1), float(k))}, append=k > 0)
    vecs = H5Reader(p).tree()["u"]
    assert len(vecs) == n
    for k in range(n):
        assert vecs[f"u_{k}"]["vector"][0, 0] == float(k)


@pytest.mark.skipif(not REF.exists(), reason="reference checkout not present")
def test_writer_matches_libhdf5_structure(tmp_path, golden_dir):
    g = np.load(golden_dir / "glow_h5.npz")
    p = tmp_path / "electrons.h5"
    f = H5CheckpointFile(p)
    f.add_checkpoint("electrons", checkpoint_arrays(g, g["electrons_0"]), append=False)
    f.add_checkpoint("electrons", checkpoint_arrays(g, g["electrons_1"]), append=True)
    mine, k_mine = structure(p)
    ref, k_ref = structure(REF / "electrons.h5")
    assert k_mine == k_ref
    assert [m[0] for m in mine] == [r[0] for r in ref]
    for (path, a), (_, b) in zip(mine, ref):
        assert a == b, path                                   # identical message bytes (dataspace, datatype, fill, attribute ...)
    # superblock fields other than addresses
    bm, br = p.read_bytes(), (REF / "electrons.h5").read_bytes()
    assert bm[:24] == br[:24] and struct.unpack_from("<Q", bm, 40)[0] == len(bm)
