"""Extracts the reference's golden results into small fixtures under tests/golden/.

Run once in the build container (where /root/reference exists):  python tools/make_golden.py
The GPU box has no /root/reference, so tests only ever read tests/golden/.  Sources:
  tests/integrated_tests/time_of_flight/20220707_results/{electrons000000.vtu, relative error.log}
  tests/integrated_tests/glow_discharge/20220707_results/{electrons,Ar_plus,Ar_star,Ar_1p0}.h5, relative error.log
  tests/integrated_tests/streamer_discharge/20220707_results/relative error.log
h5py/vtk are not installed: the VTU is ASCII (regex), the HDF5 files are superblock-v0 files with
contiguous little-endian data sets whose absolute byte offsets were located once with a throw-away
B-tree walker and are sanity-checked below.
"""
import re
import shutil
from pathlib import Path

import numpy as np

REF = Path("/root/reference/tests/integrated_tests")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"

OFF0 = dict(topology=4496, geometry=486544, cell_dofs=809760, vector=1289760)
OFF1 = dict(topology=2093712, geometry=2575760, cell_dofs=2898976, vector=3378976)
NV, NC = 20201, 40000


def rd(buf, off, dtype, count):
    return np.frombuffer(buf, dtype=dtype, count=count, offset=off).copy()


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    # --- error logs (plain text, byte copies)
    for case in ("time_of_flight", "glow_discharge", "streamer_discharge"):
        shutil.copyfile(REF / case / "20220707_results" / "relative error.log", OUT / f"{case}_relative_error.log")
    # --- ToF field
    s = (REF / "time_of_flight/20220707_results/electrons000000.vtu").read_text()
    m = re.search(r'<PointData.*?<DataArray[^>]*Name="f_3199"[^>]*>(.*?)</DataArray>', s, re.S)
    f = np.array(m.group(1).split(), dtype=np.float64)
    assert f.size == 1681
    np.save(OUT / "tof_electrons_f3199.npy", f)
    # --- glow HDF5
    out = {}
    for name in ("electrons", "Ar_plus", "Ar_star", "Ar_1p0"):
        buf = (REF / f"glow_discharge/20220707_results/{name}.h5").read_bytes()
        v0 = rd(buf, OFF0["vector"], "<f8", NV)
        ic = np.log(3.21877e22) if name == "Ar_1p0" else np.log(1e12)
        assert np.all(v0 == v0[0]) and abs(v0[0] - ic) < 1e-12, name
        out[f"{name}_0"] = v0
        if name != "Ar_1p0":
            out[f"{name}_1"] = rd(buf, OFF1["vector"], "<f8", NV)
        if name == "electrons":
            out["topology"] = rd(buf, OFF0["topology"], "<i4", NC * 3).reshape(NC, 3)
            out["geometry"] = rd(buf, OFF0["geometry"], "<f8", NV * 2).reshape(NV, 2)
            out["cell_dofs"] = rd(buf, OFF0["cell_dofs"], "<i4", NC * 3)
            assert out["topology"].min() == 0 and out["topology"].max() == NV - 1
            assert sorted(set(out["cell_dofs"].tolist())) == list(range(NV))
    np.savez_compressed(OUT / "glow_h5.npz", **out)
    # --- input decks of the glow and streamer cases (plain data files read by the reference's fedm/file_io.py)
    for case in ("glow_discharge", "streamer_discharge"):
        dst = OUT / f"{case}_input"
        if dst.exists():
            shutil.rmtree(dst)
        shutil.copytree(REF / case / "file_input", dst)
    for k, v in out.items():
        print(k, v.shape, v.dtype, float(np.min(v)), float(np.max(v)))


if __name__ == "__main__":
    main()
