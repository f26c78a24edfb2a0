"""Pins the glow-discharge oracle (LMEA, 5 coupled fields, 7 reactions, look-up tables, semi-implicit coefficients,
secondary-emission boundary fluxes, time-dependent Dirichlet value, degree-4 rules, adaptive BDF2) on the reference's
goldens: tests/integrated_tests/glow_discharge/20220707_results/{electrons,Ar_plus,Ar_star}.h5 (`vector` at t = 1e-11,
in dolfin dof order) and the 6-row 'relative error.log'.  The reference test asserts 1e-5 / allclose
(test_glow_discharge.py:48-62); the restatement reproduces FEniCS to ~1e-12."""
import numpy as np

from oracle.glow import GlowProblem


def test_glow_matches_golden_log_and_fields(golden_dir):
    p = GlowProblem(golden_dir / "glow_discharge_input")
    # reaction matrices by substring counting (fedm/file_io.py:281-296)
    assert p.inp.power.tolist() == [[1, 0, 0, 1], [1, 0, 0, 1], [0, 1, 0, 1], [0, 1, 0, 1], [0, 2, 0, 0], [0, 1, 0, 0], [1, 0, 0, 1]]
    assert p.inp.energy_loss == [11.55, 15.76, -11.55, 4.21, -7.34, 0.0, 1.0]
    x, rows, its, out = p.run()
    gold = [[float(t) for t in ln.split()] for ln in (golden_dir / "glow_discharge_relative_error.log").read_text().splitlines()]
    assert len(rows) == len(gold) == 6
    assert [k for k, _ in its] == [2] * 6                       # exactly 2 Newton iterations per step
    for (e, dto, dt), (ge, gdto, gdt) in zip(rows, gold):
        assert abs(e - ge) <= 1e-10 * ge
        assert abs(dt - gdt) <= 1e-11 * gdt and abs(dto - gdto) <= 1e-11 * gdto
    g = np.load(golden_dir / "glow_h5.npz")
    d2v = np.empty(20201, dtype=np.int64)
    d2v[g["cell_dofs"]] = g["topology"].ravel()
    O = out.reshape(-1, 5)
    for name, k in (("electrons", 3), ("Ar_plus", 2), ("Ar_star", 1)):
        ref = g[name + "_1"]
        assert np.abs((O[d2v, k] - ref) / ref).max() <= 1e-11, name
