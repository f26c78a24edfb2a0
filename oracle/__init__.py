"""CPU oracle for the FEDM BDF2 hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain numpy/scipy restatement of what FEniCS 2019.1.0 + PETSc compute
for the reference's per-time-step path (``fedm/functions.py:174-202, 240-528, 915-1130`` and
the three integrated-test scripts under ``tests/integrated_tests``).  The arithmetic of that
path lives in un-vendored third-party packages (fenics-dolfin/ffc/ufl/fiat 2019.1.0, PETSc,
MUMPS, SCOTCH -- README.md:22 of the reference; versions of PETSc/MUMPS are not pinned by the
reference), so it cannot be imported or compiled here; the oracle restates the published
algorithms (P1 Lagrange triangles, FIAT quadrature schemes, Newton with SNES stopping rules,
variable-step BDF2 in log-density form) and is *pinned* against the reference's own goldens:

* PID controller: both golden ``relative error.log`` dt-columns, bit-exact
  (tests/test_oracle_controller.py),
* time-of-flight: golden VTU field + ``relative_error`` scalar (tests/test_oracle_tof.py),
* glow discharge: golden HDF5 ``vector`` data sets + 6-row error log (tests/test_oracle_glow.py),
* dolfin serial P1 dof numbering: golden ``cell_dofs`` (tests/test_oracle_dofmap.py).

The streamer case has no mesh and no golden densities in the checkout
(``.MISSING_LARGE_BLOBS``): its oracle is built from the same code path as the two pinned
cases and is "parity unpinned" for the assembled entries themselves; it is cross-checked by
executing the reference's UNMODIFIED streamer script (its own form-building code) on a generated
mesh through the shim's numpy backend (tests/test_shim_reference_scripts.py).  The 1-D time of
flight (BASELINE configs[0], ``tof1d.py``) has no fixture either: "parity unpinned", anchored on
its exact solution (tests/test_oracle_tof1d.py).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this package -- always as the checker / the reported baseline, never as the product path.
"""
