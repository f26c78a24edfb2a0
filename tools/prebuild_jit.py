"""Pre-compile the form-compiler plug-ins that the examples and the GPU tests use (CPU container: nvcc cross-compiles sm_100a).

The scripts are executed at tiny size on the shim's dry-run backend (fedm_b200/dolfin/dryrun_backend.py: no solve happens, the
forms are only collected); each form is lowered by fedm_b200.formcompiler and built into fedm_b200/jit_cache/, where the GPU box
finds it by hash (source + library headers) instead of running nvcc at test time."""
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "compat"), str(ROOT / "examples"), str(ROOT / "tests")]


def main():
    import dolfin
    from fedm_b200.dolfin.dryrun_backend import DryRunBackend

    backend = DryRunBackend()
    t_start = time.time()
    dolfin.set_backend(backend)
    golden = ROOT / "tests" / "golden"
    try:
        with tempfile.TemporaryDirectory() as tmp:
            tmp = Path(tmp)
            import glow_discharge
            import streamer_discharge
            import time_of_flight
            streamer_discharge.main(n=6, input_dir=golden / "streamer_discharge_input", output_dir=tmp / "s", linear_solver="mumps", max_steps=1)
            dolfin.parameters["form_compiler"]["quadrature_degree"] = None
            time_of_flight.main(output_dir=tmp / "t", nx=4, ny=4, T_final=2.5e-9 + 2e-12)
            glow_discharge.main(input_dir=golden / "glow_discharge_input", output_dir=tmp / "g", nx=4, max_steps=1)
            dolfin.parameters["form_compiler"]["quadrature_degree"] = None
            from fedm_b200 import functions as F
            from helpers_cases import run_helpers
            run_helpers(dolfin, F, n=4)
    finally:
        dolfin.set_backend(None)
        dolfin.parameters["form_compiler"]["quadrature_degree"] = None
    names = sorted(set(backend.built))
    # drop plug-ins of older library headers (their hash can no longer be requested)
    cache = ROOT / "fedm_b200" / "jit_cache"
    keep = {Path(n).stem for n in names}
    for f in cache.glob("form_*"):
        if f.stem not in keep and f.suffix in (".so", ".cu"):
            f.unlink()
    print(f"prebuilt {len(names)} form plug-ins in {time.time() - t_start:.1f} s:", ", ".join(names))


if __name__ == "__main__":
    main()
