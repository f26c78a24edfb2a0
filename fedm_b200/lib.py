"""ctypes binding of libfedmx.so (C ABI: include/fedmx.h).  Fails loudly when the library is missing."""
import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libfedmx.so"

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)


class NewtonParams(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double), ("stol", C.c_double), ("max_it", C.c_int),
                ("ksp_restart", C.c_int), ("ksp_rtol", C.c_double), ("ksp_atol", C.c_double),
                ("ksp_max_it", C.c_int)]


class NewtonResult(C.Structure):
    _fields_ = [("iterations", C.c_int), ("reason", C.c_int), ("linear_iterations", C.c_int),
                ("fnorm0", C.c_double), ("fnorm", C.c_double), ("fnorm_history", C.c_double * 64)]


# name -> (restype, argtypes); the list doubles as the export check in tests/test_abi.py
SIGNATURES = {
    "fedmx_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "fedmx_ctx_destroy": (C.c_int, [C.c_void_p]),
    "fedmx_last_error": (C.c_char_p, [C.c_void_p]),
    "fedmx_version": (C.c_char_p, []),
    "fedmx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fedmx_mesh_set": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                 C.c_void_p, C.c_void_p, C.c_int64]),
    "fedmx_pattern_get": (C.c_int, [C.c_void_p, c_i64p, c_i64p, C.c_void_p, C.c_void_p]),
    "fedmx_matrix_get_bsr": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fedmx_gps_numbering": (C.c_int, [C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "fedmx_dofmap_fenics": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fedmx_pattern_get_fenics": (C.c_int, [C.c_void_p, c_i64p, c_i64p, C.c_void_p, C.c_void_p]),
    "fedmx_dirichlet_set": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "fedmx_model_set": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "fedmx_plugin_load": (C.c_int, [C.c_void_p, C.c_char_p]),
    "fedmx_plugin_set_field": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "fedmx_fields_bind": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "fedmx_assemble_residual": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p,
                                          C.c_int, C.c_void_p]),
    "fedmx_assemble_jacobian": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p,
                                          C.c_int, C.c_void_p]),
    "fedmx_bdf2_history": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                     C.c_void_p, C.c_int, C.c_int]),
    "fedmx_spmv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fedmx_norm2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, c_dp]),
    "fedmx_dot": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, c_dp]),
    "fedmx_gmres": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int,
                              C.POINTER(C.c_int), c_dp, C.POINTER(C.c_int)]),
    "fedmx_newton_solve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p,
                                     C.c_int, C.POINTER(NewtonParams), C.POINTER(NewtonResult)]),
    "fedmx_error_norm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, c_dp]),
    "fedmx_error_norm_strided": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_double, c_dp]),
    "fedmx_pid_step": (C.c_double, [C.c_double, c_dp, C.c_double, C.c_double, C.c_double]),
    "fedmx_pi34_step": (C.c_double, [C.c_double, c_dp, C.c_double, C.c_double, C.c_double]),
    "fedmx_h211b_step": (C.c_double, [C.c_double, C.c_double, c_dp, C.c_double, C.c_double, C.c_double]),
    "fedmx_interp_table": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_double, C.c_void_p, C.c_int64]),
    "fedmx_strided_copy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                     C.c_int64, C.c_int64]),
    "fedmx_strided_scale": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                      C.c_int64, C.c_int64, C.c_double, C.c_double]),
    "fedmx_exp_diff": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64,
                                 C.c_int64]),
    "fedmx_comm_unique_id": (C.c_int, [C.c_char_p]),
    "fedmx_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p]),
    "fedmx_halo_set": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fedmx_halo_exchange": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "fedmx_allreduce_sum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "fedmx_ident_zeros_set": (C.c_int, [C.c_void_p, C.c_int]),
    "fedmx_pc_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fedmx_peer_create": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_char_p]),
    "fedmx_peer_open": (C.c_int, [C.c_void_p, C.c_char_p]),
    "fedmx_peer_halo_set": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "fedmx_peer_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "fedmx_peer_error": (C.c_int, [C.c_void_p]),
    "fedmx_precond_set": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "fedmx_p2line_assemble": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fedmx_p2line_solve": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fedmx_ilu0_factor": (C.c_int, [C.c_void_p]),
    "fedmx_ilu0_solve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fedmx_ilu0_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), c_i64p, C.c_void_p]),
    "fedmx_amg_reset": (C.c_int, [C.c_void_p]),
    "fedmx_amg_add_level": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fedmx_amg_set_coarse": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "fedmx_amg_set_fine_prolongation": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "fedmx_matrix_get_component": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "fedmx_launch_count": (C.c_int64, [C.c_void_p, C.c_int]),
    "fedmx_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "fedmx_profile_get": (C.c_int, [C.c_void_p, C.c_int, c_i64p, c_dp]),
}

_lib = None


def load():
    """Loads libfedmx.so once; raises RuntimeError (no fallback) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C fedm_b200/csrc`). fedm_b200 has no CPU fallback.")
    # torch bundles its own libnccl.so.2 (newer than the system one libfedmx links against); load torch first so
    # that the process ends up with ONE NCCL (torch's) and libfedmx binds to it.
    import torch  # noqa: F401
    lib = C.CDLL(str(LIB_PATH), mode=getattr(os, "RTLD_NOW", 2))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class FedmxError(RuntimeError):
    pass
