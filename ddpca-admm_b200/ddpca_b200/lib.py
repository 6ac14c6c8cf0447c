"""ctypes binding of include/ddpca_b200.h (the drop-in C ABI)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class DdpcaError(RuntimeError):
    pass


def library_path() -> str:
    # DDPCA_B200_LIB selects another build of the same library (kernel tuning variants)
    return os.environ.get("DDPCA_B200_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libddpca_b200.so")


# every symbol include/ddpca_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "ddpca_last_error", "ddpca_abi_version", "ddpca_device_count",
    "ddpca_plan_create", "ddpca_plan_create_blocks", "ddpca_plan_create_tri", "ddpca_plan_sizes", "ddpca_plan_get", "ddpca_plan_destroy",
    "ddpca_mg_create", "ddpca_mg_create_batch", "ddpca_mg_setup_dryrun", "ddpca_mg_batch_result", "ddpca_mg_destroy", "ddpca_mg_pcg", "ddpca_mg_pcg_dev",
    "ddpca_mg_vcycle", "ddpca_mg_spmv", "ddpca_mg_restrict", "ddpca_mg_prolong_add",
    "ddpca_mg_coarse_solve", "ddpca_mg_mult_solv", "ddpca_mg_bicgstab", "ddpca_mg_gmres",
    "ddpca_mg_level_info", "ddpca_mg_launch_count", "ddpca_mg_set_stream",
    "ddpca_mg_profile", "ddpca_mg_profile_get", "ddpca_mg_last_timing",
    "ddpca_ldlt_create", "ddpca_ldlt_create_dense", "ddpca_ldlt_solve", "ddpca_ldlt_solve_dev", "ddpca_ldlt_info", "ddpca_ldlt_destroy",
    "ddpca_admm_create", "ddpca_admm_set_body", "ddpca_admm_set_body_accuprol", "ddpca_admm_set_interface",
    "ddpca_admm_set_side_op", "ddpca_admm_set_side_solver", "ddpca_admm_set_macro", "ddpca_admm_set_macro_mg", "ddpca_admm_set_body_globtran_d1", "ddpca_admm_set_macro1", "ddpca_admm_finalize",
    "ddpca_admm_step", "ddpca_admm_row_length", "ddpca_admm_get_disp", "ddpca_admm_get_side", "ddpca_admm_get_gamma",
    "ddpca_admm_launch_count", "ddpca_admm_destroy", "ddpca_admm_set_partition", "ddpca_admm_exchange_sizes",
    "ddpca_admm_set_exchange", "ddpca_admm_exchange_peers", "ddpca_admm_set_stream", "ddpca_admm_phase", "ddpca_admm_monitor_row",
    "ddpca_admm_set_smoother", "ddpca_admm_reset", "ddpca_admm_set_consforc", "ddpca_admm_body_iters", "ddpca_admm_profile", "ddpca_admm_profile_get", "ddpca_gamma_project", "ddpca_admm_set_side_iterative", "ddpca_admm_set_macro1_mg",
    "ddpca_partition_bodies", "ddpca_admm_group_create", "ddpca_admm_group_size", "ddpca_admm_group_member", "ddpca_admm_group_owner",
    "ddpca_admm_group_device", "ddpca_admm_group_finalize", "ddpca_admm_group_step", "ddpca_admm_group_launch_count", "ddpca_admm_group_destroy",
]


def load_library() -> C.CDLL:
    """Load libddpca_b200.so; raises (never falls back) if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise DdpcaError(
            f"{path} is missing: build it with `make -C ddpca-admm_b200` "
            "(or __graft_entry__.build()); there is no CPU fallback"
        )
    lib = C.CDLL(path)
    lib.ddpca_last_error.restype = C.c_char_p
    lib.ddpca_mg_launch_count.restype = C.c_long
    lib.ddpca_admm_launch_count.restype = C.c_long
    lib.ddpca_admm_group_launch_count.restype = C.c_long
    lib.ddpca_admm_group_member.restype = C.c_void_p
    _LIB = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise DdpcaError(load_library().ddpca_last_error().decode())


def device_count() -> int:
    return int(load_library().ddpca_device_count())
