"""ctypes binding of libdppb200.so (include/dpp_b200.h).  No torch types cross this boundary.

The library is the product: if it is missing or cannot be loaded this module raises -- there is
no CPU fallback (oracle/ is test infrastructure and is never imported from here).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdppb200.so")

# enums (keep in sync with include/dpp_b200.h)
KSP_CG, KSP_GMRES, KSP_PICARD = 0, 1, 2
PC_NONE, PC_JACOBI, PC_PBJACOBI, PC_FIELDSPLIT = 0, 1, 2, 3
FS_ADDITIVE, FS_MULTIPLICATIVE = 0, 1
INNER_PREONLY, INNER_CG = 0, 1
OP_MATRIX_FREE, OP_ASSEMBLED = 0, 1
KERNEL_GENERAL, KERNEL_STRUCTURED = 0, 1
ERR_NO_DEVICE = -5


class DppOptions(C.Structure):
    _fields_ = [
        ("ksp_type", C.c_int32), ("pc_type", C.c_int32), ("fieldsplit_type", C.c_int32),
        ("inner_ksp_type", C.c_int32), ("inner_pc_type", C.c_int32), ("operator_mode", C.c_int32),
        ("max_it", C.c_int32), ("gmres_restart", C.c_int32), ("inner_max_it", C.c_int32),
        ("check_every", C.c_int32),
        ("rtol", C.c_double), ("atol", C.c_double), ("dtol", C.c_double),
        ("inner_rtol", C.c_double), ("inner_atol", C.c_double),
    ]


class DppResult(C.Structure):
    _fields_ = [
        ("iterations", C.c_int32), ("converged_reason", C.c_int32), ("inner_iterations", C.c_int32),
        ("history_len", C.c_int32),
        ("residual_norm", C.c_double), ("rhs_norm", C.c_double), ("solve_ms", C.c_double),
        ("setup_ms", C.c_double), ("apply_ms", C.c_double), ("apply_count", C.c_int64),
    ]


class DppInfo(C.Structure):
    _fields_ = [
        ("kernel_family", C.c_int32), ("dim", C.c_int32), ("degree", C.c_int32),
        ("grid_nodes", C.c_int32 * 3),
        ("n_nodes", C.c_int64), ("n_cells", C.c_int64), ("n_owned_nodes", C.c_int64),
        ("rank", C.c_int32), ("world", C.c_int32), ("sm_count", C.c_int32), ("peer_memory", C.c_int32),
        ("device_bytes", C.c_int64),
    ]


# every symbol include/dpp_b200.h declares (tests/test_cabi_symbols.py checks the header against this)
_PROTOTYPES = {
    "dpp_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int,
                             C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "dpp_destroy": (None, [C.c_void_p]),
    "dpp_last_error": (C.c_char_p, [C.c_void_p]),
    "dpp_get_info": (C.c_int, [C.c_void_p, C.POINTER(DppInfo)]),
    "dpp_force_kernel_family": (C.c_int, [C.c_void_p, C.c_int]),
    "dpp_set_numbering": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dpp_set_params": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double]),
    "dpp_set_dirichlet": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "dpp_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64]),
    "dpp_comm_add_neighbor": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "dpp_comm_ipc_blob_size": (C.c_int, []),
    "dpp_comm_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dpp_comm_ipc_import": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dpp_comm_ipc_disable": (C.c_int, [C.c_void_p]),
    "dpp_fused_cg_supported": (C.c_int, [C.c_void_p]),
    "dpp_set_fused_cg": (C.c_int, [C.c_void_p, C.c_int]),
    "dpp_apply_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "dpp_apply_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "dpp_get_diagonal_host": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dpp_assemble_csr": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "dpp_get_csr_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dpp_get_csr_block_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dpp_time_assembly": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                   C.POINTER(C.c_int64)]),
    "dpp_default_options": (None, [C.POINTER(DppOptions)]),
    "dpp_solve": (C.c_int, [C.c_void_p, C.POINTER(DppOptions), C.c_void_p, C.POINTER(DppResult), C.c_void_p,
                            C.c_int32]),
    "dpp_solution_dev": (C.c_void_p, [C.c_void_p]),
    "dpp_time_apply": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "dpp_time_cg_kernels": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                     C.POINTER(C.c_double)]),
    "dpp_time_cg_block_kernels": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                                          C.POINTER(C.c_double)]),
    "dpp_kernel_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "dpp_plan_x_segments": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "dpp_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "dpp_error_norms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "dpp_darcy_velocity": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int32, C.c_void_p,
                                    C.c_void_p]),
    "dpp_lanczos": (C.c_int, [C.c_void_p, C.c_int, C.c_int32, C.c_uint64, C.c_void_p, C.c_void_p,
                             C.POINTER(C.c_int32)]),
    "dpp_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "dpp_host_free": (C.c_int, [C.c_void_p]),
}

_lib = None


class DppLibraryError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libdppb200.so (built in-tree by __graft_entry__.build() / perphil_b200/csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DppLibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C perphil_b200/csrc). perphil_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library drift apart
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
