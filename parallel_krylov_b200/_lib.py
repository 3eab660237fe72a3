"""ctypes binding of libpkrylov.so (C-ABI declared in include/pkrylov.h).

The product path is CUDA only: if the shared library is missing or fails to load this module raises — there is no
CPU fallback (the numpy restatement under oracle/ is test infrastructure and is never imported from here).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PK_LIB: developer switch to load another build of the library (A/B experiments on kernel variants)
LIB_PATH = os.environ.get("PK_LIB") or os.path.join(_HERE, "libpkrylov.so")

PK_KMAX = 32
PK_NCCL_ID_BYTES = 128
PK_IPC_HANDLE_BYTES = 64
PK_CG, PK_MRR, PK_KSKIPCG, PK_KSKIPMRR, PK_ADAPTIVEKSKIPMRR, PK_CGCG = range(6)
METHOD_IDS = {"cg": PK_CG, "mrr": PK_MRR, "kskipcg": PK_KSKIPCG, "kskipmrr": PK_KSKIPMRR,
              "adaptivekskipmrr": PK_ADAPTIVEKSKIPMRR, "cgcg": PK_CGCG}


class SolveOpts(C.Structure):
    _fields_ = [("maxiter", C.c_int64), ("tol", C.c_double), ("k", C.c_int32), ("check_every", C.c_int32),
                ("use_graph", C.c_int32), ("x_is_zero", C.c_int32), ("global_n", C.c_int64), ("d_mdiag", C.c_void_p),
                ("basis", C.c_int32), ("pad0", C.c_int32), ("lam_lo", C.c_double), ("lam_hi", C.c_double)]


class SolveResult(C.Structure):
    _fields_ = [("iterations", C.c_int64), ("entries", C.c_int64), ("converged", C.c_int32), ("final_k", C.c_int32),
                ("final_residual", C.c_double), ("elapsed_s", C.c_double), ("kernel_launches", C.c_int64),
                ("spmv_count", C.c_int64)]


_P = C.c_void_p
_I64 = C.c_int64
_SIGS = {
    "pk_version": (C.c_int, []),
    "pk_last_error": (C.c_char_p, []),
    "pk_ctx_create": (C.c_int, [C.POINTER(_P), C.c_int, _P]),
    "pk_ctx_destroy": (C.c_int, [_P]),
    "pk_ctx_sync": (C.c_int, [_P]),
    "pk_ctx_sm_count": (C.c_int, [_P]),
    "pk_prof_begin": (C.c_int, [_P, C.c_int]),
    "pk_prof_end": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "pk_mat_csr": (C.c_int, [_P, C.POINTER(_P), _I64, _I64, _I64, _P, _P, _P]),
    "pk_mat_csr64": (C.c_int, [_P, C.POINTER(_P), _I64, _I64, _I64, _P, _P, _P]),
    "pk_mat_dense": (C.c_int, [_P, C.POINTER(_P), _I64, _I64, _P, _I64]),
    "pk_mat_destroy": (C.c_int, [_P]),
    "pk_mat_kernel_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pk_mat_ld": (_I64, [_P]),
    "pk_mat_row_hashes": (C.c_int, [_P, _P]),
    "pk_mat_set_patterns": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P]),
    "pk_mat_set_halo": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _I64, _I64]),
    "pk_mat_halo_p2p_handle": (C.c_int, [_P, C.c_char_p]),
    "pk_mat_halo_p2p_open": (C.c_int, [_P, C.c_char_p, _P, _P]),
    "pk_mat_halo_p2p_disable": (C.c_int, [_P]),
    "pk_nccl_unique_id": (C.c_int, [C.c_char_p, C.c_char_p]),
    "pk_comm_init": (C.c_int, [_P, C.c_char_p, C.c_int, C.c_int, C.c_char_p]),
    "pk_comm_destroy": (C.c_int, [_P]),
    "pk_ctx_set_nocomm": (C.c_int, [_P, C.c_int]),
    "pk_p2p_handle": (C.c_int, [_P, C.c_char_p]),
    "pk_p2p_open": (C.c_int, [_P, C.c_int, C.c_int, C.c_char_p]),
    "pk_allreduce_sum": (C.c_int, [_P, _P, _I64]),
    "pk_allgather": (C.c_int, [_P, _P, _P, _I64]),
    "pk_spmv": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P]),
    "pk_matpow": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "pk_mat_matpow_info": (C.c_int, [_P, _P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pk_mat_set_band_ext": (C.c_int, [_P, _P, _I64, _I64]),
    "pk_mat_set_matpow_ext": (C.c_int, [_P, C.c_int, C.c_int, _I64, _I64, _P, _I64, _P, _P, _P, _I64, _P, _P, _P]),
    "pk_csr_band_info": (C.c_int, [_P, _I64, _I64, _P, _P, C.POINTER(C.c_int)]),
    "pk_dot": (C.c_int, [_P, _I64, _P, _P, _P]),
    "pk_gram": (C.c_int, [_P, C.c_int, _I64, _I64, _P, C.c_int, _P, C.c_int, _P]),
    "pk_work_doubles": (_I64, [C.c_int, _I64, C.c_int]),
    "pk_mat_diagonal": (C.c_int, [_P, _P]),
    "pk_mat_gershgorin": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "pk_solve": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, _P, _P, _I64, C.POINTER(SolveOpts),
                           C.POINTER(SolveResult)]),
    "pk_gen_stencil_counts": (C.c_int, [_P, _I64, _I64, _I64, _I64, _I64, _P]),
    "pk_gen_stencil_fill": (C.c_int, [_P, _I64, _I64, _I64, _I64, _I64, _P, _P, _P]),
    "pk_gen_banded_counts": (C.c_int, [_P, _I64, C.c_int, _I64, _I64, _P]),
    "pk_gen_banded_fill": (C.c_int, [_P, _I64, C.c_int, C.c_uint64, _I64, _I64, _P, _P, _P]),
    "pk_fill_hash_normal": (C.c_int, [_P, C.c_uint64, _I64, _I64, _P]),
}

_lib = None


class PkError(RuntimeError):
    pass


def exported_symbols():
    """Names include/pkrylov.h declares (used by the CPU test that checks the library exports all of them)."""
    return sorted(_SIGS)


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PkError(
            f"{LIB_PATH} not found: build it with `python -m parallel_krylov_b200.build` "
            "(nvcc, sm_100a).  parallel_krylov_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().pk_last_error()
        raise PkError(f"{what or 'libpkrylov'} failed ({status}): {msg.decode() if msg else '?'}")
