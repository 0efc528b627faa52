"""ctypes binding of libadcme_b200.so — the same symbols a Julia `ccall` binds (julia/ADCMEB200.jl).

The library is mandatory: importing this module raises if the shared object is missing, and every
compute call raises `ADCMEError` when no sm_100 GPU is usable.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ADCME_B200_LIB", os.path.join(HERE, "libadcme_b200.so"))   # override: A/B builds

AB_OK, AB_EINVAL, AB_EHANDLE, AB_ECUDA, AB_ERANGE = 0, -1, -2, -3, -4
AB_ENOCONV, AB_EBREAKDOWN, AB_EUNSUPPORTED, AB_ENCCL = -5, -6, -7, -8

_i32, _i64, _dbl, _ptr, _str = C.c_int32, C.c_int64, C.c_double, C.c_void_p, C.c_char_p

# name -> (restype, argtypes); mirrors include/adcme_b200.h one to one
SIGNATURES = {
    "ab_init": (C.c_int, [C.c_int]),
    "ab_finalize": (C.c_int, []),
    "ab_device_count": (C.c_int, []),
    "ab_last_error": (_str, []),
    "ab_version": (_str, []),
    "ab_set_stream": (C.c_int, [_ptr]),
    "ab_synchronize": (C.c_int, []),
    "ab_set_option": (C.c_int, [_str, _dbl]),
    "ab_get_option": (C.c_int, [_str, _ptr]),
    "ab_asm_create": (C.c_int, [_i32, _i32, _dbl]),
    "ab_asm_add": (C.c_int, [_i32, _i32, _ptr, _ptr, _i32]),
    "ab_asm_add_coo": (C.c_int, [_i32, _ptr, _ptr, _ptr, _i64]),
    "ab_asm_nnz": (_i64, [_i32]),
    "ab_asm_copy": (C.c_int, [_i32, _ptr, _ptr, _ptr]),
    "ab_asm_destroy": (C.c_int, [_i32]),
    "ab_csr_from_coo": (C.c_int, [_i64, _ptr, _ptr, _ptr, _i64, _i64, C.c_int, _ptr]),
    "ab_csr_from_coo32": (C.c_int, [_i64, _ptr, _ptr, _ptr, _i64, _i64, C.c_int, _ptr]),
    "ab_coo_reorder": (C.c_int, [_i64, _ptr, _ptr, _ptr, _i64, _i64, C.c_int, _ptr, _ptr, _ptr, _ptr]),
    "ab_csr_create": (C.c_int, [_i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr]),
    "ab_csr_query": (C.c_int, [_i64, _ptr, _ptr, _ptr, _ptr]),
    "ab_csr_export": (C.c_int, [_i64, _ptr, _ptr, _ptr]),
    "ab_csr_export_map": (C.c_int, [_i64, _ptr]),
    "ab_csr_update_values": (C.c_int, [_i64, _ptr]),
    "ab_csr_transpose": (C.c_int, [_i64, _ptr]),
    "ab_csr_destroy": (C.c_int, [_i64]),
    "ab_compress_count": (C.c_int, [_i64, _ptr, _ptr]),
    "ab_compress": (C.c_int, [_i64, _ptr, _ptr, _ptr, _ptr]),
    "ab_compress_grad": (C.c_int, [_i64, _ptr, _ptr, _ptr]),
    "ab_coo_to_rows_count": (C.c_int, [_i64, _ptr, _ptr]),
    "ab_coo_to_rows": (C.c_int, [_i64, _ptr, _i64, _ptr, _ptr, _ptr]),
    "ab_spmm": (C.c_int, [_i64, C.c_int, _ptr, _i64, _ptr]),
    "ab_spmm_grad_values": (C.c_int, [_i64, _ptr, _ptr, _i64, _ptr]),
    "ab_spmm_grad_dense": (C.c_int, [_i64, _ptr, _i64, _ptr]),
    "ab_solve": (C.c_int, [_i64, _str, _ptr, _ptr, _dbl, C.c_int, _ptr, _ptr]),
    "ab_solve_grad": (C.c_int, [_i64, _str, _ptr, _ptr, _ptr, _ptr, _dbl, C.c_int, _ptr, _ptr]),
    "ab_cg_fixed": (C.c_int, [_i64, _ptr, _ptr, C.c_int, _ptr]),
    "ab_sparse_solver": (C.c_int, [_ptr, _ptr, _ptr, _i64, _ptr, _i64, _str, _ptr]),
    "ab_sparse_solver_grad": (C.c_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _ptr, _i64, _str, _ptr, _ptr]),
    "ab_factorize": (C.c_int, [_ptr, _ptr, _ptr, _i64, _i64, _i64, _ptr]),
    "ab_factorized_solve": (C.c_int, [_i64, _ptr, _i64, _ptr]),
    "ab_factorized_solve_grad": (C.c_int, [_i64, _ptr, _ptr, _ptr, _ptr, _i64, _i64, _ptr, _ptr]),
    "ab_trip_size": (_i64, [_i64]),
    "ab_trip_fetch": (C.c_int, [_i64, _ptr, _ptr, _ptr]),
    "ab_trip_fetch_idx2": (C.c_int, [_i64, _ptr, _ptr]),
    "ab_trip_destroy": (C.c_int, [_i64]),
    "ab_csr_from_trip": (C.c_int, [_i64, _i64, _i64, C.c_int, _ptr]),
    "ab_zero_out_row": (C.c_int, [_i64, _ptr, _ptr, _ptr, _i64, _ptr]),
    "ab_zero_out_row_grad": (C.c_int, [_i64, _ptr, _ptr, _i64, _ptr, _i64, _ptr]),
    "ab_scatter_update": (C.c_int, [_i64, _ptr, _ptr, _ptr, _i64, _ptr, _ptr, _ptr, _i64, _i64, _ptr, _i64, _ptr, _i64, _ptr]),
    "ab_scatter_update_grad": (C.c_int, [_i64, _ptr, _ptr, _i64, _i64, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64, _ptr, _ptr]),
    "ab_sparse_indexing": (C.c_int, [_i64, _ptr, _i64, _ptr, _i64, _ptr]),
    "ab_sparse_indexing_grad": (C.c_int, [_i64, _ptr, _i64, _ptr, _i64, _ptr, _ptr, _ptr, _i64, _ptr]),
    "ab_csr_export_coo": (C.c_int, [_i64, _ptr, _ptr, _ptr, C.c_int]),
    "ab_coo_scale": (C.c_int, [_i64, _ptr, _ptr, _ptr, _i64, C.c_int, _ptr]),
    "ab_dense_to_coo": (C.c_int, [_i64, _i64, _ptr, _ptr]),
    "ab_coo_concat": (C.c_int, [_i64, _ptr, _ptr, _ptr, _i64, _ptr, _ptr, _ptr, _i64, _i64, C.c_int, _ptr, _ptr, _ptr]),
    "ab_csr_to_dense": (C.c_int, [_i64, _ptr]),
    "ab_dense_grad_values": (C.c_int, [_i64, _ptr, _i64, _i64, _ptr, _ptr]),
    "ab_csr_axpby": (C.c_int, [_i64, _dbl, _i64, _dbl, _ptr]),
    "ab_spgemm": (C.c_int, [_i64, _i64, C.c_int, _ptr]),
    "ab_tri_solve": (C.c_int, [_i64, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "ab_tri_solve_grad": (C.c_int, [_i64, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "ab_newton_step": (C.c_int, [_i64, _str, _ptr, _ptr, _dbl, _ptr, _ptr, _ptr, _dbl, C.c_int, _ptr]),
    "ab_vec_axpby": (C.c_int, [_i64, _dbl, _ptr, _dbl, _ptr, _ptr]),
    "ab_vec_dot": (C.c_int, [_i64, _ptr, _ptr, _ptr]),
    "ab_poisson3d_csr": (C.c_int, [_i64, _i64, _i64, _i64, _i64, _ptr]),
    "ab_poisson3d_kappa_csr": (C.c_int, [_i64, _i64, _i64, _ptr, _i64, _i64, _ptr]),
    "ab_poisson2d_csr": (C.c_int, [_i64, _ptr, C.c_int, _ptr]),
    "ab_poisson2d_update": (C.c_int, [_i64, _i64, _ptr, C.c_int]),
    "ab_poisson2d_grad": (C.c_int, [_i64, _ptr, _ptr, C.c_int, _ptr]),
    "ab_dist_unique_id": (C.c_int, [_ptr]),
    "ab_dist_init": (C.c_int, [C.c_int, C.c_int, _ptr]),
    "ab_dist_finalize": (C.c_int, []),
    "ab_dist_csr_create": (C.c_int, [_i64, _i64, _i64, _i64, _ptr]),
    "ab_dist_spmv": (C.c_int, [_i64, _ptr, _ptr]),
    "ab_dist_solve": (C.c_int, [_i64, _str, _ptr, _ptr, _dbl, C.c_int, _ptr, _ptr]),
    "ab_dist_solve_grad": (C.c_int, [_i64, _ptr, _ptr, _ptr, _ptr, _dbl, C.c_int, _ptr, _ptr]),
    "ab_dist_cg_fixed": (C.c_int, [_i64, _ptr, _ptr, C.c_int, _ptr]),
    "ab_dist_destroy": (C.c_int, [_i64]),
    "ab_dist_export_coo": (C.c_int, [_i64, _ptr, _ptr]),
    "ab_dist_query": (C.c_int, [_i64, _ptr, _ptr, _ptr]),
    "ab_dist_transpose": (C.c_int, [_i64, _ptr]),
    "ab_timer_reset": (C.c_int, []),
    "ab_timer_get": (_dbl, [C.c_int]),
    "ab_bench_spmv": (C.c_int, [_i64, _ptr, _ptr, C.c_int, C.c_int, _ptr]),
    "ab_bench_cg_kernels": (C.c_int, [_i64, C.c_int, _ptr, _ptr, _ptr]),
    "ab_launch_count": (_i64, []),
    "ab_spmv_solver_form": (C.c_int, [_i64, _ptr, _ptr]),
    "ab_csr_get_info": (C.c_int, [_i64, _str, _ptr]),
    "ab_test_zmarch_plan": (C.c_int, [_i64, _i64, C.c_int, _ptr, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ptr, _i64,
                                      _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr]),
}


class ADCMEError(RuntimeError):
    """Raised for every non-zero return of the C-ABI (the reference raises PyCall.PyError from a
    TensorFlow Status, /root/reference/test/sparse.jl:335-339)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[adcme_b200 {code}] {msg}")
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python adcme.jl_b200/build.py` "
            "(libadcme_b200 has no Python or CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError = header/library drift; let it surface
        fn.restype, fn.argtypes = res, args
    return lib


lib = _load()


def check(rc: int) -> None:
    if rc != AB_OK:
        raise ADCMEError(rc, (lib.ab_last_error() or b"").decode("utf-8", "replace"))


# ------------------------------------------------------------------ argument marshalling
def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def as_array(x, dtype):
    """Returns (object kept alive, pointer).  torch tensors (CPU or CUDA) pass their storage
    pointer; everything else goes through a contiguous numpy array of `dtype`."""
    if x is None:
        return None, None
    if _is_torch(x):
        import torch

        tdt = {np.float64: torch.float64, np.int64: torch.int64, np.int32: torch.int32}[dtype]
        t = x.detach()
        if t.dtype != tdt:
            t = t.to(tdt)
        t = t.contiguous()
        return t, C.c_void_p(t.data_ptr())
    a = np.ascontiguousarray(x, dtype=dtype)
    return a, C.c_void_p(a.ctypes.data)


def empty_like_kind(ref, shape, dtype):
    """Output buffer living where `ref` lives (torch device tensor / torch cpu / numpy)."""
    if _is_torch(ref):
        import torch

        tdt = {np.float64: torch.float64, np.int64: torch.int64, np.int32: torch.int32}[dtype]
        return torch.empty(shape, dtype=tdt, device=ref.device)
    return np.empty(shape, dtype=dtype)


def ptr(x):
    if x is None:
        return None
    if _is_torch(x):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(x.ctypes.data)
