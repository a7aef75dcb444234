"""ctypes binding of libbixcor.so (the C-ABI of include/bixcor.h).

There is no CPU fallback: a missing library is an ImportError-level failure and
a missing GPU surfaces as ``BixcorError`` from the first compute call.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BIXCOR_LIB: developer override to load an ablation build (make -C csrc VARIANT=... EXTRA=...), see scripts/perf_probe.py
LIB_PATH = os.environ.get("BIXCOR_LIB") or os.path.join(_HERE, "libbixcor.so")

OK = 0
ERR_NO_GPU, ERR_INVALID, ERR_CUDA, ERR_OOM, ERR_UNSUPPORTED, ERR_INTERNAL = -1, -2, -3, -4, -5, -6
PREC_F64, PREC_TF32X3, PREC_I8EXACT, PREC_OZAKI, PREC_OZAKI4, PREC_F16X3 = 0, 1, 2, 3, 4, 5
PREC_AUTO = -1  # what the drop-in patch passes: exact-integer Gram for Spearman with n <= 1860, FP64 otherwise (env BIXCOR_PRECISION)
TOM_V1, TOM_V2 = 1, 2
RBF_GAUSSIAN, RBF_BUMP, RBF_INVERSE_QUADRATIC = 1, 2, 3
XCHG_HANDLE_BYTES = 64

i64, i32, f64 = C.c_int64, C.c_int, C.c_double
ptr = C.c_void_p

# name -> (restype, argtypes); every symbol include/bixcor.h declares
SIGNATURES = {
    "bixcor_init": (i32, [i32]),
    "bixcor_init_devices": (i32, [i32, i32]),
    "bixcor_shutdown": (None, []),
    "bixcor_nccl_version": (i32, []),
    "bixcor_last_error": (C.c_char_p, []),
    "bixcor_launch_count": (i64, []),
    "bixcor_transfer_counters": (i32, [C.POINTER(i64), C.POINTER(i64), i32]),
    "bixcor_set_transport": (i32, [i32, i32]),
    "bixcor_host_expand_s32": (i32, [ptr, i64, i64, ptr, i64, i32, i32, ptr]),
    "bixcor_host_widen_f32": (i32, [ptr, i64, ptr]),
    "bixcor_set_fieller": (i32, [f64]),
    "bixcor_debug_inject": (i32, [i32]),
    "bixcor_packed_len": (i64, [i64, i32]),
    "bixcor_packed_offset": (i64, [i64, i64, i32]),
    "bixcor_upper_to_dense": (i32, [ptr, i32, i64, ptr]),
    "bixcor_dense_to_upper": (i32, [ptr, i32, i64, ptr]),
    "bixcor_upper_to_sparse_nnz": (i64, [ptr, i32, i64]),
    "bixcor_upper_to_sparse": (i32, [ptr, i32, i64, ptr, ptr, ptr]),
    "bixcor_shard_rows": (i32, [i64, i32, i32, i32, C.POINTER(i64), C.POINTER(i64)]),
    "bixcor_rank_columns": (i32, [ptr, i64, i64, ptr]),
    "bixcor_cor_dense": (i32, [ptr, i64, i64, i32, i32, ptr]),
    "bixcor_cor_upper": (i32, [ptr, i64, i64, i32, i32, f64, i32, ptr]),
    "bixcor_cor_upper_rows": (i32, [ptr, i64, i64, i32, i32, f64, i32, i64, i64, ptr]),
    "bixcor_diffcor": (i32, [ptr, i64, ptr, i64, i64, i32, i32, ptr, ptr, ptr, ptr]),
    "bixcor_diffcor_rows": (i32, [ptr, i64, ptr, i64, i64, i32, i32, i64, i64, ptr, ptr, ptr, ptr]),
    "bixcor_tom": (i32, [ptr, i64, i32, i32, i32, ptr]),
    "bixcor_tom_packed": (i32, [ptr, i32, i64, i32, i32, i32, ptr]),
    "bixcor_tom_from_exp": (i32, [ptr, i64, i64, i32, i32, i32, f64, i32, i32, ptr]),
    "bixcor_covariance": (i32, [ptr, i64, i64, ptr]),
    "bixcor_cos": (i32, [ptr, i64, i64, ptr]),
    "bixcor_cor2": (i32, [ptr, i64, i64, ptr, i64, i32, ptr]),
    "bixcor_rbh_cor": (i32, [ptr, i64, i64, ptr, i64, i32, i32, f64, ptr, ptr, ptr, i64, ptr]),
    "bixcor_cov2cor": (i32, [ptr, i64, ptr]),
    "bixcor_rbf": (i32, [ptr, i64, f64, i32, ptr]),
    "bixcor_rbf_iterate_epsilons": (i32, [ptr, i64, i32, ptr, i32, i32, ptr]),
    "bixcor_cor_upper_rbf": (i32, [ptr, i64, i64, i32, i32, f64, i32, i32, i32, ptr]),
    "bixcor_coremo_stability": (i32, [ptr, i64, i64, ptr, i32, f64, i32, i32, i32, ptr]),
    "bixcor_sparse_gram_cor": (i32, [ptr, ptr, ptr, i64, i64, i32, ptr]),
    "bixcor_sparse_pairwise_cor": (i32, [ptr, ptr, ptr, i64, i64, i32, ptr, ptr, i64, i32, ptr]),
    "bixcor_sparse_pairwise_cor_cells": (i32, [ptr, ptr, ptr, i64, i64, i32, ptr, i64, ptr, ptr, i64, i32, ptr]),
    "bixcor_dev_rank_columns": (i32, [ptr, i64, i64, ptr]),
    "bixcor_dev_cor_upper_rows": (i32, [ptr, i64, i64, i32, i32, f64, i32, i64, i64, ptr]),
    "bixcor_dev_operand_row_bytes": (i64, [i64, i32]),
    "bixcor_dev_build_operand": (i32, [ptr, i64, i64, i32, i32, i64, i64, ptr, ptr]),
    "bixcor_dev_operand_inv_norm": (i32, [ptr, i64, i64, i32, ptr]),
    "bixcor_dev_cor_upper_rows_op": (i32, [ptr, ptr, i64, i64, i32, i32, f64, i64, i64, ptr]),
    "bixcor_dev_cor_upper_rows_op_to_host": (i32, [ptr, ptr, i64, i64, i32, i32, f64, i64, i64, ptr]),
    "bixcor_xchg_create": (i32, [i64, i64, i32, i32, i32, ptr]),
    "bixcor_xchg_open": (i32, [ptr]),
    "bixcor_xchg_release_peers": (i32, []),
    "bixcor_xchg_destroy": (i32, []),
    "bixcor_dev_cor_upper_rows_xchg": (i32, [ptr, i64, i64, i32, i32, f64, i64, i64, ptr]),
    "bixcor_dev_cor_upper_rows_xchg_to_host": (i32, [ptr, i64, i64, i32, i32, f64, i64, i64, ptr]),
    "bixcor_dev_cor_dense": (i32, [ptr, i64, i64, i32, i32, ptr]),
    "bixcor_dev_diffcor_rows": (i32, [ptr, i64, ptr, i64, i64, i32, i32, i64, i64, ptr, ptr, ptr, ptr]),
    "bixcor_dev_affinity_cols": (i32, [ptr, i64, i64, i32, i32, f64, i32, i64, i64, ptr]),
    "bixcor_dev_tom_connectivity": (i32, [ptr, i64, i32, i64, i64, ptr]),
    "bixcor_dev_tom_rows": (i32, [ptr, ptr, i64, i32, i32, i32, i32, i64, i64, ptr]),
    "bixcor_dev_tom_plane_row_bytes": (i64, [i64, i32]),
    "bixcor_dev_tom_split_planes": (i32, [ptr, i64, i32, i64, i64, ptr]),
    "bixcor_dev_tom_rows_planes": (i32, [ptr, ptr, i64, i32, i32, i32, i32, i64, i64, ptr]),
    "bixcor_dev_sparse_gram_accumulate": (i32, [ptr, ptr, ptr, i64, i64, i32, ptr, ptr]),
    "bixcor_dev_sparse_gram_finalise": (i32, [ptr, ptr, i64, i64, ptr]),
    "bixcor_dev_set_stream": (i32, [ptr, i32]),
    "bixcor_dev_set_async": (i32, [i32]),
    "bixcor_dev_synchronize": (i32, []),
    "bixcor_timing_enable": (i32, [i32]),
    "bixcor_last_timing": (i32, [C.POINTER(f64)]),
}


class BixcorError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libbixcor error {code}: {message}")
        self.code = code
        self.message = message


_lib = None


def load() -> C.CDLL:
    """Load libbixcor.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C bixverse_b200/csrc`. bixverse_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != OK:
        raise BixcorError(rc, load().bixcor_last_error().decode("utf-8", "replace"))


def last_timing() -> dict:
    arr = (f64 * 4)()
    check(load().bixcor_last_timing(arr))
    return {"cols_ms": arr[0], "gemm_ms": arr[1], "other_ms": arr[2], "gemm_launches": int(arr[3])}
