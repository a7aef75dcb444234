"""ctypes view of oracle/liboracle.so (C restatement).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "liboracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_PATH)
    return _lib


def _f(a):
    return np.asfortranarray(a, dtype=np.float64)


def rank_matrix_col(x):
    x = _f(x); out = np.empty_like(x, order="F")
    lib().orc_rank_matrix_col(C.c_void_p(x.ctypes.data), C.c_int64(x.shape[0]), C.c_int64(x.shape[1]), C.c_void_p(out.ctypes.data))
    return out


def cor_dense(x, spearman):
    x = _f(x); p = x.shape[1]; out = np.empty((p, p), order="F")
    lib().orc_cor_dense(C.c_void_p(x.ctypes.data), C.c_int64(x.shape[0]), C.c_int64(p), C.c_int(int(spearman)), C.c_void_p(out.ctypes.data))
    return out


def cor_upper(x, spearman, shift):
    x = _f(x); p = x.shape[1]
    out = np.empty(p * (p - 1) // 2 if shift else p * (p + 1) // 2)
    lib().orc_cor_upper(C.c_void_p(x.ctypes.data), C.c_int64(x.shape[0]), C.c_int64(p), C.c_int(int(spearman)), C.c_int(int(shift)), C.c_void_p(out.ctypes.data))
    return out


def tom(a, signed, version):
    a = _f(a); p = a.shape[0]; out = np.empty((p, p), order="F")
    lib().orc_tom(C.c_void_p(a.ctypes.data), C.c_int64(p), C.c_int(int(signed)), C.c_int(int(version)), C.c_void_p(out.ctypes.data))
    return out


def diffcor(ra, rb, na, nb, c):
    ra = np.ascontiguousarray(ra, dtype=np.float64); rb = np.ascontiguousarray(rb, dtype=np.float64)
    z = np.empty_like(ra); pv = np.empty_like(ra)
    lib().orc_diffcor(C.c_void_p(ra.ctypes.data), C.c_void_p(rb.ctypes.data), C.c_int64(ra.shape[0]), C.c_int64(na), C.c_int64(nb), C.c_double(c), C.c_void_p(z.ctypes.data), C.c_void_p(pv.ctypes.data))
    return z, pv


def cor2(x, y, spearman):
    x = _f(x); y = _f(y); out = np.empty((x.shape[1], y.shape[1]), order="F")
    lib().orc_cor2(C.c_void_p(x.ctypes.data), C.c_int64(x.shape[0]), C.c_int64(x.shape[1]), C.c_void_p(y.ctypes.data),
                   C.c_int64(y.shape[1]), C.c_int(int(spearman)), C.c_void_p(out.ctypes.data))
    return out


def rbh_cor(x, y, k_best, spearman, min_similarity=0.0):
    x = _f(x); y = _f(y); cap = x.shape[1] * y.shape[1]
    oi = np.empty(cap, dtype=np.int32); oj = np.empty(cap, dtype=np.int32); sim = np.empty(cap)
    f = lib().orc_rbh_cor
    f.restype = C.c_int64
    cnt = f(C.c_void_p(x.ctypes.data), C.c_int64(x.shape[0]), C.c_int64(x.shape[1]), C.c_void_p(y.ctypes.data),
            C.c_int64(y.shape[1]), C.c_int(int(k_best)), C.c_int(int(spearman)), C.c_double(min_similarity),
            C.c_void_p(oi.ctypes.data), C.c_void_p(oj.ctypes.data), C.c_void_p(sim.ctypes.data), C.c_int64(cap))
    return [(int(oi[t]), int(oj[t]), float(sim[t])) for t in range(cnt)]


def rbf(d, eps, rbf_type):
    d = np.ascontiguousarray(d, dtype=np.float64); out = np.empty_like(d)
    code = {"gaussian": 1, "bump": 2, "inverse_quadratic": 3}[rbf_type]
    lib().orc_rbf(C.c_void_p(d.ctypes.data), C.c_int64(d.size), C.c_double(eps), C.c_int(code), C.c_void_p(out.ctypes.data))
    return out


# ---- timed CPU baseline (bench.py reference arm): OpenMP restatement + threaded OpenBLAS dgemm ----------------------
_blas = None


def openblas_dgemm():
    """Address of cblas_dgemm in the ILP64 OpenBLAS that numpy bundles (scipy_cblas_dgemm64_), plus its thread count.
    The image has no system BLAS; this is the same library `numpy.dot` calls."""
    global _blas
    if _blas is None:
        import glob

        d = os.path.join(os.path.dirname(np.__file__), "..", "numpy.libs")
        cands = sorted(glob.glob(os.path.join(d, "libscipy_openblas64_*.so*")))
        if not cands:
            raise RuntimeError("numpy's bundled OpenBLAS (numpy.libs/libscipy_openblas64_*.so) not found")
        h = C.CDLL(cands[0])
        fn = C.cast(h.scipy_cblas_dgemm64_, C.c_void_p).value
        get = h.scipy_openblas_get_num_threads64_
        get.restype = C.c_int
        _blas = (h, fn, get)
    return _blas[1], int(_blas[2]())


def set_blas_threads(n):
    openblas_dgemm()
    f = _blas[0].scipy_openblas_set_num_threads64_
    f.argtypes = [C.c_int]
    f(int(n))


def use_all_cores():
    """OpenMP + OpenBLAS threads := the cores this process may run on (torchrun exports OMP_NUM_THREADS=1)."""
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().orc_set_num_threads(C.c_int(ncpu))
    set_blas_threads(ncpu)
    return ncpu


def num_threads():
    f = lib().orc_num_threads
    f.restype = C.c_int
    return int(f())


def cor_upper_pipeline(x, spearman, shift, beta, out=None):
    """rs_cor_upper_triangle (+ |r|^beta) on all host cores; returns (packed vector, seconds dict)."""
    x = _f(x); n, p = x.shape
    ln = p * (p - 1) // 2 if shift else p * (p + 1) // 2
    if out is None:
        out = np.empty(ln)
    secs = (C.c_double * 4)()
    fn, _ = openblas_dgemm()
    f = lib().orc_cor_upper_pipeline
    f.restype = C.c_int
    rc = f(C.c_void_p(x.ctypes.data), C.c_int64(n), C.c_int64(p), C.c_int(int(spearman)), C.c_int(int(shift)), C.c_double(beta),
           C.c_void_p(fn), C.c_void_p(out.ctypes.data), secs)
    if rc != 0:
        raise MemoryError("orc_cor_upper_pipeline: allocation failed")
    return out, {"columns_s": secs[0], "dgemm_s": secs[1], "gather_s": secs[2], "total_s": secs[3]}


def tom_from_exp_pipeline(x, spearman, signed, version, beta):
    x = _f(x); n, p = x.shape
    out = np.empty(p * (p - 1) // 2)
    secs = (C.c_double * 4)()
    fn, _ = openblas_dgemm()
    f = lib().orc_tom_from_exp_pipeline
    f.restype = C.c_int
    rc = f(C.c_void_p(x.ctypes.data), C.c_int64(n), C.c_int64(p), C.c_int(int(spearman)), C.c_int(int(signed)), C.c_int(int(version)),
           C.c_double(beta), C.c_void_p(fn), C.c_void_p(out.ctypes.data), secs)
    if rc != 0:
        raise MemoryError("orc_tom_from_exp_pipeline: allocation failed")
    return out, {"affinity_s": secs[0], "dgemm_s": secs[1], "normalise_s": secs[2], "total_s": secs[3]}
