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
