"""Host-side mirror of bixverse's operator interface for the co-expression
correlation path, calling the CUDA library through its C ABI.

Function names, argument meaning and error behaviour follow the reference's
extendr functions so that tests read like the reference's own tinytest files:

=========================== =================================================
here                        reference
=========================== =================================================
rs_cor                      src/rust/src/base/r_cors_similarity.rs:70-77
rs_cor_upper_triangle       src/rust/src/base/r_cors_similarity.rs:251-268
rs_differential_cor         src/rust/src/methods/r_diffcor.rs:45-74
rs_tom                      src/rust/src/methods/r_coremo.rs:46-59
rs_rank_matrix_col          src/rust/src/enrichment/r_singscore.rs:89-95
rs_rbf_function[_mat]       src/rust/src/base/r_rbf.rs:36-83
rs_rbf_iterate_epsilons     src/rust/src/base/r_rbf.rs:118-131
rs_coremo_stability         src/rust/src/methods/r_coremo.rs:191-235
rs_pairwise_gene_cors_mc    src/rust/src/meta_cell/r_mc_processing.rs:258-283
rs_upper_triangle_to_dense  src/rust/src/base/r_helpers.rs:75-94
rs_dense_to_upper_triangle  src/rust/src/base/r_helpers.rs:113-134
rs_upper_triangle_to_sparse src/rust/src/data/r_sparse.rs:43-55
calculate_tom               R/functions_stats.R:43-59
calculate_tom_from_exp      R/functions_stats.R:103-121
cor_module_tom              R/methods_coexp_cor.R:118-164
UpperTriangularSymMat       R/classes_triangular_mat.R:10-211
UpperTriangleDiffcorMat     R/classes_triangular_mat.R:224-341
=========================== =================================================

Matrices are samples x genes; they are handed to the library in R's
column-major layout (``np.asfortranarray``).  Extra keyword arguments
(``precision``, ``beta``) are the north-star extensions and default to the
reference behaviour: ``precision`` defaults to ``PREC_AUTO``, the value the
drop-in patch of INTEGRATION.md passes (exact-integer Gram for Spearman with
n <= 1860, FP64 otherwise; env ``BIXCOR_PRECISION`` overrides).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import PREC_AUTO, PREC_F16X3, PREC_F64, PREC_I8EXACT, PREC_OZAKI, PREC_OZAKI4, PREC_TF32X3, TOM_V1, TOM_V2, BixcorError, check

__all__ = [
    "rs_cor", "rs_cor_upper_triangle", "rs_covariance", "rs_cos", "rs_cor2", "rs_cov2cor", "rs_differential_cor", "rs_tom", "rs_rank_matrix_col",
    "rs_upper_triangle_to_dense", "rs_dense_to_upper_triangle", "rs_upper_triangle_to_sparse", "calculate_tom", "calculate_tom_from_exp", "cor_module_tom", "rs_rbh_cor",
    "rs_rbf_function", "rs_rbf_function_mat", "rs_rbf_iterate_epsilons", "rs_coremo_stability", "cor_upper_triangle_rbf", "rs_pairwise_gene_cors_mc",
    "sparse_gene_cor_upper_triangle", "UpperTriangularSymMat", "UpperTriangleDiffcorMat", "init", "shard_rows",
    "PREC_AUTO", "PREC_F64", "PREC_TF32X3", "PREC_I8EXACT", "PREC_OZAKI", "PREC_OZAKI4", "PREC_F16X3", "BixcorError",
]


def init(n_gpus: int = 1, first_device: int = 0) -> None:
    check(_lib.load().bixcor_init_devices(int(first_device), int(n_gpus)))


def _fmat(x, name="x") -> np.ndarray:
    a = np.asarray(x)
    if a.ndim != 2:
        raise ValueError(f"{name} must be a matrix")
    return np.asfortranarray(a, dtype=np.float64)


def _p(a: np.ndarray):
    return a.ctypes.data


def _parse_tom_types(tom_type: str) -> int:
    # crate parse_tom_types: "v1" / "v2", anything else is an error (r_coremo.rs:50-52)
    if tom_type == "v1":
        return TOM_V1
    if tom_type == "v2":
        return TOM_V2
    raise ValueError(f"Invalid TOM version type: {tom_type}")


def shard_rows(p: int, shift: bool, world: int, rank: int) -> tuple[int, int]:
    import ctypes as C

    r0, r1 = C.c_int64(), C.c_int64()
    check(_lib.load().bixcor_shard_rows(p, int(bool(shift)), world, rank, C.byref(r0), C.byref(r1)))
    return r0.value, r1.value


def rs_rank_matrix_col(x) -> np.ndarray:
    a = _fmat(x)
    out = np.empty(a.shape, dtype=np.float64, order="F")
    check(_lib.load().bixcor_rank_columns(_p(a), a.shape[0], a.shape[1], _p(out)))
    return out


def rs_cor(x, spearman: bool, precision: int = PREC_AUTO) -> np.ndarray:
    a = _fmat(x)
    n, p = a.shape
    out = np.empty((p, p), dtype=np.float64, order="F")
    check(_lib.load().bixcor_cor_dense(_p(a), n, p, int(bool(spearman)), precision, _p(out)))
    return out


def rs_covariance(x) -> np.ndarray:
    """src/rust/src/base/r_cors_similarity.rs:49-55 (== base R cov)."""
    a = _fmat(x)
    n, p = a.shape
    out = np.empty((p, p), dtype=np.float64, order="F")
    check(_lib.load().bixcor_covariance(_p(a), n, p, _p(out)))
    return out


def rs_cos(x) -> np.ndarray:
    """src/rust/src/base/r_cors_similarity.rs:90-97."""
    a = _fmat(x)
    n, p = a.shape
    out = np.empty((p, p), dtype=np.float64, order="F")
    check(_lib.load().bixcor_cos(_p(a), n, p, _p(out)))
    return out


def rs_cor2(x, y, spearman: bool) -> np.ndarray:
    """src/rust/src/base/r_cors_similarity.rs:114-122 (== base R cor(x, y)); rows must match."""
    a, b = _fmat(x), _fmat(y, "y")
    if a.shape[0] != b.shape[0]:
        raise ValueError("x and y need the same number of rows")
    out = np.empty((a.shape[1], b.shape[1]), dtype=np.float64, order="F")
    check(_lib.load().bixcor_cor2(_p(a), a.shape[0], a.shape[1], _p(b), b.shape[1], int(bool(spearman)), _p(out)))
    return out


def rs_rbh_cor(module_matrices: dict, k_best: int, spearman: bool, min_similarity: float) -> dict:
    """src/rust/src/methods/r_rbh.rs:187-272.  module_matrices: name -> (matrix features x modules, rownames, colnames)
    (the named R matrices of the list).  Every origin is compared with every later target on their shared features;
    returns the reference's list: origin, target, comparisons, origin_modules, target_modules, similarity."""
    lib = _lib.load()
    names = list(module_matrices)
    res = {k: [] for k in ("origin", "target", "comparisons", "origin_modules", "target_modules", "similarity")}
    for a, na in enumerate(names[:-1]):
        xa, ra, ca = module_matrices[na]
        xa = np.asarray(xa, dtype=np.float64)
        for nb in names[a + 1:]:
            xb, rb, cb = module_matrices[nb]
            xb = np.asarray(xb, dtype=np.float64)
            pos_b = {f: k for k, f in enumerate(rb)}
            shared = [(k, pos_b[f]) for k, f in enumerate(ra) if f in pos_b]
            n_hits = 0
            if len(shared) >= 2:
                ia, ib = (list(t) for t in zip(*shared))
                sa, sb = _fmat(xa[ia]), _fmat(xb[ib], "y")
                cap = sa.shape[1] * sb.shape[1]
                oi, oj = np.empty(cap, dtype=np.int32), np.empty(cap, dtype=np.int32)
                sim = np.empty(cap, dtype=np.float64)
                cnt = np.zeros(1, dtype=np.int64)
                check(lib.bixcor_rbh_cor(_p(sa), sa.shape[0], sa.shape[1], _p(sb), sb.shape[1], int(k_best), int(bool(spearman)),
                                         float(min_similarity), _p(oi), _p(oj), _p(sim), cap, _p(cnt)))
                n_hits = int(cnt[0])
                res["origin_modules"] += [ca[i] for i in oi[:n_hits]]
                res["target_modules"] += [cb[j] for j in oj[:n_hits]]
                res["similarity"] += sim[:n_hits].tolist()
            res["origin"].append(na)
            res["target"].append(nb)
            res["comparisons"].append(n_hits)
    return res


def rs_cov2cor(x) -> np.ndarray:
    """src/rust/src/base/r_cors_similarity.rs:135-142."""
    a = _fmat(x)
    if a.shape[0] != a.shape[1]:
        raise ValueError("x must be a square covariance matrix")
    out = np.empty(a.shape, dtype=np.float64, order="F")
    check(_lib.load().bixcor_cov2cor(_p(a), a.shape[0], _p(out)))
    return out


def rs_cor_upper_triangle(x, spearman: bool, shift: bool = True, beta: float = 1.0,
                          precision: int = PREC_AUTO, rows: tuple[int, int] | None = None) -> np.ndarray:
    a = _fmat(x)
    n, p = a.shape
    lib = _lib.load()
    s = int(bool(shift))
    if rows is None:
        out = np.empty(lib.bixcor_packed_len(p, s), dtype=np.float64)
        check(lib.bixcor_cor_upper(_p(a), n, p, int(bool(spearman)), s, float(beta), precision, _p(out)))
    else:
        r0, r1 = rows
        ln = lib.bixcor_packed_offset(r1, p, s) - lib.bixcor_packed_offset(r0, p, s)
        out = np.empty(max(ln, 0), dtype=np.float64)
        check(lib.bixcor_cor_upper_rows(_p(a), n, p, int(bool(spearman)), s, float(beta), precision, r0, r1, _p(out)))
    return out


def rs_differential_cor(x_a, x_b, spearman: bool, precision: int = PREC_AUTO) -> dict:
    a, b = _fmat(x_a, "x_a"), _fmat(x_b, "x_b")
    if a.shape[1] != b.shape[1]:
        # the reference asserts (panics) here: r_diffcor.rs:51-56
        raise ValueError(
            "Input matrices must have the same number of columns. Found "
            f"{a.shape[1]} columns in first matrix and {b.shape[1]} in second."
        )
    p = a.shape[1]
    lib = _lib.load()
    ln = lib.bixcor_packed_len(p, 1)
    res = {k: np.empty(ln, dtype=np.float64) for k in ("r_a", "r_b", "z_score", "p_val")}
    check(lib.bixcor_diffcor(_p(a), a.shape[0], _p(b), b.shape[0], p, int(bool(spearman)), precision,
                             _p(res["r_a"]), _p(res["r_b"]), _p(res["z_score"]), _p(res["p_val"])))
    return res


def rs_tom(x, tom_type: str, signed: bool, precision: int = PREC_AUTO) -> np.ndarray:
    version = _parse_tom_types(tom_type)
    a = _fmat(x)
    if a.shape[0] != a.shape[1]:
        raise ValueError("x must be a square affinity matrix")
    p = a.shape[0]
    out = np.empty((p, p), dtype=np.float64, order="F")
    check(_lib.load().bixcor_tom(_p(a), p, int(bool(signed)), version, precision, _p(out)))
    return out


def calculate_tom(cor_mat, signed: bool, version: str = "v1", precision: int = PREC_AUTO) -> np.ndarray:
    if version not in ("v1", "v2"):
        raise ValueError("version must be one of 'v1', 'v2'")
    cor_mat = np.asarray(cor_mat, dtype=np.float64)
    if not signed:
        cor_mat = np.abs(cor_mat)  # R/functions_stats.R:52-54
    return rs_tom(cor_mat, version, signed, precision)


def cor_module_tom(cor_res: "UpperTriangularSymMat", signed: bool = True, version: str = "v2",
                   precision: int = PREC_AUTO) -> "UpperTriangularSymMat":
    """R/methods_coexp_cor.R:118-164 on the result container: get_sym_matrix -> rs_tom -> rs_dense_to_upper_triangle,
    fused into one library call; returns the TOM as a new shift = 1 container (the R method stores it back)."""
    ver = _parse_tom_types(version)
    p = len(cor_res.features)
    lib = _lib.load()
    out = np.empty(lib.bixcor_packed_len(p, 1), dtype=np.float64)
    check(lib.bixcor_tom_packed(_p(cor_res.values), int(cor_res.shift), p, int(bool(signed)), ver, precision, _p(out)))
    return UpperTriangularSymMat(out, cor_res.features, True)


def calculate_tom_from_exp(x, signed: bool, version: str, cor_method: str = "pearson", beta: float = 1.0,
                           precision: int = PREC_AUTO, packed: bool = False) -> np.ndarray:
    """Fused on-device rs_cor -> (abs) -> rs_tom chain; the p x p correlation never leaves HBM."""
    if version not in ("v1", "v2"):
        raise ValueError("version must be one of 'v1', 'v2'")
    a = _fmat(x)
    n, p = a.shape
    lib = _lib.load()
    out = np.empty(lib.bixcor_packed_len(p, 1), dtype=np.float64) if packed else np.empty((p, p), dtype=np.float64, order="F")
    check(lib.bixcor_tom_from_exp(_p(a), n, p, int(cor_method == "spearman"), int(bool(signed)),
                                  _parse_tom_types(version), float(beta), precision, int(bool(packed)), _p(out)))
    return out


def rs_upper_triangle_to_dense(data, shift: bool, n: int) -> np.ndarray:
    v = np.ascontiguousarray(data, dtype=np.float64)
    lib = _lib.load()
    if v.shape[0] != lib.bixcor_packed_len(n, int(bool(shift))):
        raise ValueError("packed length does not match n")
    out = np.empty((n, n), dtype=np.float64, order="F")
    check(lib.bixcor_upper_to_dense(_p(v), int(bool(shift)), n, _p(out)))
    return out


def rs_dense_to_upper_triangle(x, shift: bool) -> np.ndarray:
    a = _fmat(x)
    p = a.shape[1]
    lib = _lib.load()
    out = np.empty(lib.bixcor_packed_len(p, int(bool(shift))), dtype=np.float64)
    check(lib.bixcor_dense_to_upper(_p(a), int(bool(shift)), p, _p(out)))
    return out


def _parse_rbf_types(rbf_type: str) -> int:
    # crate parse_rbf_types; anything else is "Invalid RBF function" (r_rbf.rs:37-38)
    try:
        return {"gaussian": _lib.RBF_GAUSSIAN, "bump": _lib.RBF_BUMP, "inverse_quadratic": _lib.RBF_INVERSE_QUADRATIC}[rbf_type]
    except KeyError:
        raise ValueError(f"Invalid RBF function: {rbf_type}") from None


def rs_rbf_function(x, epsilon: float, rbf_type: str) -> np.ndarray:
    """src/rust/src/base/r_rbf.rs:36-48."""
    mode = _parse_rbf_types(rbf_type)
    v = np.ascontiguousarray(x, dtype=np.float64).ravel()
    out = np.empty_like(v)
    check(_lib.load().bixcor_rbf(_p(v), v.shape[0], float(epsilon), mode, _p(out)))
    return out


def rs_rbf_function_mat(x, epsilon: float, rbf_type: str) -> np.ndarray:
    """src/rust/src/base/r_rbf.rs:66-83."""
    mode = _parse_rbf_types(rbf_type)
    a = _fmat(x)
    out = np.empty(a.shape, dtype=np.float64, order="F")
    check(_lib.load().bixcor_rbf(_p(a), a.size, float(epsilon), mode, _p(out)))
    return out


def rs_rbf_iterate_epsilons(dist, epsilon_vec, original_dim: int, shift: bool, rbf_type: str) -> np.ndarray:
    """src/rust/src/base/r_rbf.rs:118-131; an unknown rbf_type falls back to gaussian like the reference doc says."""
    mode = {"bump": _lib.RBF_BUMP, "inverse_quadratic": _lib.RBF_INVERSE_QUADRATIC}.get(rbf_type, _lib.RBF_GAUSSIAN)
    v = np.ascontiguousarray(dist, dtype=np.float64)
    eps = np.ascontiguousarray(epsilon_vec, dtype=np.float64)
    lib = _lib.load()
    if v.shape[0] != lib.bixcor_packed_len(original_dim, int(bool(shift))):
        raise ValueError("packed length does not match original_dim")
    out = np.empty((original_dim, eps.shape[0]), dtype=np.float64, order="F")
    check(lib.bixcor_rbf_iterate_epsilons(_p(v), original_dim, int(bool(shift)), _p(eps), eps.shape[0], mode, _p(out)))
    return out


def rs_coremo_stability(data, indices, epsilon: float, rbf_type: str, spearman: bool,
                        precision: int = PREC_AUTO) -> list[np.ndarray]:
    """src/rust/src/methods/r_coremo.rs:191-235: one packed distance vector per left-out (1-based) sample."""
    import ctypes as C

    mode = _parse_rbf_types(rbf_type)
    a = _fmat(data)
    n, p = a.shape
    idx = np.ascontiguousarray(indices, dtype=np.int32)
    lib = _lib.load()
    outs = [np.empty(lib.bixcor_packed_len(p, 1), dtype=np.float64) for _ in range(idx.shape[0])]
    ptrs = (C.c_void_p * max(1, idx.shape[0]))(*[o.ctypes.data for o in outs])
    check(lib.bixcor_coremo_stability(_p(a), n, p, _p(idx), idx.shape[0], float(epsilon), mode, int(bool(spearman)),
                                      precision, C.cast(ptrs, C.c_void_p)))
    return outs


def cor_upper_triangle_rbf(x, spearman: bool, epsilon: float, rbf_type: str, shift: bool = True,
                           complement: bool = False, precision: int = PREC_AUTO) -> np.ndarray:
    """Fused rs_cor_upper_triangle -> 1 - |r| -> rs_rbf_function (R/methods_coexp_cor.R:357-370) in one Gram pass."""
    mode = _parse_rbf_types(rbf_type)
    a = _fmat(x)
    n, p = a.shape
    lib = _lib.load()
    out = np.empty(lib.bixcor_packed_len(p, int(bool(shift))), dtype=np.float64)
    check(lib.bixcor_cor_upper_rbf(_p(a), n, p, int(bool(spearman)), int(bool(shift)), float(epsilon), mode,
                                   int(bool(complement)), precision, _p(out)))
    return out


def rs_upper_triangle_to_sparse(value, shift: bool, n: int, cs_type: str = "csc") -> dict:
    """src/rust/src/data/r_sparse.rs:43-55: the reference's sparse list (data, indices 0-based, indptr, nrow, ncol, cs_type)."""
    if cs_type not in ("csr", "csc"):
        raise ValueError("Invalid cs_type")
    v = np.ascontiguousarray(value, dtype=np.float64)
    lib = _lib.load()
    if v.shape[0] != lib.bixcor_packed_len(n, int(bool(shift))):
        raise ValueError("packed length does not match n")
    nnz = lib.bixcor_upper_to_sparse_nnz(_p(v), int(bool(shift)), n)
    indptr = np.empty(n + 1, dtype=np.int64)
    indices = np.empty(max(nnz, 1), dtype=np.int32)
    data = np.empty(max(nnz, 1), dtype=np.float64)
    check(lib.bixcor_upper_to_sparse(_p(v), int(bool(shift)), n, _p(indptr), _p(indices), _p(data)))
    return {"data": data[:nnz], "indices": indices[:nnz], "indptr": indptr, "nrow": n, "ncol": n, "cs_type": cs_type}


def sparse_gene_cor_upper_triangle(indptr, indices, data, cells: int, genes: int, precision: int = PREC_AUTO) -> np.ndarray:
    """All-pairs Pearson over a cells x genes CSR matrix (north-star extension of
    rs_pairwise_gene_cors_mc, src/rust/src/meta_cell/r_mc_processing.rs:258-283)."""
    ip = np.ascontiguousarray(indptr, dtype=np.int64)
    ix = np.ascontiguousarray(indices, dtype=np.int32)
    dt = np.ascontiguousarray(data, dtype=np.float32)
    if ip.shape[0] != cells + 1:
        raise ValueError("indptr must have cells + 1 entries")
    lib = _lib.load()
    out = np.empty(lib.bixcor_packed_len(genes, 1), dtype=np.float64)
    check(lib.bixcor_sparse_gram_cor(_p(ip), _p(ix), _p(dt), cells, genes, precision, _p(out)))
    return out


def rs_pairwise_gene_cors_mc(sparse_data: dict, gene_indices_1, gene_indices_2, spearman: bool,
                             cells_to_keep=None) -> np.ndarray:
    """src/rust/src/meta_cell/r_mc_processing.rs:258-283.  `sparse_data` is the reference's sparse list:
    ``data``, ``indptr``, ``indices`` (0-based), ``nrow`` (cells), ``ncol`` (genes) and ``format`` ("csr"/"csc");
    gene indices are 0-based, like the reference demands.  With ``cells_to_keep`` (0-based cell indices) this is the
    arithmetic of rs_pairwise_gene_cors (src/rust/src/single_cell/r_sc_processing.rs:474-499) on gene columns the
    caller has read from ``counts_genes.bin``."""
    for k in ("data", "indptr", "indices", "nrow", "ncol", "format"):
        if k not in sparse_data:
            raise ValueError(f"sparse_data is missing '{k}'")
    fmt = str(sparse_data["format"]).lower()
    if fmt not in ("csr", "csc"):
        raise ValueError(f"unknown sparse format: {sparse_data['format']}")
    cells, genes = int(sparse_data["nrow"]), int(sparse_data["ncol"])
    ip = np.ascontiguousarray(sparse_data["indptr"], dtype=np.int64)
    ix = np.ascontiguousarray(sparse_data["indices"], dtype=np.int32)
    dt = np.ascontiguousarray(sparse_data["data"], dtype=np.float32)  # the reference casts to f32 (:267-269)
    if ip.shape[0] != (cells if fmt == "csr" else genes) + 1:
        raise ValueError("indptr length does not match the sparse format")
    g1 = np.ascontiguousarray(gene_indices_1, dtype=np.int32)
    g2 = np.ascontiguousarray(gene_indices_2, dtype=np.int32)
    if g1.shape != g2.shape:
        raise ValueError("gene_indices_1 and gene_indices_2 need the same length")
    out = np.empty(g1.shape[0], dtype=np.float64)
    if cells_to_keep is not None:
        keep = np.ascontiguousarray(cells_to_keep, dtype=np.int32)
        check(_lib.load().bixcor_sparse_pairwise_cor_cells(_p(ip), _p(ix), _p(dt), cells, genes, int(fmt == "csr"), _p(keep),
                                                           keep.shape[0], _p(g1), _p(g2), g1.shape[0], int(bool(spearman)),
                                                           _p(out)))
        return out
    check(_lib.load().bixcor_sparse_pairwise_cor(_p(ip), _p(ix), _p(dt), cells, genes, int(fmt == "csr"), _p(g1), _p(g2),
                                                 g1.shape[0], int(bool(spearman)), _p(out)))
    return out


class UpperTriangularSymMat:
    """Mirror of the R6 class upper_triangular_sym_mat (R/classes_triangular_mat.R:10-211)."""

    def __init__(self, values, features, shift: bool = True):
        self.values = np.ascontiguousarray(values, dtype=np.float64)
        self.features = list(features)
        self.shift = bool(shift)
        p = len(self.features)
        expected = p * (p - 1) // 2 if self.shift else p * (p + 1) // 2
        if self.values.shape[0] != expected:  # length check, classes_triangular_mat.R:33-38
            raise ValueError(f"expected {expected} values for {p} features, got {self.values.shape[0]}")

    def get_sym_matrix(self) -> np.ndarray:
        return rs_upper_triangle_to_dense(self.values, self.shift, len(self.features))

    def get_sparse_matrix(self) -> dict:
        """R/classes_triangular_mat.R:125-140 (the sparse list; Matrix::sparseMatrix is the R side's business)."""
        return rs_upper_triangle_to_sparse(self.values, self.shift, len(self.features), "csc")


class UpperTriangleDiffcorMat:
    """Mirror of R/classes_triangular_mat.R:224-341 (shift fixed to 1)."""

    REQUIRED = ("r_a", "r_b", "z_score", "p_val")

    def __init__(self, diff_cor_res: dict, features):
        missing = [k for k in self.REQUIRED if k not in diff_cor_res]
        if missing:
            raise ValueError(f"diff_cor_res is missing {missing}")
        lens = {len(diff_cor_res[k]) for k in self.REQUIRED}
        if len(lens) != 1:
            raise ValueError("all differential correlation vectors must have equal length")
        self.features = list(features)
        p = len(self.features)
        if lens.pop() != p * (p - 1) // 2:
            raise ValueError("vector length does not match the number of features")
        self.data = {k: np.asarray(diff_cor_res[k], dtype=np.float64) for k in self.REQUIRED}

    def get_cor_matrix(self, which: str = "r_a") -> np.ndarray:
        return rs_upper_triangle_to_dense(self.data[which], True, len(self.features))

    def get_data_table(self) -> dict:
        p = len(self.features)
        i, j = np.triu_indices(p, 1)
        keep = self.data["p_val"] != 0  # classes_triangular_mat.R:298-300
        out = {"feature_a": np.asarray(self.features)[i][keep], "feature_b": np.asarray(self.features)[j][keep]}
        out.update({k: v[keep] for k, v in self.data.items()})
        return out
