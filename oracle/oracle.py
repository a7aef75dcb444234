"""NumPy float64 restatement of the bixverse co-expression hot path (CPU oracle).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): the product path never
imports this module.

Why a restatement: the arithmetic of the path lives in the third-party crate
``bixverse-rs`` 0.4.4 (crates.io, checksum c858a486...eba6d;
``/root/reference/src/rust/Cargo.toml:29``, ``Cargo.lock:173-176``) which is not
vendored and cannot be built here (no rustc/cargo/R).  Each function below
restates the *published contract* visible at the reference's own call sites,
tests and docs, and is pinned by ``tests/test_oracle.py`` against

* the 4x4 TOM known-answer vectors of
  ``inst/tinytest/test_cor_modules.R:9-46`` (signed/unsigned x v1/v2),
* the triangle <-> dense vectors of ``inst/tinytest/test_sparse.R:5-9`` and the
  in-tree index loops of ``src/rust/src/base/r_helpers.rs:75-134``,
* base-R ``cor()`` equivalence (``inst/tinytest/test_R_rust_equality.R:5-72``)
  with numpy.corrcoef / scipy.stats as the base-R stand-in on the very matrix
  the reference test uses (R's ``set.seed(123); rnorm`` stream is reproduced in
  ``tests/golden/make_golden.py``),
* the reciprocal-best-hit known answers of ``inst/tinytest/test_rbh.R:109-214``
  (Pearson to 1e-7 and Spearman, on the test's own ``set.seed(42)`` matrices).

PARITY UNPINNED (no reference test or golden vector exists): the Fisher-z
z-score / p-value of ``rs_differential_cor`` (constants restated from the
upstream crate's documented method: Fieller 1.06 for Spearman), tie / NaN /
zero-variance behaviour (base-R semantics by the ``== cor()`` contract), the
TOM diagonal (never consumed by the packed path), and the two north-star
extensions that do not exist in the reference (soft-threshold power, all-pairs
sparse Gram).
"""
from __future__ import annotations

import math

import numpy as np

__all__ = [
    "rank_average",
    "rank_matrix_col",
    "column_pairwise_cor",
    "column_pairwise_cov",
    "column_pairwise_cos",
    "cor_two_matrices",
    "cov2cor",
    "upper_triangle_indices",
    "dense_to_upper_triangle",
    "upper_triangle_to_dense",
    "packed_offset",
    "cor_upper_triangle",
    "soft_threshold",
    "calculate_diff_correlation",
    "differential_cor",
    "calc_tom",
    "calculate_tom",
    "tom_from_exp",
    "sparse_gram_cor",
    "pairwise_gene_cors",
    "pairwise_gene_cors_spearman",
    "rbf_function",
    "rbf_iterate_epsilons",
    "coremo_stability",
    "rbh_cor",
    "rbh_cor_list",
    "shard_rows",
]


# --------------------------------------------------------------------------
# ranks  (crate core::math::matrix_helpers::rank_matrix_col, exposed at
# src/rust/src/enrichment/r_singscore.rs:89-95; contract == base R
# rank(ties.method = "average"), inst/tinytest/test_R_rust_equality.R:17-21)
# --------------------------------------------------------------------------
def rank_average(v: np.ndarray) -> np.ndarray:
    """1-based average-tie ranks of one vector; -0.0 ties with +0.0.

    A vector holding any NaN ranks to all-NaN (base-R ``cor`` returns NA for
    such a column; unpinned by the reference tests).
    """
    v = np.asarray(v, dtype=np.float64)
    n = v.shape[0]
    if n == 0:
        return np.empty(0, dtype=np.float64)
    if np.isnan(v).any():
        return np.full(n, np.nan)
    order = np.argsort(v, kind="stable")
    sv = v[order]
    # boundaries of runs of equal values
    new_run = np.empty(n, dtype=bool)
    new_run[0] = True
    new_run[1:] = sv[1:] != sv[:-1]
    run_id = np.cumsum(new_run) - 1
    starts = np.flatnonzero(new_run)
    ends = np.append(starts[1:], n) - 1
    avg = 0.5 * (starts + ends) + 1.0  # mean of 1-based positions in the run
    out = np.empty(n, dtype=np.float64)
    out[order] = avg[run_id]
    return out


def rank_matrix_col(x: np.ndarray) -> np.ndarray:
    """Column-wise average ranks of an n x p matrix."""
    x = np.asarray(x, dtype=np.float64)
    out = np.empty_like(x, order="F")
    for j in range(x.shape[1]):
        out[:, j] = rank_average(x[:, j])
    return out


# --------------------------------------------------------------------------
# correlation (crate core::base::cors_similarity::column_pairwise_cor, call
# site src/rust/src/base/r_cors_similarity.rs:70-77; contract == base R cor())
# --------------------------------------------------------------------------
def _standardise(x: np.ndarray) -> np.ndarray:
    """Centre each column and scale to unit Euclidean norm (so Z'Z == cor)."""
    mu = x.mean(axis=0)
    xc = x - mu
    ss = np.einsum("ij,ij->j", xc, xc)
    with np.errstate(divide="ignore", invalid="ignore"):
        z = xc / np.sqrt(ss)
    return z


def column_pairwise_cor(x: np.ndarray, spearman: bool = False) -> np.ndarray:
    """Dense p x p Pearson (or Spearman = Pearson of average ranks) matrix.

    sd == 0 columns give NaN rows/cols (IEEE propagation; base R gives NA).
    """
    x = np.asarray(x, dtype=np.float64)
    if spearman:
        x = rank_matrix_col(x)
    z = _standardise(x)
    with np.errstate(invalid="ignore"):
        return z.T @ z


def column_pairwise_cov(x: np.ndarray) -> np.ndarray:
    """Sample covariance (n-1), == base R cov(); r_cors_similarity.rs:49-55."""
    x = np.asarray(x, dtype=np.float64)
    xc = x - x.mean(axis=0)
    return (xc.T @ xc) / (x.shape[0] - 1)


def column_pairwise_cos(x: np.ndarray) -> np.ndarray:
    """Column cosine similarity (rs_cos, r_cors_similarity.rs:90-97)."""
    x = np.asarray(x, dtype=np.float64)
    xn = x / np.sqrt((x * x).sum(axis=0))
    return xn.T @ xn


def cor_two_matrices(x: np.ndarray, y: np.ndarray, spearman: bool = False) -> np.ndarray:
    """p x q correlation between the columns of x and of y (rs_cor2, r_cors_similarity.rs:114-122;
    == base R cor(x, y), inst/tinytest/test_R_rust_equality.R:39-55)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if x.shape[0] != y.shape[0]:
        raise ValueError("x and y need the same number of rows")
    if spearman:
        x, y = rank_matrix_col(x), rank_matrix_col(y)
    return _standardise(x).T @ _standardise(y)


def cov2cor(c: np.ndarray) -> np.ndarray:
    """rs_cov2cor (r_cors_similarity.rs:135-142; test_R_rust_equality.R:76-82)."""
    c = np.asarray(c, dtype=np.float64)
    d = np.sqrt(np.diag(c))
    return c / np.outer(d, d)


# --------------------------------------------------------------------------
# packed upper triangle (authoritative in-tree loops:
# src/rust/src/base/r_helpers.rs:75-94 and :113-134; gather order of
# r_cors_similarity.rs:251-268)
# --------------------------------------------------------------------------
def upper_triangle_indices(p: int, shift: int) -> tuple[np.ndarray, np.ndarray]:
    """Row-major (i outer, j from i+shift) index vectors, as the crate's
    utils::matrix_utils::upper_triangle_indices (call site r_cors_similarity.rs:257)."""
    rows, cols = np.triu_indices(p, k=int(shift))
    return rows.astype(np.int64), cols.astype(np.int64)


def packed_offset(i, p: int, shift: int):
    """Offset of the first packed element of row i (0-based)."""
    i = np.asarray(i, dtype=np.int64)
    if shift:
        return i * (2 * p - i - 1) // 2
    return i * (2 * p - i + 1) // 2


def dense_to_upper_triangle(m: np.ndarray, shift: bool = True) -> np.ndarray:
    r, c = upper_triangle_indices(m.shape[0], 1 if shift else 0)
    return np.ascontiguousarray(np.asarray(m)[r, c])


def upper_triangle_to_dense(v: np.ndarray, shift: bool, p: int) -> np.ndarray:
    """r_helpers.rs:75-94: mirror, and diagonal := 1 when shift is set."""
    v = np.asarray(v, dtype=np.float64)
    m = np.zeros((p, p), dtype=np.float64)
    r, c = upper_triangle_indices(p, 1 if shift else 0)
    assert v.shape[0] == r.shape[0], "packed length mismatch"
    m[r, c] = v
    m[c, r] = v
    if shift:
        m[np.arange(p), np.arange(p)] = 1.0
    return m


def cor_upper_triangle(x: np.ndarray, spearman: bool, shift: bool = True) -> np.ndarray:
    """rs_cor_upper_triangle (r_cors_similarity.rs:251-268)."""
    return dense_to_upper_triangle(column_pairwise_cor(x, spearman), shift)


def soft_threshold(r: np.ndarray, beta: float, signed: bool = False) -> np.ndarray:
    """North-star extension (NOT in the reference): |r|^beta, or sign(r)|r|^beta.
    beta == 1 is the identity on r itself (bit-identical, sign kept)."""
    if beta == 1.0:
        return np.array(r, dtype=np.float64, copy=True)
    a = np.abs(r) ** beta
    return np.copysign(a, r) if signed else a


# --------------------------------------------------------------------------
# differential correlation (crate methods::diffcor::calculate_diff_correlation,
# call site src/rust/src/methods/r_diffcor.rs:45-74).  PARITY UNPINNED.
# --------------------------------------------------------------------------
def calculate_diff_correlation(r_a, r_b, n_a: int, n_b: int, spearman: bool, fieller: float | None = None):
    """Fisher-z differential correlation on packed (shift = 1) vectors.

    z = (atanh r_a - atanh r_b) / sqrt(c/(n_a-3) + c/(n_b-3)), c = 1.06 for
    Spearman (Fieller), 1 for Pearson; two-sided normal p = erfc(|z|/sqrt 2).
    r = +-1 gives atanh = +-inf (rows dropped downstream,
    R/classes_triangular_mat.R:298-300).
    """
    r_a = np.asarray(r_a, dtype=np.float64)
    r_b = np.asarray(r_b, dtype=np.float64)
    c = (1.06 if spearman else 1.0) if fieller is None else float(fieller)
    se = math.sqrt(c / (n_a - 3) + c / (n_b - 3))
    with np.errstate(divide="ignore", invalid="ignore"):
        z = (np.arctanh(r_a) - np.arctanh(r_b)) / se
    from scipy.special import erfc

    p = erfc(np.abs(z) / math.sqrt(2.0))
    return {"r_a": r_a, "r_b": r_b, "z_score": z, "p_val": p}


def differential_cor(x_a, x_b, spearman: bool, fieller: float | None = None):
    """rs_differential_cor (r_diffcor.rs:45-74)."""
    x_a = np.asarray(x_a, dtype=np.float64)
    x_b = np.asarray(x_b, dtype=np.float64)
    if x_a.shape[1] != x_b.shape[1]:
        raise ValueError("Input matrices must have the same number of columns.")
    ra = cor_upper_triangle(x_a, spearman, True)
    rb = cor_upper_triangle(x_b, spearman, True)
    return calculate_diff_correlation(ra, rb, x_a.shape[0], x_b.shape[0], spearman, fieller)


# --------------------------------------------------------------------------
# TOM (crate methods::coremo::calc_tom, call site
# src/rust/src/methods/r_coremo.rs:46-59; formulas R/functions_stats.R:15-37;
# the include-diagonal-in-k / exclude-u-in-{i,j} reading is the one that
# reproduces inst/tinytest/test_cor_modules.R:9-46)
# --------------------------------------------------------------------------
def calc_tom(a: np.ndarray, signed: bool, version: str = "v1") -> np.ndarray:
    a = np.asarray(a, dtype=np.float64)
    if version not in ("v1", "v2"):
        raise ValueError(f"Invalid TOM version type: {version}")
    # rs_tom level: the matrix is used AS PASSED.  Signed: |a_ij| in k and in the denominators; unsigned: a_ij itself
    # (R/functions_stats.R:15-37).  abs() for unsigned networks is applied by the R wrappers calculate_tom /
    # calculate_tom_from_exp (R/functions_stats.R:52-54,114-116) - `calculate_tom` below - and NOT by cor_module_tom
    # (R/methods_coexp_cor.R:151).  Unsigned with negative entries is unpinned (every reference KAT passes |A|).
    absa = np.abs(a) if signed else a
    k = absa.sum(axis=1)  # includes the diagonal entry as passed
    d = np.diag(a)
    shared = a @ a - a * (d[:, None] + d[None, :])  # sum over u not in {i, j}
    mk = np.minimum(k[:, None], k[None, :])
    if version == "v1":
        t = (a + shared) / (mk + 1.0 - absa)
    else:
        t = 0.5 * (a + shared / (mk + absa))
    t[np.arange(a.shape[0]), np.arange(a.shape[0])] = 1.0  # unpinned; WGCNA convention
    return t


def calculate_tom(cor_mat: np.ndarray, signed: bool, version: str = "v1") -> np.ndarray:
    """R/functions_stats.R:43-59: abs() first when unsigned, then rs_tom."""
    cor_mat = np.asarray(cor_mat, dtype=np.float64)
    return calc_tom(cor_mat if signed else np.abs(cor_mat), signed, version)


def tom_from_exp(x, signed: bool, version: str = "v1", spearman: bool = False, beta: float = 1.0):
    """calculate_tom_from_exp (R/functions_stats.R:103-121) plus the north-star
    soft-threshold extension (beta == 1 reproduces the reference)."""
    r = column_pairwise_cor(x, spearman)
    a = soft_threshold(r, beta, signed=signed) if beta != 1.0 else r
    if not signed:
        a = np.abs(a)
    np.fill_diagonal(a, 1.0)
    return calc_tom(a, signed, version)


# --------------------------------------------------------------------------
# sparse single-cell variant (extension; semantic anchor rs_pairwise_gene_cors_mc,
# src/rust/src/meta_cell/r_mc_processing.rs:258-283: values cast to f32, Pearson
# over ALL cells including the implicit zeros).  PARITY UNPINNED.
# --------------------------------------------------------------------------
def _csr_to_dense(indptr, indices, data, cells: int, genes: int) -> np.ndarray:
    dense = np.zeros((cells, genes), dtype=np.float64)
    indptr = np.asarray(indptr, dtype=np.int64)
    rows = np.repeat(np.arange(cells, dtype=np.int64), np.diff(indptr))
    np.add.at(dense, (rows, np.asarray(indices, dtype=np.int64)), np.asarray(data, dtype=np.float32).astype(np.float64))
    return dense


def sparse_gram_cor(indptr, indices, data, cells: int, genes: int, chunk: int = 65536) -> np.ndarray:
    """All-pairs Pearson from a cells x genes CSR matrix, packed (shift = 1).

    r_ij = (G_ij - s_i s_j / N) / sqrt((G_ii - s_i^2/N)(G_jj - s_j^2/N)),
    G = X'X, s = column sums, accumulated in float64 over cell chunks.
    """
    indptr = np.asarray(indptr, dtype=np.int64)
    g = np.zeros((genes, genes), dtype=np.float64)
    s = np.zeros(genes, dtype=np.float64)
    for c0 in range(0, cells, chunk):
        c1 = min(cells, c0 + chunk)
        lo, hi = indptr[c0], indptr[c1]
        blk = _csr_to_dense(indptr[c0 : c1 + 1] - lo, indices[lo:hi], data[lo:hi], c1 - c0, genes)
        g += blk.T @ blk
        s += blk.sum(axis=0)
    cov = g - np.outer(s, s) / cells
    var = np.diag(cov).copy()
    with np.errstate(divide="ignore", invalid="ignore"):
        r = cov / np.sqrt(np.outer(var, var))
    return dense_to_upper_triangle(r, True)


def pairwise_gene_cors(indptr, indices, data, cells: int, genes: int, idx1, idx2) -> np.ndarray:
    """Pair-list Pearson (rs_pairwise_gene_cors_mc semantics) == the all-pairs
    result restricted to the listed pairs."""
    dense = _csr_to_dense(indptr, indices, data, cells, genes)
    out = np.empty(len(idx1), dtype=np.float64)
    for t, (i, j) in enumerate(zip(idx1, idx2)):
        out[t] = np.corrcoef(dense[:, i], dense[:, j])[0, 1]
    return out


def pairwise_gene_cors_spearman(indptr, indices, data, cells: int, genes: int, idx1, idx2) -> np.ndarray:
    """Spearman flavour of the pair list: Pearson of the average-tie ranks over ALL cells (zeros tie)."""
    dense = _csr_to_dense(indptr, indices, data, cells, genes)
    out = np.empty(len(idx1), dtype=np.float64)
    for t, (i, j) in enumerate(zip(idx1, idx2)):
        out[t] = np.corrcoef(rank_average(dense[:, i]), rank_average(dense[:, j]))[0, 1]
    return out


# --------------------------------------------------------------------------
# RBF affinities (crate core::math::rbf behind src/rust/src/base/r_rbf.rs:36-131) and the
# leave-one-out CoReMo stability distances (src/rust/src/methods/r_coremo.rs:191-235).
# The kernel formulas are not in the reference tree; they are the textbook definitions the
# crate documents (gaussian exp(-(eps d)^2), inverse quadratic 1/(1+(eps d)^2), bump
# exp(-1/(1-(eps d)^2) + 1) inside d < 1/eps, 0 outside).  PARITY UNPINNED: no reference test
# holds values of these functions (they only feed hclust / Leiden fixtures).
# --------------------------------------------------------------------------
def rbf_function(d, epsilon: float, rbf_type: str) -> np.ndarray:
    d = np.asarray(d, dtype=np.float64)
    t2 = (epsilon * d) ** 2
    if rbf_type == "gaussian":
        return np.exp(-t2)
    if rbf_type == "inverse_quadratic":
        return 1.0 / (1.0 + t2)
    if rbf_type == "bump":
        out = np.zeros_like(d)
        inside = d < 1.0 / epsilon
        out[inside] = np.exp(-1.0 / (1.0 - t2[inside]) + 1.0)
        return out
    raise ValueError(f"Invalid RBF function: {rbf_type}")


def rbf_iterate_epsilons(dist, epsilons, original_dim: int, shift: bool, rbf_type: str) -> np.ndarray:
    """original_dim x len(epsilons): column sums of the symmetric affinity matrix rebuilt from the packed
    distances (diagonal := 1 when shift, as rs_upper_triangle_to_dense, r_helpers.rs:83-84)."""
    out = np.empty((original_dim, len(epsilons)), dtype=np.float64)
    for e, eps in enumerate(epsilons):
        aff = upper_triangle_to_dense(rbf_function(dist, eps, rbf_type), shift, original_dim)
        out[:, e] = aff.sum(axis=0)
    return out


def coremo_stability(x, indices_1based, epsilon: float, rbf_type: str, spearman: bool) -> list[np.ndarray]:
    """r_coremo.rs:204-226: drop one sample, correlate, d = 1 - |r| (packed, shift 1), return 1 - phi(d)."""
    x = np.asarray(x, dtype=np.float64)
    res = []
    for idx in indices_1based:
        red = np.delete(x, int(idx) - 1, axis=0)
        d = 1.0 - np.abs(cor_upper_triangle(red, spearman, True))
        res.append(1.0 - rbf_function(d, epsilon, rbf_type))
    return res


# --------------------------------------------------------------------------
# multi-GPU partition of the packed triangle (host logic; mirrored by
# bixcor_shard_rows in include/bixcor.h)
# --------------------------------------------------------------------------
# --------------------------------------------------------------------------
# reciprocal best hits on correlations (rs_rbh_cor,
# src/rust/src/methods/r_rbh.rs:187-272; crate methods::rbh::calculate_rbh_cor)
# Pinned by the known-answer test inst/tinytest/test_rbh.R:109-214 (k_best = 1,
# Pearson 1e-7 and Spearman) on R's own set.seed(42) matrices.  Unpinned: the
# use of |r| as the similarity (doc ":172-173": "Minimum (absolute)
# correlations"; every KAT hit is positive), tie handling of the k-th best.
# --------------------------------------------------------------------------
def rbh_cor(x: np.ndarray, y: np.ndarray, k_best: int, spearman: bool, min_similarity: float = 0.0):
    """x (features x p), y (same features x q) -> list of (i, j, |r_ij|), origin-major, for pairs that are among the
    k_best largest of row i and of column j and exceed min_similarity."""
    sim = np.abs(cor_two_matrices(np.asarray(x, float), np.asarray(y, float), spearman))
    p, q = sim.shape

    def kth(v):
        v = np.sort(v[~np.isnan(v)])[::-1]
        return v[k_best - 1] if v.size >= k_best else -np.inf

    row_thr = np.array([kth(sim[i, :]) for i in range(p)])
    col_thr = np.array([kth(sim[:, j]) for j in range(q)])
    hits = []
    for i in range(p):
        for j in range(q):
            v = sim[i, j]
            if v >= row_thr[i] and v >= col_thr[j] and v > min_similarity:
                hits.append((i, j, float(v)))
    return hits


def rbh_cor_list(module_matrices: dict, k_best: int, spearman: bool, min_similarity: float) -> dict:
    """The list-level driver of r_rbh.rs:195-270: every origin against every later target, features intersected by
    row name; module_matrices maps name -> (matrix, rownames, colnames)."""
    names = list(module_matrices)
    res = {k: [] for k in ("origin", "target", "comparisons", "origin_modules", "target_modules", "similarity")}
    for a, na in enumerate(names[:-1]):
        xa, ra, ca = module_matrices[na]
        for nb in names[a + 1:]:
            xb, rb, cb = module_matrices[nb]
            pos_b = {f: k for k, f in enumerate(rb)}
            shared = [(k, pos_b[f]) for k, f in enumerate(ra) if f in pos_b]
            hits = []
            if len(shared) >= 2:
                ia, ib = zip(*shared)
                hits = rbh_cor(np.asarray(xa, float)[list(ia)], np.asarray(xb, float)[list(ib)], k_best, spearman, min_similarity)
            res["origin"].append(na)
            res["target"].append(nb)
            res["comparisons"].append(len(hits))
            res["origin_modules"] += [ca[i] for i, _, _ in hits]
            res["target_modules"] += [cb[j] for _, j, _ in hits]
            res["similarity"] += [v for _, _, v in hits]
    return res


def shard_rows(p: int, shift: int, world: int, rank: int, align: int = 128) -> tuple[int, int]:
    """Contiguous block-row range [r0, r1) owned by `rank`: boundaries at
    equal-pair-count quantiles of the triangle, rounded to `align` rows."""
    total = p * (p - 1) // 2 if shift else p * (p + 1) // 2

    def boundary(q: int) -> int:
        if q <= 0:
            return 0
        if q >= world:
            return p
        target = total * q / world
        # smallest r with pairs_before(r) >= target, pairs_before(r) = packed_offset(r)
        a = 2 * p - 1 if shift else 2 * p + 1
        disc = a * a - 8.0 * target
        r = (a - math.sqrt(max(disc, 0.0))) / 2.0
        r = int(round(r / align)) * align
        return max(0, min(p, r))

    return boundary(rank), boundary(rank + 1)
