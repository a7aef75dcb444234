"""Sampled-row oracle for BASELINE-size parity checks.  TEST INFRASTRUCTURE ONLY (tests/, bench.py's parity leg,
__graft_entry__.smoke()) - never imported by the product package.

At 20 000 genes the full oracle result (1.6 - 6.4 GB, 8e11 - 1.6e13 flop on the host) is too slow for a test budget,
but any single row of it is cheap: with Z the standardised (ranked) columns, row i of the correlation matrix is the
GEMV Z[:, i]' Z, and a TOM entry (i, j) only needs rows i and j of the affinity matrix.  The arithmetic is the one of
oracle/oracle.py (same helper functions), restricted to the sampled rows:

    correlation   column_pairwise_cor      src/rust/src/base/r_cors_similarity.rs:70-77, 251-268
    Fisher z / p  calculate_diff_correlation  src/rust/src/methods/r_diffcor.rs:45-74   (PARITY UNPINNED, see oracle.py)
    TOM           calc_tom                 src/rust/src/methods/r_coremo.rs:46-59, R/functions_stats.R:15-37
"""
from __future__ import annotations

import numpy as np

from . import oracle as o

# rows the full-size tests and bench.py sample at p = 20000: both ends, a tile edge (127 | 128), interior rows, the last
# ragged block (19871 is in the last 128-row block of 20000 = 156 * 128 + 32) and the last row that owns a pair
ROWS_20K = (0, 1, 127, 128, 5000, 12345, 19871, 19998)


def pick_rows(r0: int, r1: int, p: int, want: int = 8, preferred=ROWS_20K) -> list[int]:
    """Up to `want` rows of [r0, min(r1, p - 1)): the preferred ones that fall inside, topped up with evenly spaced rows
    (always the first and the last row of the range, where a shard boundary bug would show)."""
    hi = min(r1, p - 1)  # row p - 1 owns no pair
    if hi <= r0:
        return []
    rows = {r for r in preferred if r0 <= r < hi}
    for t in np.linspace(r0, hi - 1, num=want):
        if len(rows) >= want:
            break
        rows.add(int(round(t)))
    rows.update((r0, hi - 1))
    return sorted(rows)[: max(want, 2)]


def packed_row(vec_slice, base_offset: int, i: int, p: int, shift: int = 1):
    """Row i of a packed (shift) vector held as a slice that starts at packed offset `base_offset`."""
    s = 1 if shift else 0
    a = int(o.packed_offset(i, p, s)) - base_offset
    return vec_slice[a : a + (p - i - s)]


class SampledOracle:
    """Standardised columns of one expression matrix (samples x genes) on the host, rows of the results on demand."""

    def __init__(self, x: np.ndarray, spearman: bool, ranks: np.ndarray | None = None):
        x = np.asarray(x, dtype=np.float64)
        self.n, self.p = x.shape
        self.spearman = bool(spearman)
        v = (ranks if ranks is not None else o.rank_matrix_col(x)) if spearman else x
        self.z = np.asfortranarray(o._standardise(v))  # n x p, unit-norm centred columns (oracle.py: cor = Z'Z)

    def cor_row_full(self, i: int) -> np.ndarray:
        """Row i of column_pairwise_cor(x, spearman): all p entries."""
        return self.z[:, i] @ self.z

    def cor_upper_row(self, i: int, shift: int = 1, beta: float = 1.0) -> np.ndarray:
        """Row i of rs_cor_upper_triangle (entries j >= i + shift) with the soft-threshold extension."""
        r = self.cor_row_full(i)[i + (1 if shift else 0) :]
        return o.soft_threshold(r, beta) if beta != 1.0 else r

    def affinity_row(self, i: int, signed: bool, beta: float) -> np.ndarray:
        """Row i of the affinity matrix calculate_tom_from_exp builds (oracle.tom_from_exp): soft-threshold power with the
        sign kept for signed networks, abs() for unsigned ones, unit diagonal."""
        r = self.cor_row_full(i)
        a = o.soft_threshold(r, beta, signed=signed) if beta != 1.0 else (r if signed else np.abs(r))
        a[i] = 1.0
        return a

    def tom_entries(self, i: int, cols, signed: bool, version: str = "v1", beta: float = 1.0) -> np.ndarray:
        """TOM(i, j) for the listed columns j != i: the formulas of oracle.calc_tom on rows i and j of the affinity."""
        ai = self.affinity_row(i, signed, beta)
        ki = (np.abs(ai) if signed else ai).sum()
        out = np.empty(len(cols))
        for t, j in enumerate(cols):
            aj = self.affinity_row(int(j), signed, beta)
            kj = (np.abs(aj) if signed else aj).sum()
            aij = ai[j]
            shared = ai @ aj - aij * (ai[i] + aj[j])  # sum over u not in {i, j}
            den = abs(aij) if signed else aij
            mk = min(ki, kj)
            out[t] = (aij + shared) / (mk + 1.0 - den) if version == "v1" else 0.5 * (aij + shared / (mk + den))
        return out


def tom_sample_cols(i: int, p: int, want: int = 12) -> np.ndarray:
    """Columns j > i checked for row i of the packed TOM: the neighbours, a tile edge, the far end, evenly spaced ones."""
    if i + 1 >= p:
        return np.zeros(0, dtype=np.int64)
    cand = {i + 1, min(p - 1, i + 2), min(p - 1, (i // 128 + 1) * 128), min(p - 1, (i // 128 + 1) * 128 - 1), p - 1, p - 2}
    for t in np.linspace(i + 1, p - 1, num=want):
        cand.add(int(round(t)))
    return np.array(sorted(c for c in cand if i < c < p)[: want + 6], dtype=np.int64)


def diffcor_rows(oa: SampledOracle, ob: SampledOracle, i: int):
    """Row i of rs_differential_cor: (r_a, r_b, z, p) for j > i."""
    d = o.calculate_diff_correlation(oa.cor_upper_row(i), ob.cor_upper_row(i), oa.n, ob.n, oa.spearman)
    return d["r_a"], d["r_b"], d["z_score"], d["p_val"]
