"""Randomised shapes through every precision path (ragged tile edges, odd block-row counts for the CTA-pair tile
lists, tie-heavy data, random row shards) against the oracle.  Seeded: the same 24 cases every run."""
import numpy as np
import pytest

from oracle import oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api(lib_built):
    from bixverse_b200 import api as a

    a.init(1, 0)
    return a


def _cases():
    rng = np.random.default_rng(20260101)
    out = []
    for _ in range(24):
        n = int(rng.choice([3, 17, 64, 129, 500, 1023, 1024, 1500]))
        p = int(rng.choice([2, 127, 128, 129, 255, 383, 640, 1111]))
        out.append((n, p, bool(rng.integers(0, 2)), bool(rng.integers(0, 2)), int(rng.integers(0, 1 << 30))))
    return out


@pytest.mark.parametrize("n,p,spearman,ties,seed", _cases())
def test_random_shapes_all_precisions(api, n, p, spearman, ties, seed):
    from bixverse_b200 import _lib

    rng = np.random.default_rng(seed)
    x = rng.normal(size=(n, p))
    if ties:
        x = np.round(x, 1)
    ref_dense = o.column_pairwise_cor(x, spearman)
    ok_dense = ~np.isnan(ref_dense)
    precs = [(_lib.PREC_F64, 1e-10), (_lib.PREC_TF32X3, 1e-5), (_lib.PREC_F16X3, 1e-5), (_lib.PREC_OZAKI, 1e-10)]
    if spearman and n <= 1860:
        precs.append((_lib.PREC_I8EXACT, 1e-12))
    nb = (p + 127) // 128
    for prec, tol in precs:
        for shift in (True, False):
            ref = o.dense_to_upper_triangle(ref_dense, shift)
            got = api.rs_cor_upper_triangle(x, spearman, shift, precision=prec)
            ok = ~np.isnan(ref)
            assert np.array_equal(np.isnan(got), ~ok)
            if ok.any():
                assert np.abs(got[ok] - ref[ok]).max() < tol, (prec, shift)
        if nb >= 2:  # a random tile-aligned row shard equals the matching slice of the full vector
            b0 = int(rng.integers(0, nb - 1))
            b1 = int(rng.integers(b0 + 1, nb + 1))
            r0, r1 = 128 * b0, min(p, 128 * b1)
            full = api.rs_cor_upper_triangle(x, spearman, True, beta=6.0, precision=prec)
            part = api.rs_cor_upper_triangle(x, spearman, True, beta=6.0, precision=prec, rows=(r0, r1))
            lo, hi = int(o.packed_offset(r0, p, 1)), int(o.packed_offset(r1, p, 1))
            assert np.array_equal(part, full[lo:hi], equal_nan=True), (prec, r0, r1)
        dense = api.rs_cor(x, spearman, precision=prec)
        rd = ref_dense.copy()
        np.fill_diagonal(rd, 1.0)
        m = ok_dense & ~np.eye(p, dtype=bool)
        if m.any():
            assert np.abs(dense[m] - rd[m]).max() < tol, prec
