"""Pins the CPU oracle against the reference's golden vectors (CPU only)."""
import json
import os

import numpy as np
import pytest
from scipy.stats import rankdata, spearmanr

from oracle import oracle as o


@pytest.fixture(scope="module")
def kat(golden_dir):
    with open(os.path.join(golden_dir, "reference_kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def req(golden_dir):
    with open(os.path.join(golden_dir, "r_equality_seed123.json")) as f:
        return {k: (np.array(v) if isinstance(v, list) else v) for k, v in json.load(f).items()}


# inst/tinytest/test_cor_modules.R:9-110 (tolerance 1e-5)
@pytest.mark.parametrize("signed,version", [(True, "v1"), (True, "v2"), (False, "v1"), (False, "v2")])
def test_tom_known_answer(kat, signed, version):
    cor = o.upper_triangle_to_dense(kat["tom_input_packed"], True, kat["tom_n"])
    a = cor if signed else np.abs(cor)  # calculate_tom takes abs() first when unsigned
    got = o.dense_to_upper_triangle(o.calc_tom(a, signed, version), True)
    exp = np.array(kat[f"tom_{'signed' if signed else 'unsigned'}_{version}"])
    assert np.abs(got - exp).max() < kat["tom_tolerance"]


def test_tom_invalid_version():
    with pytest.raises(ValueError, match="Invalid TOM version type"):
        o.calc_tom(np.eye(3), True, "v3")


# inst/tinytest/test_sparse.R:5-9 and r_helpers.rs:75-134
def test_triangle_dense_roundtrip(kat):
    v = np.array(kat["sparse_input_packed"], dtype=float)
    d = o.upper_triangle_to_dense(v, True, kat["sparse_n"])
    assert np.array_equal(d, np.array(kat["sparse_dense_expected"]))
    assert np.array_equal(o.dense_to_upper_triangle(d, True), v)
    # shift = 0 keeps the diagonal
    v0 = o.dense_to_upper_triangle(d, False)
    assert v0.shape[0] == 10 and np.array_equal(o.upper_triangle_to_dense(v0, False, 4), d)


def test_packed_offsets_match_loop_order():
    for p in (1, 2, 5, 17):
        for shift in (0, 1):
            r, c = o.upper_triangle_indices(p, shift)
            idx = 0
            for i in range(p):
                assert o.packed_offset(i, p, shift) == idx
                for j in range(i + shift, p):
                    assert (r[idx], c[idx]) == (i, j)
                    idx += 1
            assert idx == r.shape[0]


# inst/tinytest/test_R_rust_equality.R:5-72 on the reference's own set.seed(123) matrix
def test_pearson_spearman_cov_equal_base_r_standin(req):
    mat = req["mat"]
    assert mat.shape == (12, 14)
    assert np.abs(o.column_pairwise_cor(mat, False) - req["pearson"]).max() < 1e-14
    assert np.abs(o.column_pairwise_cor(mat, True) - req["spearman"]).max() < 1e-14
    assert np.abs(o.column_pairwise_cov(mat) - req["cov"]).max() < 1e-14
    assert np.array_equal(o.rank_matrix_col(mat), req["ranks"])
    # independent implementation
    assert np.abs(o.column_pairwise_cor(mat, True) - spearmanr(mat).statistic).max() < 1e-14
    # upper-triangle class round trip (test_R_rust_equality.R:60-72)
    packed = o.cor_upper_triangle(mat, False, True)
    assert np.abs(o.upper_triangle_to_dense(packed, True, 14) - req["pearson"]).max() < 1e-14


def test_two_matrix_cor_and_cov2cor_equal_base_r_standin(req):
    """inst/tinytest/test_R_rust_equality.R:37-55,76-82."""
    mat, mat2 = req["mat"], req["mat_2"]
    assert np.abs(o.cor_two_matrices(mat, mat2, False) - req["pearson_2"]).max() < 1e-14
    assert np.abs(o.cor_two_matrices(mat, mat2, True) - req["spearman_2"]).max() < 1e-14
    assert np.abs(o.cov2cor(o.column_pairwise_cov(mat)) - req["pearson"]).max() < 1e-14
    c = o.column_pairwise_cos(mat)
    assert np.allclose(np.diag(c), 1.0) and np.allclose(c, c.T)


def test_rank_average_ties_signed_zero_nan():
    rng = np.random.default_rng(7)
    x = np.round(rng.normal(size=(257, 11)), 1)
    x[3, 0], x[9, 0] = 0.0, -0.0
    assert np.array_equal(o.rank_matrix_col(x), np.apply_along_axis(rankdata, 0, x))
    x[5, 2] = np.nan
    r = o.rank_matrix_col(x)
    assert np.isnan(r[:, 2]).all() and not np.isnan(r[:, 1]).any()
    assert np.array_equal(o.rank_average(np.array([2.0, 2.0, 2.0])), np.array([2.0, 2.0, 2.0]))


def test_zero_variance_column_is_nan():
    x = np.random.default_rng(0).normal(size=(10, 4))
    x[:, 1] = 3.0
    r = o.column_pairwise_cor(x, False)
    assert np.isnan(r[1]).all() and np.isnan(r[:, 1]).all() and not np.isnan(r[0, 2])


def test_soft_threshold_identity_and_power():
    r = np.array([-0.5, 0.25, 1.0])
    assert np.array_equal(o.soft_threshold(r, 1.0), r)
    assert np.allclose(o.soft_threshold(r, 6.0), np.abs(r) ** 6)
    assert np.allclose(o.soft_threshold(r, 3.0, signed=True), np.sign(r) * np.abs(r) ** 3)


def test_diffcor_definition():
    rng = np.random.default_rng(3)
    xa, xb = rng.normal(size=(30, 6)), rng.normal(size=(40, 6))
    res = o.differential_cor(xa, xb, spearman=False)
    ra = o.cor_upper_triangle(xa, False, True)
    rb = o.cor_upper_triangle(xb, False, True)
    z = (np.arctanh(ra) - np.arctanh(rb)) / np.sqrt(1 / 27 + 1 / 37)
    assert np.allclose(res["z_score"], z, rtol=1e-14)
    from scipy.stats import norm

    assert np.allclose(res["p_val"], 2 * norm.sf(np.abs(z)), rtol=1e-12)
    sp = o.differential_cor(xa, xb, spearman=True)
    assert np.allclose(sp["z_score"] * np.sqrt(1.06), (np.arctanh(sp["r_a"]) - np.arctanh(sp["r_b"])) / np.sqrt(1 / 27 + 1 / 37))
    with pytest.raises(ValueError, match="same number of columns"):
        o.differential_cor(xa, xb[:, :5], False)


def test_tom_from_exp_matches_chain():
    x = np.random.default_rng(5).normal(size=(25, 9))
    cor = o.column_pairwise_cor(x, False)
    assert np.allclose(o.tom_from_exp(x, True, "v1"), o.calc_tom(cor, True, "v1"))
    assert np.allclose(o.tom_from_exp(x, False, "v2"), o.calc_tom(np.abs(cor), False, "v2"))


def test_sparse_gram_matches_dense_and_pairs():
    from bixverse_b200.synthetic import synthetic_sparse_cells

    cells, genes = 700, 40
    ip, ix, d = synthetic_sparse_cells(cells, genes, 11, density=0.2)
    packed = o.sparse_gram_cor(ip, ix, d, cells, genes, chunk=256)
    dense = o._csr_to_dense(ip, ix, d, cells, genes)
    ref = o.cor_upper_triangle(dense, False, True)
    ok = ~np.isnan(ref)
    assert np.abs(packed[ok] - ref[ok]).max() < 1e-12
    i, j = o.upper_triangle_indices(genes, 1)
    sel = np.arange(0, i.shape[0], 37)
    pl = o.pairwise_gene_cors(ip, ix, d, cells, genes, i[sel], j[sel])
    okp = ~np.isnan(pl)
    assert np.abs(pl[okp] - packed[sel][okp]).max() < 1e-12


def test_c_restatement_agrees_with_numpy_oracle(lib_built):
    from oracle import c_binding as cb

    rng = np.random.default_rng(21)
    x = np.round(rng.normal(size=(33, 12)), 1)
    assert np.array_equal(cb.rank_matrix_col(x), o.rank_matrix_col(x))
    for sp in (False, True):
        assert np.abs(cb.cor_dense(x, sp) - o.column_pairwise_cor(x, sp)).max() < 1e-14
        for shift in (False, True):
            assert np.abs(cb.cor_upper(x, sp, shift) - o.cor_upper_triangle(x, sp, shift)).max() < 1e-14
    a = o.column_pairwise_cor(x, False)
    for signed in (True, False):
        for ver in (1, 2):
            assert np.abs(cb.tom(a, signed, ver) - o.calc_tom(a, signed, f"v{ver}")).max() < 1e-13
    ra, rb = o.cor_upper_triangle(x[:20], False), o.cor_upper_triangle(x[13:], False)
    z, pv = cb.diffcor(ra, rb, 20, 20, 1.0)
    ref = o.calculate_diff_correlation(ra, rb, 20, 20, False)
    assert np.allclose(z, ref["z_score"], rtol=1e-13) and np.allclose(pv, ref["p_val"], rtol=1e-12)


def test_shard_rows_partition_properties():
    for p in (100, 2000, 20000, 20011):
        for shift in (0, 1):
            for world in (1, 2, 3, 4, 8):
                bounds = [o.shard_rows(p, shift, world, r) for r in range(world)]
                assert bounds[0][0] == 0 and bounds[-1][1] == p
                for a, b in zip(bounds[:-1], bounds[1:]):
                    assert a[1] == b[0] and a[0] % 128 == 0
                if p >= 20000:
                    pairs = [o.packed_offset(b, p, shift) - o.packed_offset(a, p, shift) for a, b in bounds]
                    assert max(pairs) / (sum(pairs) / world) < 1.06  # balanced by pair count


def test_rbf_and_stability_restatement():
    """Internal consistency of the RBF / leave-one-out restatement (parity unpinned, see oracle.py)."""
    d = np.linspace(0.0, 1.0, 11)
    assert o.rbf_function(d, 2.0, "gaussian")[0] == 1.0 and o.rbf_function(d, 2.0, "bump")[0] == 1.0
    assert o.rbf_function(d, 2.0, "inverse_quadratic")[5] == pytest.approx(1.0 / (1.0 + 1.0))
    assert np.all(o.rbf_function(d, 2.0, "bump")[d >= 0.5] == 0.0)
    rng = np.random.default_rng(1)
    x = rng.normal(size=(20, 9))
    dist = 1.0 - np.abs(o.cor_upper_triangle(x, False, True))
    k = o.rbf_iterate_epsilons(dist, [1.0, 3.0], 9, True, "gaussian")
    dense = o.upper_triangle_to_dense(o.rbf_function(dist, 3.0, "gaussian"), True, 9)
    assert np.allclose(k[:, 1], dense.sum(axis=0)) and np.allclose(np.diag(dense), 1.0)
    res = o.coremo_stability(x, [1, 20], 2.0, "gaussian", True)
    red = x[1:]
    ref = 1.0 - np.exp(-((2.0 * (1.0 - np.abs(o.cor_upper_triangle(red, True, True)))) ** 2))
    assert np.allclose(res[0], ref) and len(res) == 2 and res[1].shape == ref.shape


def _rbh_fixture(golden_dir):
    with open(os.path.join(golden_dir, "rbh_cor_seed42.json")) as f:
        g = json.load(f)
    mats = {
        "origin_1": (np.array(g["matrix_a"]), g["rownames_a"], g["colnames_a"]),
        "origin_2": (np.array(g["matrix_b"]), g["rownames_b"], g["colnames_b"]),
    }
    disjoint = {"origin_1": mats["origin_1"], "origin_2": (np.array(g["matrix_c"]), g["rownames_c"], g["colnames_c"])}
    return g, mats, disjoint


@pytest.mark.parametrize("spearman", [False, True])
def test_rbh_cor_known_answer(golden_dir, spearman):
    """inst/tinytest/test_rbh.R:109-214: k_best = 1, minimum_similarity = 0, on the test's own set.seed(42) matrices."""
    g, mats, disjoint = _rbh_fixture(golden_dir)
    res = o.rbh_cor_list(mats, 1, spearman, 0.0)
    assert res["origin"] == ["origin_1"] and res["target"] == ["origin_2"] and res["comparisons"] == [3]
    assert res["origin_modules"] == g["expected_origin_modules"]
    assert res["target_modules"] == g["expected_target_modules"]
    np.testing.assert_allclose(res["similarity"], g["expected_spearman" if spearman else "expected_pearson"], atol=g["tolerance"])
    none = o.rbh_cor_list(disjoint, 1, spearman, 0.0)  # no overlapping features (:196-214)
    assert none["comparisons"] == [0] and none["similarity"] == []


def test_rbh_cor_k_best_and_threshold():
    rng = np.random.default_rng(5)
    x, y = rng.normal(size=(40, 7)), rng.normal(size=(40, 9))
    y[:, 2] = -x[:, 4] + 0.01 * rng.normal(size=40)  # strong negative hit: similarity is |r|
    h1 = o.rbh_cor(x, y, 1, False, 0.0)
    assert (4, 2) in [(i, j) for i, j, _ in h1] and all(v > 0 for _, _, v in h1)
    h3 = o.rbh_cor(x, y, 3, False, 0.0)
    assert set((i, j) for i, j, _ in h1) <= set((i, j) for i, j, _ in h3) and len(h3) > len(h1)
    assert o.rbh_cor(x, y, 3, False, 0.9) == [h for h in h3 if h[2] > 0.9]


def test_c_restatement_next_rows(lib_built, golden_dir):
    """The plain-C oracle of the 'next' rows: pinned by the reference's rbh known answers, consistent with numpy."""
    from oracle import c_binding as cb

    g, mats, _ = _rbh_fixture(golden_dir)
    a, b = mats["origin_1"][0], mats["origin_2"][0][:4]
    for spearman, key in ((False, "expected_pearson"), (True, "expected_spearman")):
        hits = cb.rbh_cor(a, b, 1, spearman, 0.0)
        assert [(g["colnames_a"][i], g["colnames_b"][j]) for i, j, _ in hits] == list(zip(g["expected_origin_modules"], g["expected_target_modules"]))
        np.testing.assert_allclose([v for _, _, v in hits], g[key], atol=g["tolerance"])
    rng = np.random.default_rng(4)
    x, y = np.round(rng.normal(size=(25, 6)), 1), rng.normal(size=(25, 9))
    for sp in (False, True):
        assert np.abs(cb.cor2(x, y, sp) - o.cor_two_matrices(x, y, sp)).max() < 1e-14
        for k in (1, 3):
            hc, hn = cb.rbh_cor(x, y, k, sp, 0.05), o.rbh_cor(x, y, k, sp, 0.05)
            assert [(i, j) for i, j, _ in hc] == [(i, j) for i, j, _ in hn]
            assert np.allclose([v for _, _, v in hc], [v for _, _, v in hn], atol=1e-14)
    d = np.abs(rng.normal(size=200))
    for t in ("gaussian", "bump", "inverse_quadratic"):
        assert np.allclose(cb.rbf(d, 1.3, t), o.rbf_function(d, 1.3, t), rtol=1e-14, atol=0)


def test_golden_fixtures_regenerate_identically(golden_dir, tmp_path, monkeypatch):
    """The committed fixtures are exactly what tests/golden/make_golden.py writes (R's RNG restated, self-checked)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(golden_dir, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.setattr(mod, "HERE", str(tmp_path))
    mod.main()
    for name in ("reference_kat.json", "r_equality_seed123.json", "rbh_cor_seed42.json"):
        with open(os.path.join(golden_dir, name)) as f, open(tmp_path / name) as g:
            assert json.load(f) == json.load(g), name


def test_sampled_row_oracle_equals_the_full_oracle():
    """oracle/sampled.py (rows of the BASELINE-size results on demand) against the full-matrix oracle."""
    from oracle import sampled as sm

    rng = np.random.default_rng(5)
    x = np.round(rng.normal(size=(40, 300)), 1)  # ties
    xb = rng.normal(size=(31, 300))
    p = 300
    for sp in (False, True):
        so = sm.SampledOracle(x, sp)
        full = o.cor_upper_triangle(x, sp, True)
        full6 = o.soft_threshold(full, 6.0)
        for i in sm.pick_rows(0, p, p):
            assert np.abs(sm.packed_row(full, 0, i, p) - so.cor_upper_row(i)).max() < 1e-14
            assert np.abs(sm.packed_row(full6, 0, i, p) - so.cor_upper_row(i, 1, 6.0)).max() < 1e-14
        sb = sm.SampledOracle(xb, sp)
        d = o.differential_cor(x, xb, sp)
        for i in (0, 128, 298):
            ra, rb, z, pv = sm.diffcor_rows(so, sb, i)
            for got, key in ((ra, "r_a"), (rb, "r_b"), (z, "z_score"), (pv, "p_val")):
                assert np.allclose(got, sm.packed_row(d[key], 0, i, p), rtol=1e-12, atol=1e-14)
    so = sm.SampledOracle(x, False)
    for signed in (True, False):
        for ver in ("v1", "v2"):
            t = o.tom_from_exp(x, signed, ver, False, 6.0)
            for i in (0, 127, 250):
                cols = sm.tom_sample_cols(i, p)
                assert np.abs(so.tom_entries(i, cols, signed, ver, 6.0) - t[i, cols]).max() < 1e-13
    # shard-aware row picking: inside the range, both ends present, a shard that only holds the pair-less last row is empty
    assert sm.pick_rows(128, 256, 300)[0] == 128 and sm.pick_rows(128, 256, 300)[-1] == 255
    assert sm.pick_rows(299, 300, 300) == [] and sm.pick_rows(120, 120, 120) == []
    assert set(sm.pick_rows(0, 20000, 20000)) == set(sm.ROWS_20K)


def _packed_key_rank_model(col):
    """NumPy model of the register rank kernel's fast path (bixverse_b200/csrc/columns.cuh, rank_columns_warp_body): every
    element becomes one double - the value with -0 folded into +0, +-inf replaced by +-DBL_MAX and its low 10 mantissa bits
    replaced by the sample index; the replaced bits (+ an "infinite" flag) are parked by sample index.  After the sort,
    neighbours that agree above the index bits must have equal parked bits (a true tie), otherwise the kernel falls back to
    the exact (key, index) sort.  Returns (2 * average ranks or None, fallback?)."""
    n = col.size
    assert n <= 1024 and not np.isnan(col).any()
    b = (col + 0.0).view(np.uint64).copy()
    parked = (b & np.uint64(1023)).astype(np.uint32)
    isinf = (b & np.uint64(0x7FF0000000000000)) == np.uint64(0x7FF0000000000000)
    b[isinf] = (b[isinf] & np.uint64(0x8000000000000000)) | np.uint64(0x7FEFFFFFFFFFFFFF)
    parked[isinf] = 1024 | 1023
    packed = ((b & ~np.uint64(1023)) | np.arange(n, dtype=np.uint64)).view(np.float64)
    assert np.unique(packed).size == n  # all keys distinct: one comparison per exchange suffices
    order = np.argsort(packed, kind="stable")  # the sorting network orders by the packed DOUBLE values
    bits = packed.view(np.uint64)[order]
    same = np.zeros(n, dtype=bool)
    same[1:] = (bits[1:] ^ bits[:-1]) < np.uint64(1024)
    pk = parked[order]
    mixed = same.copy()
    mixed[1:] &= pk[1:] != pk[:-1]
    if mixed.any():
        return None, True
    start = ~same
    run_id = np.cumsum(start) - 1
    lo = np.flatnonzero(start)[run_id]
    ends = np.r_[np.flatnonzero(start)[1:] - 1, n - 1]
    hi = ends[run_id]
    rk2 = np.empty(n, dtype=np.int64)
    rk2[order] = lo + hi + 2
    return rk2, False


def test_packed_key_rank_model_is_exact_or_falls_back():
    """The correctness argument of the packed-key rank kernel, checked on the CPU: whenever the fast path accepts a column its
    ranks equal the reference's average-tie ranks exactly; columns it cannot decide (values closer than 1024 ulp that are not
    equal, finite values next to +-DBL_MAX beside infinities) are handed to the exact sort, never mis-ranked."""
    rng = np.random.default_rng(11)
    big, tiny = np.finfo(np.float64).max, np.finfo(np.float64).smallest_subnormal
    accepted = rejected = 0
    cols = []
    for n in (2, 31, 64, 333, 1000, 1024):
        cols += [rng.normal(size=n), np.round(rng.normal(size=n), 1), np.full(n, 2.5), rng.integers(-3, 4, size=n).astype(float),
                 np.where(rng.random(n) < 0.3, np.inf, rng.normal(size=n)), np.where(rng.random(n) < 0.3, -np.inf, np.round(rng.normal(size=n), 1)),
                 rng.integers(-3, 4, size=n) * tiny * 2048, np.where(rng.random(n) < 0.5, 0.0, -0.0)]
        near = 1.0 + rng.integers(0, 1024, size=n) * 2.0 ** -52            # distinct values inside one 1024-ulp bucket
        cols += [near, -near, np.r_[near[: n // 2], np.full(n - n // 2, 1.0)]]
        edge = rng.normal(size=n)
        edge[: min(n, 3)] = [np.inf, big, np.nextafter(big, 0)][: min(n, 3)]
        cols.append(edge)
        cols.append(rng.integers(-3, 4, size=n) * tiny)                      # subnormals: all inside the bucket around zero
    for col in cols:
        rk2, fallback = _packed_key_rank_model(col)
        want = 2.0 * o.rank_matrix_col(col[:, None])[:, 0]
        if fallback:
            rejected += 1
            # the fast path only rejects when two different values share the bits above the index field
            bb = (col + 0.0).view(np.uint64) >> np.uint64(10)
            inf_like = np.isinf(col) | (np.abs(col) >= np.nextafter(big, 0) * (1 - 2.0 ** -42))
            assert any(np.unique(col[bb == u]).size > 1 for u in np.unique(bb)) or (inf_like.any() and np.unique(col[inf_like]).size > 1)
        else:
            accepted += 1
            assert np.array_equal(rk2.astype(np.float64), want), col[:8]
    assert accepted > 40 and rejected > 10
