"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Run with -m gpu on a B200.

Bars: ranks and packed indexing bit-exact; correlations / TOM / z within 1e-10 on the FP64 path
(BIXCOR_PREC_F64) and within 1e-5 on the TF32x3 path; the int8 Spearman path is held to 1e-12.
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle as o

pytestmark = pytest.mark.gpu

TOL64 = 1e-10
TOL32 = 1e-5  # north-star bar for the FP32 path


@pytest.fixture(scope="module")
def api(lib_built):
    import torch

    assert torch.cuda.is_available(), "GPU tests need a GPU"
    from bixverse_b200 import api as _api

    _api.init(1, 0)
    return _api


def _maxerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
    m = ~np.isnan(a)
    return float(np.abs(a[m] - b[m]).max()) if m.any() else 0.0


# ----------------------------------------------------------------------------- golden vectors
def test_r_equality_golden(api, golden_dir):
    """inst/tinytest/test_R_rust_equality.R:5-72 on the reference's own set.seed(123) matrix."""
    g = json.load(open(os.path.join(golden_dir, "r_equality_seed123.json")))
    mat = np.array(g["mat"])
    assert _maxerr(api.rs_cor(mat, False), np.array(g["pearson"])) < 1e-14 * 100
    assert _maxerr(api.rs_cor(mat, True), np.array(g["spearman"])) < 1e-14 * 100
    assert np.array_equal(api.rs_rank_matrix_col(mat), np.array(g["ranks"]))
    packed = api.rs_cor_upper_triangle(mat, False, True)
    cls = api.UpperTriangularSymMat(packed, [f"feature_{i}" for i in range(14)], True)
    assert _maxerr(cls.get_sym_matrix(), np.array(g["pearson"])) < 1e-12


@pytest.mark.parametrize("signed,version", [(True, "v1"), (True, "v2"), (False, "v1"), (False, "v2")])
def test_tom_known_answer(api, golden_dir, signed, version):
    """inst/tinytest/test_cor_modules.R:9-110."""
    kat = json.load(open(os.path.join(golden_dir, "reference_kat.json")))
    cor = api.rs_upper_triangle_to_dense(kat["tom_input_packed"], True, 4)
    tom = api.calculate_tom(cor, signed, version)
    got = api.rs_dense_to_upper_triangle(tom, True)
    exp = np.array(kat[f"tom_{'signed' if signed else 'unsigned'}_{version}"])
    assert np.abs(got - exp).max() < kat["tom_tolerance"]
    assert _maxerr(tom, o.calculate_tom(cor, signed, version)) < 1e-13


# ----------------------------------------------------------------------------- ranks (bit-exact)
@pytest.mark.parametrize("n,p", [(2, 3), (31, 5), (32, 4), (33, 7), (200, 64), (1000, 50), (1025, 9), (4097, 3)])
def test_ranks_bit_exact(api, n, p):
    rng = np.random.default_rng(n * 131 + p)
    x = rng.normal(size=(n, p))
    assert np.array_equal(api.rs_rank_matrix_col(x), o.rank_matrix_col(x))
    xt = np.round(x, 1)  # tie heavy
    xt[0, 0], xt[n - 1, 0] = 0.0, -0.0
    assert np.array_equal(api.rs_rank_matrix_col(xt), o.rank_matrix_col(xt))
    xc = xt.copy()
    xc[:, p - 1] = 2.5  # one all-ties column
    if p > 1:
        xc[n // 2, 0] = np.nan  # one NaN column
    assert np.array_equal(api.rs_rank_matrix_col(xc), o.rank_matrix_col(xc), equal_nan=True)
    xi = x.copy()
    xi[0, 0], xi[1, 0] = np.inf, -np.inf
    assert np.array_equal(api.rs_rank_matrix_col(xi), o.rank_matrix_col(xi))


# ----------------------------------------------------------------------------- correlations
@pytest.mark.parametrize("n,p", [(12, 14), (200, 2000), (57, 129), (16, 128), (17, 300), (1000, 257)])
@pytest.mark.parametrize("spearman", [False, True])
def test_cor_upper_and_dense(api, n, p, spearman):
    from bixverse_b200.synthetic import synthetic_bulk

    x = synthetic_bulk(n, p, seed=n + p) if n >= 50 else np.random.default_rng(n + p).normal(size=(n, p))
    dense_ref = o.column_pairwise_cor(x, spearman)
    for shift in (True, False):
        got = api.rs_cor_upper_triangle(x, spearman, shift)
        assert _maxerr(got, o.dense_to_upper_triangle(dense_ref, shift)) < TOL64
    dense = api.rs_cor(x, spearman)
    ref = dense_ref.copy()
    np.fill_diagonal(ref, 1.0)
    assert _maxerr(dense, ref) < TOL64
    assert np.array_equal(dense, dense.T, equal_nan=True)


def test_config1_pearson_packed(api):
    """BASELINE config 1: Pearson 2000 genes x 200 samples, packed shift=1 (P = 1 999 000)."""
    from bixverse_b200.synthetic import synthetic_bulk

    x = synthetic_bulk(200, 2000, seed=123)
    got = api.rs_cor_upper_triangle(x, False, True)
    assert got.shape[0] == 1_999_000
    assert _maxerr(got, o.cor_upper_triangle(x, False, True)) < TOL64


def test_soft_threshold_beta(api):
    from bixverse_b200.synthetic import synthetic_bulk

    x = synthetic_bulk(100, 300, seed=9, tie_heavy=True)
    r = o.cor_upper_triangle(x, True, True)
    base = api.rs_cor_upper_triangle(x, True, True, beta=1.0)
    assert _maxerr(base, r) < TOL64
    assert np.array_equal(api.rs_cor_upper_triangle(x, True, True), base)  # beta == 1 is a bit-identical no-op
    assert _maxerr(api.rs_cor_upper_triangle(x, True, True, beta=6.0), o.soft_threshold(r, 6.0)) < TOL64


def test_zero_variance_and_nan_columns(api):
    x = np.random.default_rng(4).normal(size=(40, 140))
    x[:, 3] = 1.0
    x[7, 130] = np.nan
    for spearman in (False, True):
        got = api.rs_cor_upper_triangle(x, spearman, True)
        ref = o.cor_upper_triangle(x, spearman, True)
        assert _maxerr(got, ref) < TOL64
        assert np.isnan(got).sum() == np.isnan(ref).sum() > 0


def test_row_sharded_calls_tile_the_packed_vector(api):
    from bixverse_b200.synthetic import synthetic_bulk

    n, p = 64, 700
    x = synthetic_bulk(n, p, seed=77)
    full = api.rs_cor_upper_triangle(x, True, True, beta=2.0)
    for world in (2, 3):
        parts = [api.rs_cor_upper_triangle(x, True, True, beta=2.0, rows=api.shard_rows(p, True, world, r)) for r in range(world)]
        assert np.array_equal(np.concatenate(parts), full)


def test_pageable_and_pinned_hosts_agree(api):
    import torch

    n, p = 100, 900
    x = np.random.default_rng(1).normal(size=(n, p))
    a = api.rs_cor_upper_triangle(x, False, True)
    xp = torch.from_numpy(np.ascontiguousarray(x.T)).pin_memory()  # p x n row-major == n x p column-major
    from bixverse_b200 import _lib

    lib = _lib.load()
    out = torch.empty(p * (p - 1) // 2, dtype=torch.float64).pin_memory()
    _lib.check(lib.bixcor_cor_upper(xp.data_ptr(), n, p, 0, 1, 1.0, 0, out.data_ptr()))
    assert np.array_equal(out.numpy(), a)


# ----------------------------------------------------------------------------- differential correlation
@pytest.mark.parametrize("spearman", [False, True])
def test_diffcor(api, spearman):
    from bixverse_b200.synthetic import synthetic_pair

    xa, xb = synthetic_pair(60, 75, 333, seed=2025)
    got = api.rs_differential_cor(xa, xb, spearman)
    ref = o.differential_cor(xa, xb, spearman)
    assert _maxerr(got["r_a"], ref["r_a"]) < TOL64 and _maxerr(got["r_b"], ref["r_b"]) < TOL64
    assert _maxerr(got["z_score"], ref["z_score"]) < 1e-9
    assert np.allclose(got["p_val"], ref["p_val"], rtol=1e-8, atol=1e-300, equal_nan=True)
    m = api.UpperTriangleDiffcorMat(got, [f"g{i}" for i in range(333)])
    assert m.get_cor_matrix("r_a").shape == (333, 333)


# ----------------------------------------------------------------------------- TOM
@pytest.mark.parametrize("p", [5, 130, 400])
@pytest.mark.parametrize("signed", [True, False])
@pytest.mark.parametrize("version", ["v1", "v2"])
def test_tom_random(api, p, signed, version):
    x = np.random.default_rng(p).normal(size=(30, p))
    cor = o.column_pairwise_cor(x, False)
    np.fill_diagonal(cor, 1.0)
    got = api.calculate_tom(cor, signed, version)
    ref = o.calc_tom(np.abs(cor) if not signed else cor, signed, version)
    assert _maxerr(got, ref) < TOL64
    assert np.array_equal(got, got.T)


@pytest.mark.parametrize("cor_method", ["pearson", "spearman"])
def test_tom_from_exp(api, cor_method):
    from bixverse_b200.synthetic import synthetic_bulk

    x = synthetic_bulk(80, 390, seed=2026)
    sp = cor_method == "spearman"
    for signed in (True, False):
        ref = o.tom_from_exp(x, signed, "v1", sp, 1.0)
        assert _maxerr(api.calculate_tom_from_exp(x, signed, "v1", cor_method), ref) < TOL64
        refb = o.tom_from_exp(x, signed, "v2", sp, 6.0)
        gotb = api.calculate_tom_from_exp(x, signed, "v2", cor_method, beta=6.0, packed=True)
        assert _maxerr(gotb, o.dense_to_upper_triangle(refb, True)) < TOL64


@pytest.mark.parametrize("shift", [True, False])
def test_cor_module_tom_fused(api, shift):
    """cor_module_tom == unpack (diag := 1 when shift) -> rs_tom -> repack, in one call."""
    from bixverse_b200.synthetic import synthetic_bulk

    x = synthetic_bulk(80, 450, seed=13)
    packed = api.rs_cor_upper_triangle(x, True, shift)
    res = api.UpperTriangularSymMat(packed, [f"g{i}" for i in range(450)], shift)
    for signed, version in ((True, "v2"), (False, "v1")):
        got = api.cor_module_tom(res, signed, version)
        dense = o.upper_triangle_to_dense(packed, shift, 450)
        ref = o.dense_to_upper_triangle(o.calc_tom(dense, signed, version), True)  # rs_tom level: the matrix as passed
        assert got.shift and _maxerr(got.values, ref) < TOL64 * max(1.0, float(np.abs(ref).max()))
        chain = api.rs_dense_to_upper_triangle(api.rs_tom(api.rs_upper_triangle_to_dense(packed, shift, 450), version, signed), True)
        assert _maxerr(got.values, chain) < 1e-13


def test_tom_invalid_version_is_error(api):
    with pytest.raises(ValueError, match="Invalid TOM version type"):
        api.rs_tom(np.eye(4), "v9", True)
    from bixverse_b200 import _lib

    a = np.eye(4, order="F")
    out = np.empty((4, 4), order="F")
    rc = _lib.load().bixcor_tom(a.ctypes.data, 4, 1, 7, 0, out.ctypes.data)
    assert rc == _lib.ERR_INVALID and b"Invalid TOM version type" in _lib.load().bixcor_last_error()


# ----------------------------------------------------------------------------- sparse single-cell
def test_sparse_gram_cor(api):
    from bixverse_b200.synthetic import synthetic_sparse_cells

    cells, genes = 9000, 300  # > one 8192-cell panel: exercises accumulation
    ip, ix, d = synthetic_sparse_cells(cells, genes, 2027, density=0.05)
    got = api.sparse_gene_cor_upper_triangle(ip, ix, d, cells, genes)
    ref = o.sparse_gram_cor(ip, ix, d, cells, genes)
    assert _maxerr(got, ref) < 1e-9
    i, j = o.upper_triangle_indices(genes, 1)
    sel = np.arange(0, i.shape[0], 997)
    pl = o.pairwise_gene_cors(ip, ix, d, cells, genes, i[sel], j[sel])  # rs_pairwise_gene_cors_mc semantics
    ok = ~np.isnan(pl)
    assert np.abs(pl[ok] - got[sel][ok]).max() < 1e-9


@pytest.mark.parametrize("kind", ["tf32", "f16"])
def test_sparse_gram_cor_tf32_fast_path(api, kind):
    from bixverse_b200 import _lib
    from bixverse_b200.synthetic import synthetic_sparse_cells

    prec = {"tf32": _lib.PREC_TF32X3, "f16": _lib.PREC_F16X3}[kind]
    cells, genes = 20000, 300  # three 8192-cell panels
    ip, ix, d = synthetic_sparse_cells(cells, genes, 2028, density=0.05)
    got = api.sparse_gene_cor_upper_triangle(ip, ix, d, cells, genes, precision=prec)
    ref = o.sparse_gram_cor(ip, ix, d, cells, genes)
    assert _maxerr(got, ref) < TOL32
    if kind == "f16":  # values that do not fit the scaled binary16 planes are reported, not turned into infinities
        big = d.copy()
        big[5] = 100.0
        with pytest.raises(api.BixcorError):
            api.sparse_gene_cor_upper_triangle(ip, ix, big, cells, genes, precision=prec)
        a = np.eye(200) * 1.0
        a[3, 7] = a[7, 3] = 80.0
        with pytest.raises(api.BixcorError):
            api.rs_tom(a, "v1", True, precision=prec)


# ----------------------------------------------------------------------------- fast / exact tensor paths
def _precisions():
    from bixverse_b200 import _lib

    return _lib.PREC_TF32X3, _lib.PREC_I8EXACT


@pytest.mark.parametrize("n,p", [(200, 2000), (1000, 257), (57, 129)])
def test_spearman_int8_exact_path(api, n, p):
    from bixverse_b200.synthetic import synthetic_bulk

    _, i8 = _precisions()
    x = synthetic_bulk(n, p, seed=3 * n + p, tie_heavy=True)
    got = api.rs_cor_upper_triangle(x, True, True, precision=i8)
    assert _maxerr(got, o.cor_upper_triangle(x, True, True)) < 1e-12
    got6 = api.rs_cor_upper_triangle(x, True, False, beta=6.0, precision=i8)
    assert _maxerr(got6, o.soft_threshold(o.cor_upper_triangle(x, True, False), 6.0)) < 1e-12


def _split_prec(name):
    from bixverse_b200 import _lib

    return {"tf32": _lib.PREC_TF32X3, "f16": _lib.PREC_F16X3}[name]


@pytest.mark.parametrize("n,p", [(200, 2000), (1000, 257), (57, 129)])
@pytest.mark.parametrize("spearman", [False, True])
@pytest.mark.parametrize("kind", ["tf32", "f16"])
def test_tf32x3_fast_path(api, n, p, spearman, kind):
    """3 x TF32 and 3 x binary16 hi / lo splits (tcgen05 kind::tf32 / kind::f16): the north star's 1e-5 FP32-path bar."""
    from bixverse_b200.synthetic import synthetic_bulk

    tf32 = _split_prec(kind)
    x = synthetic_bulk(n, p, seed=5 * n + p)
    got = api.rs_cor_upper_triangle(x, spearman, True, precision=tf32)
    assert _maxerr(got, o.cor_upper_triangle(x, spearman, True)) < TOL32
    dense = api.rs_cor(x, spearman, precision=tf32)
    ref = o.column_pairwise_cor(x, spearman)
    np.fill_diagonal(ref, 1.0)
    assert _maxerr(dense, ref) < TOL32


@pytest.mark.parametrize("kind", ["tf32", "f16"])
def test_tom_tf32x3_fast_path(api, kind):
    from bixverse_b200.synthetic import synthetic_bulk

    tf32 = _split_prec(kind)
    x = synthetic_bulk(100, 520, seed=8)
    for signed in (True, False):
        for version in ("v1", "v2"):
            got = api.calculate_tom_from_exp(x, signed, version, "pearson", beta=2.0, precision=tf32)
            assert _maxerr(got, o.tom_from_exp(x, signed, version, False, 2.0)) < TOL32
    cor = np.corrcoef(synthetic_bulk(60, 1300, seed=9).T)  # K = 1300: many promotion chunks, ragged last K block
    got = api.rs_tom(cor, "v2", True, precision=tf32)
    off = ~np.eye(1300, dtype=bool)
    assert np.abs(got - o.calc_tom(cor, True, "v2"))[off].max() < TOL32
    packed = api.cor_module_tom(api.UpperTriangularSymMat(o.dense_to_upper_triangle(cor, True), list(range(1300)), True),
                                True, "v2", precision=tf32)
    assert _maxerr(packed.values, o.dense_to_upper_triangle(o.calc_tom(cor, True, "v2"), True)) < TOL32


def test_diffcor_fast_paths(api):
    from bixverse_b200.synthetic import synthetic_pair

    tf32, i8 = _precisions()
    xa, xb = synthetic_pair(60, 75, 333, seed=2025)
    ref = o.differential_cor(xa, xb, True)
    got = api.rs_differential_cor(xa, xb, True, precision=i8)
    assert _maxerr(got["r_a"], ref["r_a"]) < 1e-12 and _maxerr(got["z_score"], ref["z_score"]) < 1e-9
    got = api.rs_differential_cor(xa, xb, False, precision=tf32)
    refp = o.differential_cor(xa, xb, False)
    assert _maxerr(got["r_b"], refp["r_b"]) < TOL32 and _maxerr(got["z_score"], refp["z_score"]) < 1e-3


# ----------------------------------------------------------------------------- full-size properties (BASELINE sizes)
def test_config2_full_size_properties(api):
    """Spearman 20k x 1k + |r|^6 at full size: sampled rows against the oracle, plus
    size-independent properties (range, symmetry of a mirrored sub-block, sharded == unsharded)."""
    from bixverse_b200.synthetic import synthetic_bulk

    n, p = 1000, 20000
    x = synthetic_bulk(n, p, seed=2024, tie_heavy=True)
    got = api.rs_cor_upper_triangle(x, True, True, beta=6.0)
    assert got.shape[0] == 199_990_000
    assert np.nanmin(got) >= 0.0 and np.nanmax(got) <= 1.0 + 1e-12
    # oracle on sampled rows: correlation of gene i against all j > i
    ranks = o.rank_matrix_col(x[:, :])
    z = ranks - ranks.mean(axis=0)
    z /= np.sqrt((z * z).sum(axis=0))
    for i in (0, 1, 127, 128, 5000, 12345, 19871, 19998):
        ref = np.abs(z[:, i] @ z[:, i + 1 :]) ** 6.0
        off = int(o.packed_offset(i, p, 1))
        assert np.abs(got[off : off + p - 1 - i] - ref).max() < TOL64, i
    r0, r1 = api.shard_rows(p, True, 8, 5)
    part = api.rs_cor_upper_triangle(x, True, True, beta=6.0, rows=(r0, r1))
    assert np.array_equal(part, got[int(o.packed_offset(r0, p, 1)) : int(o.packed_offset(r1, p, 1))])


# ----------------------------------------------------------------------------- several devices in one process
def test_in_process_multi_device(api):
    """bixcor_init(n): the packed triangle is split over the process's devices by pair count."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from bixverse_b200.synthetic import synthetic_bulk, synthetic_pair

    x = synthetic_bulk(100, 1500, seed=31)
    single = api.rs_cor_upper_triangle(x, True, True, beta=3.0)
    xa, xb = synthetic_pair(60, 70, 700, seed=5)
    dsingle = api.rs_differential_cor(xa, xb, False)
    try:
        api.init(2, 0)
        multi = api.rs_cor_upper_triangle(x, True, True, beta=3.0)
        dmulti = api.rs_differential_cor(xa, xb, False)
    finally:
        api.init(1, 0)
    bad = np.flatnonzero(single != multi)
    assert bad.size == 0, f"{bad.size} of {single.size} differ, first at {bad[:5]}, single {single[bad[:3]]} multi {multi[bad[:3]]}"
    for k in dsingle:
        assert np.array_equal(dsingle[k], dmulti[k], equal_nan=True)


@pytest.mark.parametrize("prec_name", ["f64", "tf32", "ozaki"])
def test_in_process_multi_device_tom(api, prec_name):
    """bixcor_init(n) with n > 1: slab affinities per device, NVLink peer exchange, k through the host, TOM rows per device."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from bixverse_b200 import _lib
    from bixverse_b200.synthetic import synthetic_bulk

    prec = {"f64": _lib.PREC_F64, "tf32": _lib.PREC_TF32X3, "ozaki": _lib.PREC_OZAKI}[prec_name]
    tol = TOL32 if prec_name == "tf32" else TOL64
    x = synthetic_bulk(90, 900, seed=41)
    cor = np.corrcoef(x.T)
    try:
        api.init(2, 0)
        for signed in (True, False):
            ref = o.tom_from_exp(x, signed, "v1", False, 2.0)
            off = ~np.eye(900, dtype=bool)
            dense = api.calculate_tom_from_exp(x, signed, "v1", "pearson", beta=2.0, precision=prec, packed=False)
            assert np.abs(dense - ref)[off].max() < tol
            packed = api.calculate_tom_from_exp(x, signed, "v1", "pearson", beta=2.0, precision=prec, packed=True)
            assert _maxerr(packed, o.dense_to_upper_triangle(ref, True)) < tol
            got = api.calculate_tom(cor, signed, "v2", precision=prec)
            ref2 = o.calc_tom(cor if signed else np.abs(cor), signed, "v2")
            assert np.abs(got - ref2)[off].max() < tol
    finally:
        api.init(1, 0)


@pytest.mark.parametrize("prec_name", ["f64", "i8", "tf32", "f16"])
def test_operand_exchange_api(api, prec_name):
    """Gene slabs of the Gram operand built separately (as different ranks would) give the same packed result."""
    import torch

    from bixverse_b200 import _lib
    from bixverse_b200.distributed import DeviceBackend
    from bixverse_b200.synthetic import synthetic_bulk

    prec = {"f64": _lib.PREC_F64, "i8": _lib.PREC_I8EXACT, "tf32": _lib.PREC_TF32X3, "f16": _lib.PREC_F16X3}[prec_name]
    n, p = 90, 700
    x = synthetic_bulk(n, p, seed=12, tie_heavy=True)
    dev = torch.device("cuda", 0)
    be = DeviceBackend(dev)
    xd = torch.from_numpy(np.ascontiguousarray(x.T)).to(dev)
    rb = be.operand_row_bytes(n, prec)
    operand = be.empty_bytes(p * rb)
    inv = be.zeros(p)
    for g0, g1 in ((0, 256), (256, 640), (640, p)):
        be.build_operand(xd, n, p, True, prec, g0, g1, operand, inv)
    inv2 = be.zeros(p)
    be.operand_inv_norm(operand, n, p, prec, inv2)  # norms rebuilt from the records == norms written by the rank kernel
    if prec_name == "i8":
        assert torch.equal(inv2, inv)
    direct = be.cor_upper_rows(xd, n, p, True, 1, 6.0, prec, 0, p)
    for r0, r1 in ((0, p), (128, 512)):
        via = be.cor_upper_rows_op(operand, inv, n, p, prec, 1, 6.0, r0, r1)
        o0 = int(o.packed_offset(r0, p, 1))
        assert torch.equal(via, direct[o0 : o0 + via.numel()])
    ref = o.soft_threshold(o.cor_upper_triangle(x, True, True), 6.0)
    assert _maxerr(direct.cpu().numpy(), ref) < (TOL32 if prec_name in ("tf32", "f16") else TOL64)


@pytest.mark.parametrize("pinned", [True, False])
def test_host_sharded_driver_single_rank(api, pinned):
    """host -> host driver of the multi-process path (slab upload, operand build, band-pipelined D2H), world = 1."""
    import torch

    from bixverse_b200 import _lib
    from bixverse_b200.distributed import DeviceBackend, host_sharded_cor_upper
    from bixverse_b200.synthetic import synthetic_bulk

    n, p = 120, 1500
    x = synthetic_bulk(n, p, seed=21, tie_heavy=True)
    xt = torch.from_numpy(np.ascontiguousarray(x.T))
    out = torch.empty(p * (p - 1) // 2, dtype=torch.float64)
    if pinned:
        xt, out = xt.pin_memory(), out.pin_memory()
    be = DeviceBackend(torch.device("cuda", 0))
    ws = {}
    for _ in range(2):  # second call reuses the cached workspace
        out.fill_(-7.0)
        o0, o1 = host_sharded_cor_upper(be, xt, n, p, True, out, 1, 6.0, _lib.PREC_I8EXACT, workspace=ws)
        assert (o0, o1) == (0, out.numel())
        ref = o.soft_threshold(o.cor_upper_triangle(x, True, True), 6.0)
        assert _maxerr(out.numpy(), ref) < TOL64


# ----------------------------------------------------------------------------- sibling Gram kernels
def test_covariance_cor2_cov2cor_golden(api, golden_dir):
    """inst/tinytest/test_R_rust_equality.R:23-55,76-82 on the reference's own matrices."""
    g = json.load(open(os.path.join(golden_dir, "r_equality_seed123.json")))
    mat, mat2 = np.array(g["mat"]), np.array(g["mat_2"])
    cov = api.rs_covariance(mat)
    assert _maxerr(cov, np.array(g["cov"])) < 1e-13
    assert _maxerr(api.rs_cor2(mat, mat2, False), np.array(g["pearson_2"])) < 1e-13
    assert _maxerr(api.rs_cor2(mat, mat2, True), np.array(g["spearman_2"])) < 1e-13
    assert _maxerr(api.rs_cov2cor(cov), np.array(g["pearson"])) < 1e-13


@pytest.mark.parametrize("spearman", [False, True])
def test_rbh_cor_known_answer_and_oracle(api, golden_dir, spearman):
    """rs_rbh_cor: the reference's own known answers (inst/tinytest/test_rbh.R:109-214) and larger seeded cases
    against the oracle (k_best 1 / 3, thresholds, negative correlations, more than one tile per side)."""
    import json, os

    with open(os.path.join(golden_dir, "rbh_cor_seed42.json")) as f:
        g = json.load(f)
    mats = {
        "origin_1": (np.array(g["matrix_a"]), g["rownames_a"], g["colnames_a"]),
        "origin_2": (np.array(g["matrix_b"]), g["rownames_b"], g["colnames_b"]),
        "origin_3": (np.array(g["matrix_c"]), g["rownames_c"], g["colnames_c"]),
    }
    res = api.rs_rbh_cor(mats, 1, spearman, 0.0)
    assert res["origin"] == ["origin_1", "origin_1", "origin_2"] and res["target"] == ["origin_2", "origin_3", "origin_3"]
    assert res["comparisons"] == [3, 0, 0]
    assert res["origin_modules"] == g["expected_origin_modules"] and res["target_modules"] == g["expected_target_modules"]
    np.testing.assert_allclose(res["similarity"], g["expected_spearman" if spearman else "expected_pearson"], atol=g["tolerance"])

    rng = np.random.default_rng(77)
    feats = [f"g{i}" for i in range(300)]
    big = {}
    for name, ncol, lo in (("ica", 150, 0), ("nmf", 140, 40), ("pca", 33, 100)):
        m = rng.normal(size=(200, ncol))
        big[name] = (m, feats[lo:lo + 200], [f"{name}_{k}" for k in range(ncol)])
    big["nmf"][0][:160, 5] = -big["ica"][0][40:, 9]  # shared features line up: a strong NEGATIVE correlation must be a hit
    for k_best, thr in ((1, 0.0), (3, 0.1), (2, 0.25)):
        got = api.rs_rbh_cor(big, k_best, spearman, thr)
        ref = o.rbh_cor_list(big, k_best, spearman, thr)
        for key in ("origin", "target", "comparisons", "origin_modules", "target_modules"):
            assert got[key] == ref[key], key
        assert _maxerr(np.array(got["similarity"]), np.array(ref["similarity"])) < TOL64


def test_siblings_larger_shapes(api):
    from bixverse_b200.synthetic import synthetic_bulk

    x = synthetic_bulk(120, 390, seed=41)
    y = synthetic_bulk(120, 200, seed=42)
    assert _maxerr(api.rs_covariance(x), o.column_pairwise_cov(x)) < 1e-10
    c2 = api.rs_cor2(x, y, True)
    rx, ry = o.rank_matrix_col(x), o.rank_matrix_col(y)
    ref = np.corrcoef(rx.T, ry.T)[:390, 390:]
    assert c2.shape == (390, 200) and _maxerr(c2, ref) < 1e-10
    xn = x / np.sqrt((x * x).sum(axis=0))
    assert _maxerr(api.rs_cos(x), xn.T @ xn) < 1e-10


# ----------------------------------------------------------------------------- RBF affinities + CoReMo stability
@pytest.mark.parametrize("rbf_type", ["gaussian", "bump", "inverse_quadratic"])
def test_rbf_function_and_epsilon_sweep(api, rbf_type):
    rng = np.random.default_rng(5)
    d = np.concatenate([np.linspace(0.0, 1.0, 101), rng.uniform(0, 1, 5000)])
    for eps in (0.25, 1.0, 2.0, 7.5):
        assert _maxerr(api.rs_rbf_function(d, eps, rbf_type), o.rbf_function(d, eps, rbf_type)) < 1e-14
    m = rng.uniform(0, 1, size=(37, 11))
    assert _maxerr(api.rs_rbf_function_mat(m, 2.0, rbf_type), o.rbf_function(m, 2.0, rbf_type)) < 1e-14
    eps_vec = np.sort(np.array([0.25] + list(np.arange(0.5, 10.5, 0.5))))[::-1]  # R/methods_coexp_cor.R:314,353
    for p, shift in ((300, True), (131, False), (1100, True)):
        x = rng.normal(size=(40, p))
        dist = 1.0 - np.abs(o.cor_upper_triangle(x, False, shift))
        dist[dist < 0] = 0.0
        got = api.rs_rbf_iterate_epsilons(dist, eps_vec, p, shift, rbf_type)
        ref = o.rbf_iterate_epsilons(dist, eps_vec, p, shift, rbf_type)
        assert got.shape == (p, eps_vec.shape[0])
        assert np.abs(got - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())
    with pytest.raises(ValueError):
        api.rs_rbf_function(d, 1.0, "laplace")


@pytest.mark.parametrize("spearman", [False, True])
@pytest.mark.parametrize("rbf_type", ["gaussian", "bump"])
def test_coremo_stability(api, spearman, rbf_type):
    from bixverse_b200.synthetic import synthetic_bulk

    n, p = 60, 400
    x = synthetic_bulk(n, p, seed=31, tie_heavy=spearman)
    idx = [1, 2, 17, n]
    got = api.rs_coremo_stability(x, idx, 2.0, rbf_type, spearman)
    ref = o.coremo_stability(x, idx, 2.0, rbf_type, spearman)
    assert len(got) == len(idx)
    for g, r in zip(got, ref):
        assert _maxerr(g, r) < TOL64
    if spearman:
        _, i8 = _precisions()
        got8 = api.rs_coremo_stability(x, idx, 2.0, rbf_type, True, precision=i8)
        for g, r in zip(got8, ref):
            assert _maxerr(g, r) < 1e-12
    fused = api.cor_upper_triangle_rbf(x, spearman, 2.0, rbf_type)
    assert _maxerr(fused, o.rbf_function(1.0 - np.abs(o.cor_upper_triangle(x, spearman, True)), 2.0, rbf_type)) < TOL64
    tf32, _ = _precisions()
    fused32 = api.cor_upper_triangle_rbf(x, spearman, 2.0, rbf_type, precision=tf32)
    assert _maxerr(fused32, fused) < 2e-5
    with pytest.raises(api.BixcorError):
        api.rs_coremo_stability(x, [0], 2.0, rbf_type, spearman)


# ----------------------------------------------------------------------------- sparse pair list
@pytest.mark.parametrize("fmt", ["csr", "csc"])
@pytest.mark.parametrize("spearman", [False, True])
def test_pairwise_gene_cors_mc(api, fmt, spearman):
    import scipy.sparse as sp

    from bixverse_b200.synthetic import synthetic_sparse_cells

    cells, genes = 900, 60
    ip, ix, d = synthetic_sparse_cells(cells, genes, 11, density=0.15)
    rng = np.random.default_rng(3)
    g1 = rng.integers(0, genes, size=200).astype(np.int32)
    g2 = rng.integers(0, genes, size=200).astype(np.int32)
    m = sp.csr_matrix((d, ix, ip), shape=(cells, genes))
    if fmt == "csc":
        m = m.tocsc()
    sparse_list = {"data": m.data, "indptr": m.indptr, "indices": m.indices, "nrow": cells, "ncol": genes, "format": fmt}
    got = api.rs_pairwise_gene_cors_mc(sparse_list, g1, g2, spearman)
    ref = (o.pairwise_gene_cors_spearman if spearman else o.pairwise_gene_cors)(ip, ix, d, cells, genes, g1, g2)
    ok = ~np.isnan(ref)
    assert ok.sum() > 150 and np.abs(got[ok] - ref[ok]).max() < 1e-10
    if not spearman:  # the pair list is the all-pairs Gram restricted to the listed pairs (finding 3)
        allp = api.sparse_gene_cor_upper_triangle(ip, ix, d, cells, genes)
        dense = o.upper_triangle_to_dense(allp, True, genes)
        off = (g1 != g2) & ok
        assert np.abs(got[off] - dense[g1[off], g2[off]]).max() < 1e-10
    with pytest.raises(api.BixcorError):
        api.rs_pairwise_gene_cors_mc(sparse_list, [genes], [0], spearman)


@pytest.mark.parametrize("fmt", ["csr", "csc"])
@pytest.mark.parametrize("spearman", [False, True])
def test_pairwise_gene_cors_cells_to_keep(api, fmt, spearman):
    """rs_pairwise_gene_cors semantics: the correlation runs over the listed cells only (zeros included)."""
    import scipy.sparse as sp

    from bixverse_b200.synthetic import synthetic_sparse_cells

    cells, genes = 1100, 40
    ip, ix, d = synthetic_sparse_cells(cells, genes, 21, density=0.2)
    rng = np.random.default_rng(8)
    keep = rng.permutation(cells)[:613].astype(np.int32)  # unsorted on purpose
    g1 = rng.integers(0, genes, size=120).astype(np.int32)
    g2 = rng.integers(0, genes, size=120).astype(np.int32)
    m = sp.csr_matrix((d, ix, ip), shape=(cells, genes))
    sub = sp.csr_matrix(m[keep])
    ref = (o.pairwise_gene_cors_spearman if spearman else o.pairwise_gene_cors)(
        sub.indptr.astype(np.int64), sub.indices.astype(np.int32), sub.data, len(keep), genes, g1, g2)
    mm = m.tocsc() if fmt == "csc" else m
    sparse_list = {"data": mm.data, "indptr": mm.indptr, "indices": mm.indices, "nrow": cells, "ncol": genes, "format": fmt}
    got = api.rs_pairwise_gene_cors_mc(sparse_list, g1, g2, spearman, cells_to_keep=keep)
    ok = ~np.isnan(ref)
    assert ok.sum() > 100 and np.abs(got[ok] - ref[ok]).max() < 1e-10
    for bad in ([0, 0, 1], [cells, 1, 2], [5]):
        with pytest.raises(api.BixcorError):
            api.rs_pairwise_gene_cors_mc(sparse_list, g1, g2, spearman, cells_to_keep=bad)


@pytest.mark.parametrize("fmt", ["csr", "csc"])
def test_pairwise_spearman_many_cells(api, fmt):
    """Spearman over more cells than the dense rank kernel takes: ranks of the non-zeros + analytic zero rank
    (negative values, heavy ties and explicit zeros included) against dense rankdata."""
    import scipy.sparse as sp

    cells, genes = 40000, 24
    rng = np.random.default_rng(31)
    dense = np.zeros((cells, genes), dtype=np.float32)
    for g in range(genes):
        nnz = int(cells * rng.uniform(0.01, 0.12))
        pos = rng.choice(cells, nnz, replace=False)
        vals = np.round(rng.normal(0.5, 1.0, nnz), 1 if g % 2 else 3)  # ties; some values round to 0 and some are negative
        dense[pos, g] = vals.astype(np.float32)
    dense[:, 5] = np.abs(dense[:, 5])  # an all-non-negative gene
    dense[:, 7] = 0.0                   # an all-zero gene: zero variance -> NaN
    dense[:, 9] = 0.7 * dense[:, 3] + dense[:, 9] * (rng.random(cells) < 0.3)
    m = sp.csr_matrix(dense)
    g1, g2 = np.triu_indices(genes, 1)
    g1, g2 = g1.astype(np.int32), g2.astype(np.int32)
    ref = o.pairwise_gene_cors_spearman(m.indptr.astype(np.int64), m.indices.astype(np.int32), m.data, cells, genes, g1, g2)
    mm = m.tocsc() if fmt == "csc" else m
    sparse_list = {"data": mm.data, "indptr": mm.indptr, "indices": mm.indices, "nrow": cells, "ncol": genes, "format": fmt}
    got = api.rs_pairwise_gene_cors_mc(sparse_list, g1, g2, True)
    ok = ~np.isnan(ref)
    assert ok.sum() == len(g1) - (genes - 1) and np.isnan(got[~ok]).all()
    assert np.abs(got[ok] - ref[ok]).max() < 1e-10
    # genes with more than 16384 non-zero cells (too many for one CTA's shared memory) take the chunked dense rank kernels
    dense[:20000, 2] = 1.5                                        # one huge tie run
    dense[:, 11] = np.round(rng.normal(0.0, 1.0, cells), 2)       # a fully dense gene with ties and negatives
    m2 = sp.csr_matrix(dense)
    ref2 = o.pairwise_gene_cors_spearman(m2.indptr.astype(np.int64), m2.indices.astype(np.int32), m2.data, cells, genes, g1, g2)
    mm2 = m2.tocsc() if fmt == "csc" else m2
    sl2 = {"data": mm2.data, "indptr": mm2.indptr, "indices": mm2.indices, "nrow": cells, "ncol": genes, "format": fmt}
    got2 = api.rs_pairwise_gene_cors_mc(sl2, g1, g2, True)
    ok2 = ~np.isnan(ref2)
    assert np.isnan(got2[~ok2]).all() and np.abs(got2[ok2] - ref2[ok2]).max() < 1e-10


# ----------------------------------------------------------------------------- FP64-class Gram on the int8 tensor cores (Ozaki slices)
@pytest.mark.parametrize("n,p", [(12, 14), (200, 2000), (57, 129), (1000, 257), (1500, 300)])
@pytest.mark.parametrize("spearman", [False, True])
def test_ozaki_int8_fp64_class_correlation(api, n, p, spearman):
    from bixverse_b200 import _lib
    from bixverse_b200.synthetic import synthetic_bulk

    oz = _lib.PREC_OZAKI
    x = synthetic_bulk(n, p, seed=7 * n + p) if n >= 50 else np.random.default_rng(n + p).normal(size=(n, p))
    dense_ref = o.column_pairwise_cor(x, spearman)
    for shift in (True, False):
        got = api.rs_cor_upper_triangle(x, spearman, shift, precision=oz)
        assert _maxerr(got, o.dense_to_upper_triangle(dense_ref, shift)) < TOL64
    got6 = api.rs_cor_upper_triangle(x, spearman, True, beta=6.0, precision=oz)
    assert _maxerr(got6, o.soft_threshold(o.dense_to_upper_triangle(dense_ref, True), 6.0)) < TOL64
    dense = api.rs_cor(x, spearman, precision=oz)
    ref = dense_ref.copy()
    np.fill_diagonal(ref, 1.0)
    assert _maxerr(dense, ref) < TOL64
    assert np.array_equal(dense, dense.T, equal_nan=True)
    r0, r1 = 128, min(p, 256)
    if p > 128:
        part = api.rs_cor_upper_triangle(x, spearman, True, precision=oz, rows=(r0, r1))
        full = api.rs_cor_upper_triangle(x, spearman, True, precision=oz)
        assert np.array_equal(part, full[int(o.packed_offset(r0, p, 1)) : int(o.packed_offset(r1, p, 1))])


@pytest.mark.parametrize("p", [5, 130, 400])
@pytest.mark.parametrize("signed", [True, False])
def test_ozaki_tom(api, p, signed):
    from bixverse_b200 import _lib
    from bixverse_b200.synthetic import synthetic_bulk

    rng = np.random.default_rng(p)
    cor = np.corrcoef(rng.normal(size=(3 * p, p)).T)
    for version in ("v1", "v2"):
        got = api.calculate_tom(cor, signed, version, precision=_lib.PREC_OZAKI)
        ref = o.calc_tom(cor if signed else np.abs(cor), signed, version)
        off = ~np.eye(p, dtype=bool)
        assert np.abs(got - ref)[off].max() < TOL64
    x = synthetic_bulk(150, 500, seed=9)
    for packed in (False, True):
        got = api.calculate_tom_from_exp(x, signed, "v1", "pearson", beta=2.0, precision=_lib.PREC_OZAKI, packed=packed)
        ref = o.tom_from_exp(x, signed, "v1", False, 2.0)
        if packed:
            assert _maxerr(got, o.dense_to_upper_triangle(ref, True)) < TOL64
        else:
            off = ~np.eye(500, dtype=bool)
            assert np.abs(got - ref)[off].max() < TOL64


def test_ozaki4_medium_precision(api):
    """Four leading Ozaki slices (three int8 passes): the cost class of 3xTF32, errors around 1e-8."""
    from bixverse_b200 import _lib
    from bixverse_b200.synthetic import synthetic_bulk

    oz4 = _lib.PREC_OZAKI4
    x = synthetic_bulk(300, 700, seed=77)
    for spearman in (False, True):
        got = api.rs_cor_upper_triangle(x, spearman, True, precision=oz4)
        assert _maxerr(got, o.cor_upper_triangle(x, spearman, True)) < 1e-7
    for signed in (True, False):
        got = api.calculate_tom_from_exp(x, signed, "v1", "pearson", beta=2.0, precision=oz4, packed=True)
        assert _maxerr(got, o.dense_to_upper_triangle(o.tom_from_exp(x, signed, "v1", False, 2.0), True)) < 1e-7


def test_ozaki_diffcor(api):
    from bixverse_b200 import _lib
    from bixverse_b200.synthetic import synthetic_pair

    xa, xb = synthetic_pair(60, 75, 333, seed=2025)
    for spearman in (False, True):
        ref = o.differential_cor(xa, xb, spearman)
        got = api.rs_differential_cor(xa, xb, spearman, precision=_lib.PREC_OZAKI)
        assert _maxerr(got["r_a"], ref["r_a"]) < TOL64 and _maxerr(got["r_b"], ref["r_b"]) < TOL64
        assert _maxerr(got["z_score"], ref["z_score"]) < 1e-8
        ok = ref["p_val"] > 1e-280
        assert np.abs(got["p_val"][ok] / ref["p_val"][ok] - 1.0).max() < 1e-6


def test_ozaki_nan_and_constant_columns(api):
    from bixverse_b200 import _lib

    rng = np.random.default_rng(4)
    x = rng.normal(size=(40, 200))
    x[:, 3] = 2.0
    x[5, 7] = np.nan
    got = api.rs_cor(x, False, precision=_lib.PREC_OZAKI)
    ref = api.rs_cor(x, False)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.abs(got[ok] - ref[ok]).max() < TOL64


# ----------------------------------------------------------------------------- boundary behaviour (round 2)
def test_row_range_contract(api):
    """bixcor.h: row_begin % 128 == 0, row_end % 128 == 0 or row_end == p; empty ranges are no-ops wherever they start
    (a ragged row_end wrote past the caller's slice; empty shards of a small module were rejected)."""
    from bixverse_b200.synthetic import synthetic_bulk

    x = synthetic_bulk(50, 300, seed=3)
    for bad in ((0, 200), (128, 129), (64, 300)):
        with pytest.raises(api.BixcorError) as ei:
            api.rs_cor_upper_triangle(x, False, True, rows=bad)
        assert ei.value.code == -2
    assert api.rs_cor_upper_triangle(x, False, True, rows=(300, 300)).shape[0] == 0
    assert api.rs_cor_upper_triangle(x, False, True, rows=(77, 77)).shape[0] == 0
    # a 120-gene module over 8 ranks: seven empty shards + one that owns the whole triangle
    x = synthetic_bulk(40, 120, seed=4)
    full = api.rs_cor_upper_triangle(x, True, True)
    parts = [api.rs_cor_upper_triangle(x, True, True, rows=api.shard_rows(120, True, 8, r)) for r in range(8)]
    assert np.array_equal(np.concatenate(parts), full)
    xa, xb = synthetic_bulk(30, 120, seed=5), synthetic_bulk(33, 120, seed=6)
    lib = __import__("bixverse_b200._lib", fromlist=["load"]).load()
    a, b = np.asfortranarray(xa), np.asfortranarray(xb)
    dummy = np.zeros(4)
    assert lib.bixcor_diffcor_rows(a.ctypes.data, 30, b.ctypes.data, 33, 120, 0, 0, 120, 120, *[dummy.ctypes.data] * 4) == 0


def test_precision_auto_policy(api, monkeypatch):
    """BIXCOR_PREC_AUTO (what INTEGRATION.md's patched bodies pass): Spearman with n <= 1860 -> exact-integer tcgen05
    Gram, everything else FP64; env BIXCOR_PRECISION replaces the rule; an undefined choice falls back to it."""
    from bixverse_b200 import _lib
    from bixverse_b200.synthetic import synthetic_bulk

    monkeypatch.delenv("BIXCOR_PRECISION", raising=False)
    x = synthetic_bulk(200, 500, seed=8, tie_heavy=True)
    auto = lambda sp, **kw: api.rs_cor_upper_triangle(x, sp, True, precision=_lib.PREC_AUTO, **kw)
    assert np.array_equal(auto(True), api.rs_cor_upper_triangle(x, True, True, precision=_lib.PREC_I8EXACT))
    assert np.array_equal(auto(False), api.rs_cor_upper_triangle(x, False, True, precision=_lib.PREC_F64))
    assert np.array_equal(auto(True, beta=6.0), api.rs_cor_upper_triangle(x, True, True, beta=6.0, precision=_lib.PREC_I8EXACT))
    xl = np.random.default_rng(2).normal(size=(1900, 130))  # n > 1860: the int32 recombination could overflow -> FP64
    assert np.array_equal(api.rs_cor_upper_triangle(xl, True, True, precision=_lib.PREC_AUTO),
                          api.rs_cor_upper_triangle(xl, True, True, precision=_lib.PREC_F64))
    assert _maxerr(api.rs_cor(x, True, precision=_lib.PREC_AUTO), o.column_pairwise_cor(x, True)) < 1e-12  # dense epilogue too
    d_auto = api.rs_differential_cor(x[:90], x[90:], True, precision=_lib.PREC_AUTO)
    d_i8 = api.rs_differential_cor(x[:90], x[90:], True, precision=_lib.PREC_I8EXACT)
    assert all(np.array_equal(d_auto[k], d_i8[k], equal_nan=True) for k in d_auto)
    t_auto = api.calculate_tom_from_exp(x, True, "v1", "spearman", precision=_lib.PREC_AUTO)
    assert _maxerr(t_auto, o.tom_from_exp(x, True, "v1", True)) < TOL64
    monkeypatch.setenv("BIXCOR_PRECISION", "tf32")
    assert np.array_equal(auto(False), api.rs_cor_upper_triangle(x, False, True, precision=_lib.PREC_TF32X3))
    monkeypatch.setenv("BIXCOR_PRECISION", "i8")
    assert np.array_equal(auto(False), api.rs_cor_upper_triangle(x, False, True, precision=_lib.PREC_F64))  # Pearson: undefined
    monkeypatch.setenv("BIXCOR_PRECISION", "f64")
    assert np.array_equal(auto(True), api.rs_cor_upper_triangle(x, True, True, precision=_lib.PREC_F64))
    monkeypatch.delenv("BIXCOR_PRECISION")
    with pytest.raises(api.BixcorError):
        api.rs_cor_upper_triangle(x, False, True, precision=-7)


def test_exception_barrier_on_a_gpu_entry_point(api):
    from bixverse_b200 import _lib

    x = np.random.default_rng(3).normal(size=(20, 40))
    lib = _lib.load()
    lib.bixcor_debug_inject(1)
    with pytest.raises(api.BixcorError) as ei:
        api.rs_cor(x, False)
    assert ei.value.code == _lib.ERR_OOM and "bixcor_cor_dense" in ei.value.message
    lib.bixcor_debug_inject(2)
    with pytest.raises(api.BixcorError) as ei:
        api.rs_tom(np.eye(5), "v1", True)
    assert ei.value.code == _lib.ERR_INTERNAL
    assert _maxerr(api.rs_cor(x, False), o.column_pairwise_cor(x, False)) < TOL64  # and the library still works


@pytest.mark.parametrize("version", ["v1", "v2"])
def test_rs_tom_unsigned_uses_the_matrix_as_passed(api, version):
    """abs() for unsigned networks belongs to the R wrappers (R/functions_stats.R:52-54), not to rs_tom:
    cor_module_tom (R/methods_coexp_cor.R:151) hands rs_tom the raw correlation.  rs_tom(signed = FALSE) on a matrix with
    negative entries follows the documented unsigned formulas on a_ij itself; calculate_tom takes abs() first."""
    rng = np.random.default_rng(11)
    a = rng.uniform(-0.2, 0.9, size=(150, 150))
    a = 0.5 * (a + a.T)
    np.fill_diagonal(a, 1.0)
    raw = api.rs_tom(a, version, False)
    ref_raw = o.calc_tom(a, False, version)
    off = ~np.eye(150, dtype=bool)
    assert np.abs(raw - ref_raw)[off].max() < 1e-10 * max(1.0, np.abs(ref_raw[off]).max())
    wrapped = api.calculate_tom(a, False, version)
    assert _maxerr(wrapped, o.calculate_tom(a, False, version)) < TOL64
    assert np.abs(raw - wrapped)[off].max() > 1e-3  # the two semantics differ once an entry is negative


# ----------------------------------------------------------------------------- Spearman beyond 16384 samples (round 2)
@pytest.mark.parametrize("n,p", [(16385, 3), (40000, 5), (70001, 2)])
def test_ranks_bit_exact_large_n(api, n, p):
    """rank_matrix_col has no sample limit in the reference (r_singscore.rs:89-95 sorts any column); above one CTA's
    shared memory the library sorts chunks and ranks by counting (columns.cuh: rank_big_*).  Bit-exact incl. ties."""
    rng = np.random.default_rng(n)
    x = rng.normal(size=(n, p))
    x[:, 0] = np.round(x[:, 0], 1)                      # heavy ties
    if p > 1:
        x[: n // 2, 1] = 0.0                            # one tie run across several chunks
        x[7, 1], x[8, 1] = -0.0, 0.0
    if p > 2:
        x[::3, 2] = np.inf
        x[1::3, 2] = -np.inf
    assert np.array_equal(api.rs_rank_matrix_col(x), o.rank_matrix_col(x))
    xn = x.copy()
    xn[n - 1, p - 1] = np.nan
    got = api.rs_rank_matrix_col(xn)
    assert np.isnan(got[:, p - 1]).all() and np.array_equal(got[:, : p - 1], o.rank_matrix_col(x)[:, : p - 1])


@pytest.mark.parametrize("prec_name", ["auto", "tf32"])
def test_spearman_correlation_large_n(api, prec_name):
    from bixverse_b200 import _lib

    rng = np.random.default_rng(77)
    n, p = 20000, 140
    f = rng.normal(size=(n, 3))
    x = np.round(f @ rng.normal(size=(3, p)) + rng.normal(size=(n, p)), 1)
    prec = {"auto": _lib.PREC_AUTO, "tf32": _lib.PREC_TF32X3}[prec_name]
    got = api.rs_cor_upper_triangle(x, True, True, precision=prec)
    ref = o.cor_upper_triangle(x, True, True)
    assert _maxerr(got, ref) < (TOL64 if prec_name == "auto" else TOL32)


# ----------------------------------------------------------------------------- in-process NCCL (round 2)
@pytest.mark.parametrize("use_nccl", [True, False])
def test_in_process_collectives(api, use_nccl, monkeypatch):
    """bixcor_init(2): the library opens libnccl itself (ncclCommInitAll): TOM slabs by grouped ncclBroadcast + k by
    ncclAllReduce, sparse Gram sharded by cells with ncclAllReduce of the partial Gram / column sums, dense correlation by
    gene slabs, leave-one-out stability by samples, empty shards of a small module.  BIXCOR_NCCL=0: the peer-copy path."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from bixverse_b200 import _lib
    from bixverse_b200.synthetic import synthetic_bulk, synthetic_sparse_cells

    lib = _lib.load()
    x = synthetic_bulk(90, 900, seed=43)
    cells, genes = 40000, 300
    ip, ix, dd = synthetic_sparse_cells(cells, genes, 2029, density=0.05)
    small = synthetic_bulk(40, 120, seed=4)
    single = {"dense": api.rs_cor(x, True), "sparse": api.sparse_gene_cor_upper_triangle(ip, ix, dd, cells, genes),
              "stab": api.rs_coremo_stability(x[:, :300], [1, 5, 9, 90], 2.0, "gaussian", True),
              "small": api.rs_cor_upper_triangle(small, True, True), "small_tom": api.calculate_tom_from_exp(small, True, "v1", "pearson")}
    monkeypatch.setenv("BIXCOR_NCCL", "1" if use_nccl else "0")
    try:
        api.init(2, 0)
        assert (lib.bixcor_nccl_version() >= 20000) == use_nccl
        for signed in (True, False):
            ref = o.tom_from_exp(x, signed, "v2", True, 1.0)
            packed = api.calculate_tom_from_exp(x, signed, "v2", "spearman", packed=True)
            assert _maxerr(packed, o.dense_to_upper_triangle(ref, True)) < TOL64
        assert _maxerr(api.rs_cor(x, True), single["dense"]) < 1e-14  # strips compute (i, j) and (j, i) separately: last-bit differences
        got_sparse = api.sparse_gene_cor_upper_triangle(ip, ix, dd, cells, genes)
        assert _maxerr(got_sparse, o.sparse_gram_cor(ip, ix, dd, cells, genes)) < 1e-9
        assert _maxerr(got_sparse, single["sparse"]) < 1e-12  # (the sum order differs between one and two shards)
        stab = api.rs_coremo_stability(x[:, :300], [1, 5, 9, 90], 2.0, "gaussian", True)
        assert all(np.array_equal(a, b) for a, b in zip(stab, single["stab"]))
        assert np.array_equal(api.rs_cor_upper_triangle(small, True, True), single["small"])  # p = 120: one device owns nothing
        assert np.array_equal(api.calculate_tom_from_exp(small, True, "v1", "pearson"), single["small_tom"])
    finally:
        monkeypatch.delenv("BIXCOR_NCCL")
        api.init(1, 0)


@pytest.mark.parametrize("kind", ["f16", "tf32"])
def test_tom_from_exchanged_operand_planes(api, kind):
    """The multi-rank TOM of the split-float kinds on ONE device: the hi / lo operand planes are built slab by slab (as
    different ranks would), phase 3 runs from the planes alone (a_ij and the diagonal rebuilt from hi + lo) for row ranges
    that tile the packed vector.  bixcor_dev_tom_split_planes / bixcor_dev_tom_rows_planes (bixcor.h)."""
    import torch

    from bixverse_b200 import _lib
    from bixverse_b200 import distributed as D
    from bixverse_b200.synthetic import synthetic_bulk

    prec = {"f16": _lib.PREC_F16X3, "tf32": _lib.PREC_TF32X3}[kind]
    dev = torch.device("cuda", 0)
    be = D.DeviceBackend(dev)
    n, p = 90, 1000
    x = synthetic_bulk(n, p, seed=61)
    xd = torch.from_numpy(np.ascontiguousarray(x.T)).to(dev)
    for signed in (True, False):
        ref = o.dense_to_upper_triangle(o.tom_from_exp(x, signed, "v2", False, 2.0), True)
        a_full = be.empty(p, p)
        k = be.zeros(p)
        rb = be.tom_plane_row_bytes(p, prec)
        planes = be.empty_bytes(p * rb)
        for c0, c1 in ((0, 384), (384, 1000)):  # two "ranks": affinity slab, its connectivity, its planes
            be.affinity_cols(xd, n, p, False, signed, 2.0, prec, c0, c1, a_full)
            be.tom_connectivity(a_full, p, signed, c0, c1, k)
            be.tom_split_planes(a_full, p, prec, c0, c1, planes)
        parts = [be.tom_rows_planes(planes, k, p, signed, _lib.TOM_V2, prec, r0, r1).cpu().numpy() for r0, r1 in ((0, 256), (256, 1000))]
        assert _maxerr(np.concatenate(parts), ref) < TOL32
        whole = be.tom_rows(a_full, k, p, signed, _lib.TOM_V2, prec, 0, p).cpu().numpy()  # the f64-matrix form of the same kind
        assert _maxerr(np.concatenate(parts), whole) < 2e-6


def test_int32_transport_of_the_exact_spearman_gram(api, monkeypatch):
    """The exact Spearman Gram travels as int32 (4 bytes per pair) and the host staging threads finish r = S / (|d_i| |d_j|)
    and the power: bit-identical to the f64 transport (BIXCOR_S32_TRANSPORT=0), for pageable and page-locked buffers, on
    interior, diagonal and ragged tiles, NaN / constant columns included; the library's byte counters show the halved D2H."""
    import ctypes as C

    import torch

    from bixverse_b200 import _lib
    from bixverse_b200.synthetic import synthetic_bulk

    lib = _lib.load()
    n, p = 120, 1300
    x = synthetic_bulk(n, p, seed=91, tie_heavy=True)
    x[:, 7] = 3.0        # zero variance -> NaN row / column
    x[5, 40] = np.nan    # NaN column
    P_ = p * (p - 1) // 2
    res = {}
    try:
        for mode in ("1", "0"):
            monkeypatch.setenv("BIXCOR_S32_TRANSPORT", mode)
            lib.bixcor_shutdown()
            api.init(1, 0)
            for beta in (1.0, 6.0):
                lib.bixcor_transfer_counters(None, None, 1)
                res[(mode, beta, "pageable")] = api.rs_cor_upper_triangle(x, True, True, beta=beta)
                down = C.c_int64()
                lib.bixcor_transfer_counters(None, C.byref(down), 1)
                assert down.value == (4 if mode == "1" else 8) * P_
                xp = torch.from_numpy(np.ascontiguousarray(x.T)).pin_memory()
                out = torch.empty(P_, dtype=torch.float64).pin_memory()
                _lib.check(lib.bixcor_cor_upper(xp.data_ptr(), n, p, 1, 1, beta, _lib.PREC_AUTO, out.data_ptr()))
                res[(mode, beta, "pinned")] = out.numpy().copy()
            res[(mode, 2.0, "pageable")] = api.rs_cor_upper_triangle(x, True, True, beta=2.0)  # other powers: f64 transport
    finally:
        monkeypatch.delenv("BIXCOR_S32_TRANSPORT")
        lib.bixcor_shutdown()
        api.init(1, 0)
    for beta in (1.0, 6.0):
        ref = res[("0", beta, "pageable")]
        for key in (("1", beta, "pageable"), ("1", beta, "pinned"), ("0", beta, "pinned")):
            assert np.array_equal(res[key], ref, equal_nan=True), key
        oref = o.soft_threshold(o.cor_upper_triangle(x, True, True), beta)
        ok = ~np.isnan(oref)
        assert np.isnan(ref[~ok]).all() and np.abs(ref[ok] - oref[ok]).max() < 1e-12
    assert np.array_equal(res[("1", 2.0, "pageable")], res[("0", 2.0, "pageable")], equal_nan=True)


def test_binary32_transport_of_the_split_float_results(api, monkeypatch):
    """Results of the <= 1e-5-class kinds (3xTF32 / 3xF16) are rounded to binary32 before they cross PCIe and widened by the
    host staging threads: against the f64 transport (BIXCOR_F32_TRANSPORT=0) every value agrees to float rounding (relative
    2^-24), the byte counters show the halved D2H, and both stay inside the 1e-5 class against the oracle - packed
    correlation (with and without a power, ragged last block), packed and dense TOM, pageable and page-locked buffers."""
    import ctypes as C

    import torch

    from bixverse_b200 import _lib
    from bixverse_b200.synthetic import synthetic_bulk

    lib = _lib.load()
    n, p = 150, 1100
    x = synthetic_bulk(n, p, seed=57)
    P_ = p * (p - 1) // 2
    res = {}
    try:
        for mode in ("1", "0"):
            monkeypatch.setenv("BIXCOR_F32_TRANSPORT", mode)
            lib.bixcor_shutdown()
            api.init(1, 0)
            for prec in (_lib.PREC_TF32X3, _lib.PREC_F16X3):
                for beta in (1.0, 6.0):
                    lib.bixcor_transfer_counters(None, None, 1)
                    res[(mode, prec, "cor", beta)] = api.rs_cor_upper_triangle(x, False, True, beta=beta, precision=prec)
                    down = C.c_int64()
                    lib.bixcor_transfer_counters(None, C.byref(down), 1)
                    assert down.value == (4 if mode == "1" else 8) * P_
                xp = torch.from_numpy(np.ascontiguousarray(x.T)).pin_memory()
                out = torch.empty(P_, dtype=torch.float64).pin_memory()
                _lib.check(lib.bixcor_cor_upper(xp.data_ptr(), n, p, 0, 1, 1.0, prec, out.data_ptr()))
                res[(mode, prec, "cor_pinned", 1.0)] = out.numpy().copy()
                res[(mode, prec, "tom_packed", 6.0)] = api.calculate_tom_from_exp(x, True, "v1", "pearson", beta=6.0, precision=prec, packed=True)
                res[(mode, prec, "tom_dense", 6.0)] = api.calculate_tom_from_exp(x, True, "v1", "pearson", beta=6.0, precision=prec)
            # the FP64-class kinds never use it
            lib.bixcor_transfer_counters(None, None, 1)
            api.rs_cor_upper_triangle(x, False, True, precision=_lib.PREC_F64)
            down = C.c_int64()
            lib.bixcor_transfer_counters(None, C.byref(down), 1)
            assert down.value == 8 * P_
    finally:
        monkeypatch.delenv("BIXCOR_F32_TRANSPORT")
        lib.bixcor_shutdown()
        api.init(1, 0)
    cor_ref = o.cor_upper_triangle(x, False, True)
    tom_ref = o.tom_from_exp(x, True, "v1", beta=6.0)
    iu = np.triu_indices(p, 1)  # row-major upper triangle == the packed order (shift = 1)
    for key, narrow in res.items():
        if key[0] != "1":
            continue
        wide = res[("0",) + key[1:]]
        assert narrow.shape == wide.shape
        assert np.array_equal(narrow, wide.astype(np.float32).astype(np.float64)), key  # exactly the float rounding of the f64 result
        if key[2].startswith("cor"):
            assert np.abs(narrow - o.soft_threshold(cor_ref, key[3])).max() < TOL32, key
        elif key[2] == "tom_dense":
            assert np.abs(narrow - tom_ref).max() < TOL32, key
        else:
            assert np.abs(narrow - tom_ref[iu]).max() < TOL32, key


def test_fused_operand_exchange_single_rank(api):
    """bixcor_xchg_* with a world of one (the multi-rank form runs in tests/test_gpu_dist_nccl.py): the exchange block, the
    rank kernel's peer-store variant (no peers), the flag barrier on itself and the two alternating buffers give exactly the
    result of the plain call, step after step, for the fused (int8, n <= 1024) and the copy-kernel forms."""
    import torch

    from bixverse_b200 import _lib
    from bixverse_b200 import distributed as D
    from bixverse_b200.synthetic import synthetic_bulk

    dev = torch.device("cuda", 0)
    be = D.DeviceBackend(dev)
    for n, p, precs in ((90, 700, (_lib.PREC_AUTO, _lib.PREC_F64, _lib.PREC_F16X3)), (1100, 260, (_lib.PREC_AUTO,))):
        xs = [synthetic_bulk(n, p, seed=s, tie_heavy=(s == 1)) for s in (1, 2)]
        xd = [torch.from_numpy(np.ascontiguousarray(x.T)).to(dev) for x in xs]
        for prec in precs:
            want = [be.cor_upper_rows(x, n, p, True, 1, 6.0, prec, 0, p).cpu().numpy() for x in xd]
            D.xchg_setup(be, n, p, prec)
            for step in range(4):
                got, (o0, o1) = D.sharded_cor_upper_xchg(be, xd[step % 2], n, p, True, 1, 6.0)
                assert (o0, o1) == (0, p * (p - 1) // 2)
                assert np.array_equal(got.cpu().numpy(), want[step % 2]), (n, p, prec, step)
            D.xchg_teardown(be)
    # misuse is an error, not a crash
    lib = _lib.load()
    out = torch.empty(10, dtype=torch.float64, device=dev)
    assert lib.bixcor_dev_cor_upper_rows_xchg(xd[0].data_ptr(), 1100, 260, 1, 1, 1.0, 0, 260, out.data_ptr()) == _lib.ERR_INVALID


@pytest.mark.parametrize("n", [37, 64, 500, 1000, 1024])
def test_ranks_bit_exact_packed_key_corner_cases(api, n):
    """The register rank kernel sorts ONE double per element (value with its low 10 mantissa bits replaced by the sample
    index) and verifies afterwards that elements agreeing above those bits are true ties; otherwise the column is redone with
    the exact (key, index) sort.  Columns built to hit that machinery: values closer than 1024 ulp (differing only in the
    replaced bits), mixtures of such near-ties and true ties, values next to +-DBL_MAX together with +-inf, subnormals and
    signed zeros, negative near-ties (whose bit patterns order backwards), a constant column."""
    rng = np.random.default_rng(n)
    cols = []
    base = 1.0 + rng.integers(0, 4, size=n) * 2.0 ** -20
    cols.append(base + rng.integers(0, 1024, size=n) * 2.0 ** -52)          # near-ties: differ only in the low 10 bits
    c = np.full(n, 3.0)
    c[::2] = np.nextafter(3.0, 4.0)                                         # two values one ulp apart, each tied many times
    cols.append(c)
    c = -(1.0 + rng.integers(0, 1024, size=n) * 2.0 ** -52)                 # negative near-ties
    cols.append(c)
    big = np.finfo(np.float64).max
    c = rng.normal(size=n)
    c[: min(n, 6)] = [np.inf, big, np.nextafter(big, 0), -np.inf, -big, np.inf][: min(n, 6)]
    cols.append(c)
    tiny = np.finfo(np.float64).smallest_subnormal
    c = rng.integers(-3, 4, size=n) * tiny                                  # subnormals around +-0
    c[0], c[1] = 0.0, -0.0
    cols.append(c)
    cols.append(np.full(n, -7.25))                                          # constant
    cols.append(np.round(rng.normal(size=n), 1))                            # ordinary tie-heavy column
    cols.append(rng.normal(size=n))                                         # ordinary column
    x = np.stack(cols, axis=1)
    assert np.array_equal(api.rs_rank_matrix_col(x), o.rank_matrix_col(x))
    # the same columns through the Spearman correlation (the int8 digit emission of the same kernel)
    got = api.rs_cor_upper_triangle(x, True, True)
    ref = o.cor_upper_triangle(x, True, True)
    ok = ~np.isnan(ref)
    assert np.isnan(got[~ok]).all() and (not ok.any() or np.abs(got[ok] - ref[ok]).max() < 1e-12)
