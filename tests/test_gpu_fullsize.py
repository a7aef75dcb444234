"""BASELINE-size parity on the GPU, every precision path: the CUDA results of configs 2-5 (BASELINE.json) checked through
the C ABI against sampled rows of the CPU oracle (oracle/sampled.py) - the full oracle result would cost minutes per
config on the host, a row costs one GEMV.  Sampled rows cover both ends of the triangle, a 128-row tile edge, interior
rows and the ragged last block (20000 = 156 * 128 + 32).

    C2  rs_cor_upper_triangle, Spearman, 20000 x 1000, |r|^6    src/rust/src/base/r_cors_similarity.rs:251-268
    C3  rs_differential_cor, 20000 genes, 500 vs 500 samples      src/rust/src/methods/r_diffcor.rs:45-74
    C4  calculate_tom_from_exp signed v1, 20000 x 1000, power 6   R/functions_stats.R:103-121, r_coremo.rs:46-59
    C5  sparse all-pairs Gram (extension), scaled to 262144 cells x 5000 genes
Tolerances (north star): 1e-12 exact-integer path, 1e-10 FP64-class paths, 1e-5 FP32-class paths.
"""
import numpy as np
import pytest

from oracle import c_binding as cb
from oracle import oracle as o
from oracle import sampled as sm

pytestmark = pytest.mark.gpu

P, N = 20000, 1000
BETA = 6.0


@pytest.fixture(scope="module")
def api(lib_built):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from bixverse_b200 import api as a

    a.init(1, 0)
    return a


def _prec(name):
    from bixverse_b200 import _lib

    return {"auto": _lib.PREC_AUTO, "f64": _lib.PREC_F64, "i8": _lib.PREC_I8EXACT, "tf32": _lib.PREC_TF32X3,
            "f16": _lib.PREC_F16X3, "ozaki": _lib.PREC_OZAKI}[name]


TOL = {"auto": 1e-12, "i8": 1e-12, "f64": 1e-10, "ozaki": 1e-10, "tf32": 1e-5, "f16": 1e-5}


@pytest.fixture(scope="module")
def c2_input():
    from bixverse_b200.synthetic import synthetic_bulk

    x = synthetic_bulk(N, P, seed=2024, tie_heavy=True)
    return x, sm.SampledOracle(x, True, ranks=cb.rank_matrix_col(x))


@pytest.fixture(scope="module")
def c4_input():
    from bixverse_b200.synthetic import synthetic_bulk

    x = synthetic_bulk(N, P, seed=2026)
    return x, sm.SampledOracle(x, False)


# ----------------------------------------------------------------------------- C2
@pytest.mark.parametrize("prec", ["auto", "i8", "tf32", "f16", "ozaki"])
def test_config2_spearman_power6_sampled_rows(api, c2_input, prec):
    """Config 2 on every tensor path (the FP64 DMMA path is test_gpu_parity.py::test_config2_full_size_properties).
    `auto` is what the drop-in patch passes; it must reach the exact-integer kernel and be bit-identical to `i8`."""
    x, so = c2_input
    got = api.rs_cor_upper_triangle(x, True, True, beta=BETA, precision=_prec(prec))
    assert got.shape[0] == P * (P - 1) // 2
    err = 0.0
    for i in sm.ROWS_20K:
        err = max(err, float(np.abs(sm.packed_row(got, 0, i, P) - so.cor_upper_row(i, 1, BETA)).max()))
    assert err <= TOL[prec], f"{prec}: max err {err}"
    assert np.isfinite(got[:: 4099]).all() and got.min() >= 0.0 and got.max() <= 1.0 + 1e-12
    if prec == "auto":
        ref = api.rs_cor_upper_triangle(x, True, True, beta=BETA, precision=_prec("i8"))
        assert np.array_equal(got, ref)


def test_config2_beta1_identity_rows(api, c2_input):
    """beta = 1 is the reference semantic (no soft threshold): sign kept, exact-integer path, FP64 parity."""
    x, so = c2_input
    got = api.rs_cor_upper_triangle(x, True, True, precision=_prec("auto"))
    for i in sm.ROWS_20K:
        assert np.abs(sm.packed_row(got, 0, i, P) - so.cor_upper_row(i)).max() <= 1e-12


# ----------------------------------------------------------------------------- C3
@pytest.mark.parametrize("spearman,prec", [(False, "f64"), (True, "f64"), (True, "auto"), (False, "tf32")])
def test_config3_diffcor_sampled_rows(api, spearman, prec):
    from bixverse_b200.synthetic import synthetic_pair

    xa, xb = synthetic_pair(500, 500, P, seed=2025)
    got = api.rs_differential_cor(xa, xb, spearman, precision=_prec(prec))
    oa = sm.SampledOracle(xa, spearman, ranks=cb.rank_matrix_col(xa) if spearman else None)
    ob = sm.SampledOracle(xb, spearman, ranks=cb.rank_matrix_col(xb) if spearman else None)
    tol_r = TOL[prec]
    for i in sm.ROWS_20K:
        ra, rb, z, pv = sm.diffcor_rows(oa, ob, i)
        g = {k: sm.packed_row(got[k], 0, i, P) for k in ("r_a", "r_b", "z_score", "p_val")}
        assert np.abs(g["r_a"] - ra).max() <= tol_r and np.abs(g["r_b"] - rb).max() <= tol_r
        if tol_r <= 1e-10:  # FP64-class: z and p follow (the usual bars of tests/test_gpu_parity.py::test_diffcor)
            assert np.abs(g["z_score"] - z).max() <= 1e-9
            assert np.allclose(g["p_val"], pv, rtol=1e-8, atol=1e-300)
        else:  # FP32-class r: dz = dr / (1 - r^2) / se
            se = np.sqrt(1.0 / 497 + 1.0 / 497)
            bound = 2.5 * tol_r / se / np.minimum(1 - ra * ra, 1 - rb * rb)
            assert (np.abs(g["z_score"] - z) <= bound).all()
    assert all(got[k].shape[0] == P * (P - 1) // 2 for k in got)
    m = api.UpperTriangleDiffcorMat(got, [f"g{i}" for i in range(P)])  # the result container takes it as is
    assert m.data["r_a"].shape[0] == P * (P - 1) // 2


# ----------------------------------------------------------------------------- C4
@pytest.mark.parametrize("prec", ["f64", "ozaki", "tf32", "f16", "auto"])
def test_config4_tom_signed_v1_power6_sampled_entries(api, c4_input, prec):
    x, so = c4_input
    got = api.calculate_tom_from_exp(x, True, "v1", "pearson", beta=BETA, precision=_prec(prec), packed=True)
    assert got.shape[0] == P * (P - 1) // 2
    tol = 1e-10 if prec == "auto" else TOL[prec]
    err = 0.0
    for i in sm.ROWS_20K:
        cols = sm.tom_sample_cols(i, P)
        ref = so.tom_entries(i, cols, True, "v1", BETA)
        err = max(err, float(np.abs(sm.packed_row(got, 0, i, P)[cols - i - 1] - ref).max()))
    assert err <= tol, f"{prec}: max err {err}"
    assert np.isfinite(got[:: 4099]).all()


def test_config4_tom_v2_unsigned_spearman_rows(api, c2_input):
    """The other TOM variant at full size on the reference semantic (beta = 1): unsigned v2 from Spearman correlations;
    AUTO = exact-integer correlation feeding an FP64 A * A."""
    x, so = c2_input
    got = api.calculate_tom_from_exp(x, False, "v2", "spearman", precision=_prec("auto"), packed=True)
    for i in (0, 128, 12345, 19998):
        cols = sm.tom_sample_cols(i, P)
        ref = so.tom_entries(i, cols, False, "v2", 1.0)
        assert np.abs(sm.packed_row(got, 0, i, P)[cols - i - 1] - ref).max() <= 1e-10


# ----------------------------------------------------------------------------- C5 (scaled)
@pytest.mark.parametrize("prec", ["f64", "tf32", "f16"])
def test_config5_sparse_gram_scaled_sampled_rows(api, prec):
    """Config 5 at a quarter of its cells (262 144 x 5 000, 250 non-zeros per cell = 6.6e7 non-zeros, 32 dense panels):
    rows of the all-pairs Pearson against the sparse-algebra restatement of oracle.sparse_gram_cor."""
    import scipy.sparse as sp

    from bixverse_b200.synthetic import synthetic_sparse_cells_fast

    cells, genes = 262144, 5000
    ip, ix, d = synthetic_sparse_cells_fast(cells, genes, 2027)
    got = api.sparse_gene_cor_upper_triangle(ip, ix, d, cells, genes, precision=_prec(prec))
    xs = sp.csr_matrix((d.astype(np.float64), ix, ip), shape=(cells, genes))
    xt = xs.T.tocsr()  # genes x cells
    s = np.asarray(xs.sum(axis=0)).ravel()
    ss = np.asarray(xs.multiply(xs).sum(axis=0)).ravel()
    var = ss - s * s / cells
    tol = 1e-9 if prec == "f64" else 1e-5
    for i in (0, 1, 127, 128, 2500, 4870, 4998):
        g = np.asarray((xt[i] @ xs).todense()).ravel()  # row i of X'X
        r = (g - s[i] * s / cells) / np.sqrt(var[i] * var)
        assert np.abs(sm.packed_row(got, 0, i, genes) - r[i + 1:]).max() <= tol
