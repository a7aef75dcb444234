"""C-ABI surface checks that need no GPU: the library loads, exports every symbol
include/bixcor.h declares, host-only helpers work, compute fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "bixcor.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bixcor_\w+)\s*\(", src)))


def test_header_symbols_all_exported(lib_built):
    from bixverse_b200 import _lib

    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bixcor.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes signature table out of sync with the header"


def test_packed_helpers_match_oracle(lib_built):
    from bixverse_b200 import api, _lib

    lib = _lib.load()
    for p in (1, 2, 7, 130):
        for shift in (0, 1):
            assert lib.bixcor_packed_len(p, shift) == o.upper_triangle_indices(p, shift)[0].shape[0]
            for i in range(p):
                assert lib.bixcor_packed_offset(i, p, shift) == int(o.packed_offset(i, p, shift))
    rng = np.random.default_rng(2)
    m = rng.normal(size=(9, 9))
    m = m + m.T
    for shift in (True, False):
        v = api.rs_dense_to_upper_triangle(m, shift)
        assert np.array_equal(v, o.dense_to_upper_triangle(m, shift))
        d = api.rs_upper_triangle_to_dense(v, shift, 9)
        assert np.array_equal(d, o.upper_triangle_to_dense(v, shift, 9))
    with pytest.raises(ValueError):
        api.rs_upper_triangle_to_dense(np.zeros(5), True, 9)
    # p >= 2048 takes the multi-threaded row-block path of the host helpers (pure index logic, no GPU involved)
    p = 2100
    for shift in (True, False):
        v = rng.normal(size=lib.bixcor_packed_len(p, int(shift)))
        d = api.rs_upper_triangle_to_dense(v, shift, p)
        assert np.array_equal(d, o.upper_triangle_to_dense(v, shift, p))
        assert np.array_equal(api.rs_dense_to_upper_triangle(d, shift), v)


def test_reference_sparse_vector(lib_built, golden_dir):
    import json

    from bixverse_b200 import api

    kat = json.load(open(os.path.join(golden_dir, "reference_kat.json")))
    d = api.rs_upper_triangle_to_dense(kat["sparse_input_packed"], True, 4)
    assert np.array_equal(d, np.array(kat["sparse_dense_expected"]))


def test_upper_triangle_to_sparse_matches_reference_test(lib_built):
    """inst/tinytest/test_sparse.R:5-50: the packed vector c(0.3, 0, 1, 0, 0.2, -0.5), n = 4, shift TRUE as CSC equals
    Matrix's sparse form of rs_upper_triangle_to_dense(...) (scipy.sparse is the Matrix stand-in)."""
    import scipy.sparse as sp

    from bixverse_b200 import api

    data = np.array([0.3, 0.0, 1.0, 0.0, 0.2, -0.5])
    got = api.rs_upper_triangle_to_sparse(data, True, 4, "csc")
    dense = api.rs_upper_triangle_to_dense(data, True, 4)
    ref = sp.csc_matrix(dense)
    ref.sort_indices()
    assert got["nrow"] == 4 and got["ncol"] == 4 and got["cs_type"] == "csc"
    assert np.array_equal(got["indptr"], ref.indptr) and np.array_equal(got["indices"], ref.indices)
    assert np.array_equal(got["data"], ref.data)
    assert list(got["indptr"]) == [0, 3, 6, 8, 12]
    rng = np.random.default_rng(0)
    for p, shift in ((1, True), (7, True), (9, False), (40, True), (2300, True), (2100, False)):  # >= 2048: threaded passes
        v = rng.normal(size=p * (p - 1) // 2 if shift else p * (p + 1) // 2)
        v[rng.uniform(size=v.shape) < 0.4] = 0.0
        got = api.rs_upper_triangle_to_sparse(v, shift, p, "csr")
        ref = sp.csc_matrix(o.upper_triangle_to_dense(v, shift, p))
        ref.sort_indices()
        assert np.array_equal(got["indptr"], ref.indptr) and np.array_equal(got["indices"], ref.indices)
        assert np.array_equal(got["data"], ref.data)
    with pytest.raises(ValueError):
        api.rs_upper_triangle_to_sparse(data, True, 4, "coo")


def test_shard_rows_matches_oracle(lib_built):
    from bixverse_b200 import api

    for p in (100, 2000, 20000, 20011):
        for shift in (False, True):
            for world in (1, 2, 4, 8):
                for r in range(world):
                    assert api.shard_rows(p, shift, world, r) == o.shard_rows(p, int(shift), world, r)


def test_no_cpu_fallback(lib_built):
    """Without a GPU every compute entry point must fail with BIXCOR_ERR_NO_GPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from bixverse_b200 import api

    x = np.random.default_rng(0).normal(size=(8, 5))
    for call in (lambda: api.rs_cor(x, False), lambda: api.rs_cor_upper_triangle(x, True),
                 lambda: api.rs_tom(np.eye(4), "v1", True), lambda: api.rs_rank_matrix_col(x)):
        with pytest.raises(api.BixcorError) as ei:
            call()
        assert ei.value.code == -1 and "no CPU fallback" in ei.value.message


def test_host_side_argument_errors(lib_built):
    from bixverse_b200 import api

    with pytest.raises(ValueError, match="same number of columns"):
        api.rs_differential_cor(np.zeros((5, 3)), np.zeros((5, 4)), False)
    with pytest.raises(ValueError, match="Invalid TOM version type: v3"):
        api.rs_tom(np.eye(3), "v3", True)
    with pytest.raises(ValueError):
        api.UpperTriangularSymMat(np.zeros(4), ["a", "b", "c"], shift=True)
    m = api.UpperTriangularSymMat(np.array([0.1, 0.2, 0.3]), ["a", "b", "c"], shift=True)
    assert m.get_sym_matrix()[0, 2] == 0.2 and m.get_sym_matrix()[1, 1] == 1.0
    with pytest.raises(ValueError, match="missing"):
        api.UpperTriangleDiffcorMat({"r_a": [0.0]}, ["a", "b"])
    d = api.UpperTriangleDiffcorMat({"r_a": [0.1], "r_b": [0.2], "z_score": [1.0], "p_val": [0.0]}, ["a", "b"])
    assert d.get_data_table()["r_a"].shape[0] == 0  # p_val == 0 rows dropped


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "bixverse_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "liboracle" not in txt, f


def test_exception_barrier_stops_injected_failures(lib_built):
    """include/bixcor.h: no C++ exception crosses the ABI.  bixcor_debug_inject makes the NEXT entry point throw from
    inside its body; the barrier turns std::bad_alloc into BIXCOR_ERR_OOM and anything else into BIXCOR_ERR_INTERNAL,
    and the library keeps working afterwards.  (Host-only entry point, so this runs without a GPU.)"""
    from bixverse_b200 import _lib

    lib = _lib.load()
    packed = np.array([0.3, 0.0, 1.0, 0.0, 0.2, -0.5])
    out = np.empty((4, 4), order="F")
    for what, code, text in ((1, _lib.ERR_OOM, b"std::bad_alloc"), (2, _lib.ERR_INTERNAL, b"injected failure")):
        assert lib.bixcor_debug_inject(what) == 0
        rc = lib.bixcor_upper_to_dense(packed.ctypes.data, 1, 4, out.ctypes.data)
        assert rc == code, rc
        msg = lib.bixcor_last_error()
        assert text in msg and b"bixcor_upper_to_dense" in msg
        assert lib.bixcor_upper_to_dense(packed.ctypes.data, 1, 4, out.ctypes.data) == 0  # one-shot: the next call is clean
        assert out[0, 1] == 0.3 and out[3, 3] == 1.0
    lib.bixcor_debug_inject(1)
    assert lib.bixcor_upper_to_sparse_nnz(packed.ctypes.data, 1, 4) == -1  # count-returning entry point: -1 on failure


def test_small_p_shards_are_empty_not_invalid(lib_built):
    """for 99 <= p <= 127 and 5-8 ranks the 128-aligned boundaries clamp to p and a rank owns (p, p)."""
    from bixverse_b200 import api

    for p in (99, 120, 127):
        ranges = [api.shard_rows(p, True, 8, r) for r in range(8)]
        assert ranges[0][0] == 0 and ranges[-1][1] == p
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        assert sum(1 for r0, r1 in ranges if r1 > r0) == 1  # one rank owns everything, the others are empty


def test_host_halves_of_the_transports(lib_built):
    """What the staging threads do to a piece of a result (no GPU involved): the exact Spearman Gram arrives as int32 and is
    completed to r = S / (|d_i| |d_j|) (and |r|^6) with the device epilogue's IEEE multiplications in its order - so the
    values must equal the same chain evaluated in numpy bit for bit, for any piece of the packed vector (ragged starts / ends,
    pieces spanning many rows, unaligned destinations that take the scalar head / tail of the AVX2 loop), both shifts; the
    binary32 transport widens float -> double exactly."""
    import ctypes as C

    from bixverse_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(7)
    for p, shift in ((37, 1), (130, 1), (64, 0), (301, 1)):
        total = lib.bixcor_packed_len(p, shift)
        s_all = rng.integers(-(2 ** 31) + 1, 2 ** 31 - 1, size=total, dtype=np.int64).astype(np.int32)
        inv = 1.0 / np.sqrt(rng.uniform(1.0, 3.0e8, size=p))
        inv[3] = np.nan  # a constant / NaN gene: its 1/|d| is NaN and poisons its row and column
        # reference chain: ((double)S * inv_i) * inv_j, then sq = v * v, (sq * sq) * sq
        ii, jj = np.triu_indices(p, 1 if shift else 0)
        v = (s_all.astype(np.float64) * inv[ii]) * inv[jj]
        sq = v * v
        want = {0: v, 1: (sq * sq) * sq}
        for pow6 in (0, 1):
            for g0, count in ((0, total), (1, total - 1), (5, min(total - 5, 77)), (total // 3 + 1, total // 2), (total - 1, 1), (2, 0)):
                buf = np.full(count + 3, -123.0)
                for off in (0, 1):  # off = 1: the destination is 8 (mod 16/32): scalar head of the vector loop
                    dst = buf[off:off + count]
                    src = np.ascontiguousarray(s_all[g0:g0 + count])
                    _lib.check(lib.bixcor_host_expand_s32(src.ctypes.data, count, g0, inv.ctypes.data, p, shift, pow6, dst.ctypes.data))
                    assert np.array_equal(dst, want[pow6][g0:g0 + count], equal_nan=True), (p, shift, pow6, g0, count, off)
                    assert buf[off + count] == -123.0  # nothing written past the piece
        assert lib.bixcor_host_expand_s32(s_all.ctypes.data, total, 1, inv.ctypes.data, p, shift, 0, np.empty(total).ctypes.data) == _lib.ERR_INVALID
    f = np.concatenate([rng.normal(size=1001).astype(np.float32), np.float32([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, 3.4e38])])
    for off in (0, 1, 3):
        out = np.empty(f.size + 4)
        _lib.check(lib.bixcor_host_widen_f32(f.ctypes.data, f.size, out[off:off + f.size].ctypes.data))
        assert np.array_equal(out[off:off + f.size], f.astype(np.float64), equal_nan=True)
        assert np.array_equal(np.signbit(out[off:off + f.size]), np.signbit(f))
