"""Runs the five BASELINE.json configs at full size through the host C ABI on one B200, checks sampled
entries against the CPU oracle and records wall times (pinned host buffers are NOT used here: this is the
plain pageable-memory path an R caller would take).  Output: one JSON object per line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from bixverse_b200 import _lib, api
from bixverse_b200.synthetic import synthetic_bulk, synthetic_pair, synthetic_sparse_cells_fast
from oracle import oracle as o

api.init(1, 0)
OUT = []


def emit(d):
    OUT.append(d)
    print(json.dumps(d), flush=True)


def timed(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t0)
    return r, float(np.median(ts))


def zmat(x, spearman):
    r = o.rank_matrix_col(x) if spearman else np.asarray(x)
    z = r - r.mean(axis=0)
    return z / np.sqrt((z * z).sum(axis=0))


ROWS = (0, 1, 127, 128, 5000, 12345, 19871, 19998)

# ---------------------------------------------------------------- C1
x = synthetic_bulk(200, 2000, seed=123)
got, t = timed(lambda: api.rs_cor_upper_triangle(x, False, True), 5)
ref = o.cor_upper_triangle(x, False, True)
t0 = time.perf_counter(); o.cor_upper_triangle(x, False, True); tc = time.perf_counter() - t0
emit({"config": "C1 Pearson 2000x200 packed", "pairs": int(got.shape[0]), "wall_ms": 1e3 * t, "pairs_per_s": got.shape[0] / t,
      "max_err": float(np.abs(got - ref).max()), "cpu_oracle_ms": 1e3 * tc})

# ---------------------------------------------------------------- C2
n, p = 1000, 20000
x = synthetic_bulk(n, p, seed=2024, tie_heavy=True)
z = zmat(x, True)
for name, prec in (("f64", _lib.PREC_F64), ("i8", _lib.PREC_I8EXACT), ("tf32", _lib.PREC_TF32X3)):
    got, t = timed(lambda: api.rs_cor_upper_triangle(x, True, True, beta=6.0, precision=prec), 3)
    err = 0.0
    for i in ROWS:
        refrow = np.abs(z[:, i] @ z[:, i + 1:]) ** 6.0
        off = int(o.packed_offset(i, p, 1))
        err = max(err, float(np.abs(got[off: off + p - 1 - i] - refrow).max()))
    emit({"config": "C2 Spearman 20000x1000 |r|^6 packed", "precision": name, "pairs": int(got.shape[0]), "wall_ms": 1e3 * t,
          "pairs_per_s": got.shape[0] / t, "max_err_sampled_rows": err, "host": "pageable"})
ranks = api.rs_rank_matrix_col(x[:, :2000])
emit({"config": "C2 ranks bit-exact (first 2000 genes)", "equal": bool(np.array_equal(ranks, o.rank_matrix_col(x[:, :2000])))})
del got

# ---------------------------------------------------------------- C3
xa, xb = synthetic_pair(500, 500, 20000, seed=2025)
for name, prec, sp in (("f64-pearson", _lib.PREC_F64, False), ("i8-spearman", _lib.PREC_I8EXACT, True), ("f64-spearman", _lib.PREC_F64, True)):
    res, t = timed(lambda: api.rs_differential_cor(xa, xb, sp, precision=prec), 2)
    za, zb = zmat(xa, sp), zmat(xb, sp)
    c = 1.06 if sp else 1.0
    se = np.sqrt(c / 497 + c / 497)
    e_r = e_z = e_p = 0.0
    from scipy.special import erfc
    for i in ROWS:
        ra, rb = za[:, i] @ za[:, i + 1:], zb[:, i] @ zb[:, i + 1:]
        zz = (np.arctanh(ra) - np.arctanh(rb)) / se
        pv = erfc(np.abs(zz) / np.sqrt(2))
        off = int(o.packed_offset(i, p, 1)); sl = slice(off, off + p - 1 - i)
        e_r = max(e_r, float(np.abs(res["r_a"][sl] - ra).max()), float(np.abs(res["r_b"][sl] - rb).max()))
        e_z = max(e_z, float(np.abs(res["z_score"][sl] - zz).max()))
        e_p = max(e_p, float(np.max(np.abs(res["p_val"][sl] - pv) / np.maximum(pv, 1e-300))))
    emit({"config": "C3 diffcor 20000 genes, 500 vs 500", "precision": name, "pairs": int(res["r_a"].shape[0]), "wall_ms": 1e3 * t,
          "pairs_per_s": res["r_a"].shape[0] / t, "max_err_r": e_r, "max_err_z": e_z, "max_relerr_p": e_p})
del res

# ---------------------------------------------------------------- C4
x = synthetic_bulk(1000, 20000, seed=2026)
t0 = time.perf_counter()
r = o.column_pairwise_cor(x, False)
a = o.soft_threshold(r, 6.0, signed=True)
np.fill_diagonal(a, 1.0)
k = np.abs(a).sum(axis=1)
t_cpu_cor = time.perf_counter() - t0
for name, prec in (("f64", _lib.PREC_F64), ("ozaki-int8", _lib.PREC_OZAKI), ("tf32", _lib.PREC_TF32X3)):
    for packed in (True, False):
        got, t = timed(lambda: api.calculate_tom_from_exp(x, True, "v1", "pearson", beta=6.0, precision=prec, packed=packed), 2)
        err = 0.0
        for i in ROWS:
            shared = a[i] @ a - a[i] * (a[i, i] + np.diag(a))
            tom = (a[i] + shared) / (np.minimum(k[i], k) + 1.0 - np.abs(a[i]))
            if packed:
                off = int(o.packed_offset(i, p, 1))
                err = max(err, float(np.abs(got[off: off + p - 1 - i] - tom[i + 1:]).max()))
            else:
                tom[i] = 1.0
                err = max(err, float(np.abs(got[i] - tom).max()), float(np.abs(got[:, i] - tom).max()))
        emit({"config": "C4 TOM-from-exp signed v1 beta=6, 20000x1000", "precision": name, "output": "packed" if packed else "dense",
              "wall_ms": 1e3 * t, "alg_tflops": 8.4e12 / t / 1e12, "max_err_sampled_rows": err})
        del got
emit({"config": "C4 cpu oracle correlation+adjacency only (no A*A)", "wall_ms": 1e3 * t_cpu_cor})
del r, a

# ---------------------------------------------------------------- C5
cells, genes = 1_000_000, 5000
t0 = time.perf_counter()
ip, ix, d = synthetic_sparse_cells_fast(cells, genes, seed=2027)
tgen = time.perf_counter() - t0
# oracle on the first 250 genes (dense 1M x 250)
gsub = 250
sel = ix < gsub
rows = np.repeat(np.arange(cells, dtype=np.int64), np.diff(ip))[sel]
dense = np.zeros((cells, gsub))
dense[rows, ix[sel]] = d[sel].astype(np.float64)
ref = np.corrcoef(dense.T)
for pname, prec in (("f64 (dense-panel DMMA)", _lib.PREC_F64), ("tf32 (dense-panel 3xTF32 tcgen05)", _lib.PREC_TF32X3)):
    got, t = timed(lambda: api.sparse_gene_cor_upper_triangle(ip, ix, d, cells, genes, precision=prec), 1)
    err = 0.0
    for i in range(gsub - 1):
        off = int(o.packed_offset(i, genes, 1))
        err = max(err, float(np.nanmax(np.abs(got[off: off + gsub - 1 - i] - ref[i, i + 1: gsub]))))
    emit({"config": "C5 sparse CSR 1M cells x 5000 genes (5% density)", "nnz": int(ip[-1]), "pairs": int(got.shape[0]), "wall_ms": 1e3 * t,
          "pairs_per_s": got.shape[0] / t, "max_err_first_250_genes": err, "gen_s": tgen, "precision": pname})

with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "configs_r01.json"), "w") as f:
    for d_ in OUT:
        f.write(json.dumps(d_) + "\n")
