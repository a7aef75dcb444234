"""Host -> host wall time of every drop-in entry point at BASELINE scale (20 000 genes x 1 000 samples unless stated),
one GPU, with the CPU oracle port timed beside it on a bounded sample (API_CPU=0 skips the CPU legs).
Writes gpurun_out/api_timings_r01.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bixverse_b200 import _lib, api
from bixverse_b200.synthetic import synthetic_bulk, synthetic_sparse_cells_fast

CPU = os.environ.get("API_CPU", "1") != "0"
if CPU:
    from oracle import oracle as o
    torch.set_num_threads(os.cpu_count() or 1)
api.init(1, 0)
n, p = 1000, 20000
P = p * (p - 1) // 2
x = synthetic_bulk(n, p, seed=2024, tie_heavy=True)
rows = []


def timed(fn, reps=3):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts[1:] if reps > 1 else ts)), r


def emit(name, ref, gpu_ms, cpu_ms=None, cpu_sample=None, note=None):
    row = {"api": name, "reference": ref, "gpu_host_to_host_ms": gpu_ms}
    if cpu_ms is not None:
        row["cpu_port_ms"] = cpu_ms
        row["cpu_sample"] = cpu_sample
    if note:
        row["note"] = note
    rows.append(row)
    print(json.dumps(row), flush=True)


ps = 4000  # CPU sample: first 4000 genes (cost scales with p^2; TOM with p^3)
xs = np.ascontiguousarray(x[:, :ps])
# rs_cor (dense p x p out, 3.2 GB)
t, _ = timed(lambda: api.rs_cor(x, True), 2)
c = timed(lambda: o.column_pairwise_cor(xs, True), 1)[0] if CPU else None
emit("rs_cor(spearman) dense 20000^2", "r_cors_similarity.rs:70-77", 1e3 * t, c and 1e3 * c, f"first {ps} genes")
# rs_cor_upper_triangle
t, packed = timed(lambda: api.rs_cor_upper_triangle(x, True, True, precision=_lib.PREC_I8EXACT))
c = timed(lambda: o.cor_upper_triangle(xs, True, True), 1)[0] if CPU else None
emit("rs_cor_upper_triangle(spearman)", "r_cors_similarity.rs:251-268", 1e3 * t, c and 1e3 * c, f"first {ps} genes")
# rs_differential_cor
xa, xb = x[:500], x[500:]
t, _ = timed(lambda: api.rs_differential_cor(xa, xb, False), 2)
c = timed(lambda: o.differential_cor(xa[:, :ps], xb[:, :ps], False), 1)[0] if CPU else None
emit("rs_differential_cor(pearson) 500 vs 500", "r_diffcor.rs:45-74", 1e3 * t, c and 1e3 * c, f"first {ps} genes")
# rs_rbf_iterate_epsilons on the packed distance vector
dist = 1.0 - np.abs(packed)
eps = np.array([0.5, 1.0, 1.5, 2.0, 2.5])
t, _ = timed(lambda: api.rs_rbf_iterate_epsilons(dist, eps, p, True, "gaussian"), 2)
ds = 1.0 - np.abs(o.cor_upper_triangle(xs, True, True)) if CPU else None
c = timed(lambda: o.rbf_iterate_epsilons(ds, eps, ps, True, "gaussian"), 1)[0] if CPU else None
emit("rs_rbf_iterate_epsilons, 5 epsilons", "r_rbf.rs:118-131", 1e3 * t, c and 1e3 * c, f"first {ps} genes")
# rs_coremo_stability: 4 left-out samples
idx = [1, 250, 500, 1000]
t, _ = timed(lambda: api.rs_coremo_stability(x, idx, 2.0, "gaussian", True, precision=_lib.PREC_I8EXACT), 2)
c = timed(lambda: o.coremo_stability(xs, idx[:1], 2.0, "gaussian", True), 1)[0] if CPU else None
emit("rs_coremo_stability, per left-out sample", "r_coremo.rs:191-235", 1e3 * t / len(idx), c and 1e3 * c, f"first {ps} genes, 1 sample")
# cor_module_tom fused
res = api.UpperTriangularSymMat(packed, [str(i) for i in range(p)], True)
t, _ = timed(lambda: api.cor_module_tom(res, True, "v2", precision=_lib.PREC_TF32X3), 2)
if CPU:
    dsq = o.upper_triangle_to_dense(o.cor_upper_triangle(xs, True, True), True, ps)
    c = timed(lambda: o.dense_to_upper_triangle(o.calc_tom(dsq, True, "v2"), True), 1)[0]
emit("cor_module_tom (packed in, packed out; 3xTF32)", "methods_coexp_cor.R:118-164", 1e3 * t, CPU and 1e3 * c or None, f"first {ps} genes")
del packed, dist, res
# rs_upper_triangle_to_sparse at 5000 genes (host index logic)
p5 = 5000
v5 = api.rs_cor_upper_triangle(x[:, :p5], False, True)
v5[np.abs(v5) < 0.3] = 0.0
t, _ = timed(lambda: api.rs_upper_triangle_to_sparse(v5, True, p5), 2)
emit("rs_upper_triangle_to_sparse, 5000 genes", "r_sparse.rs:43-55", 1e3 * t, note="host index logic, no GPU")
# pair-list sparse correlations: 200 000 cells x 5 000 genes, 20 000 pairs
cells, genes = 200000, 5000
ip, ix, dt = synthetic_sparse_cells_fast(cells, genes, 2027)
rng = np.random.default_rng(1)
g1 = rng.integers(0, 400, size=20000).astype(np.int32)
g2 = rng.integers(0, genes, size=20000).astype(np.int32)
sl = {"data": dt, "indptr": ip, "indices": ix, "nrow": cells, "ncol": genes, "format": "csr"}
for spearman in (False, True):
    t, _ = timed(lambda: api.rs_pairwise_gene_cors_mc(sl, g1, g2, spearman), 2)
    cfn = o.pairwise_gene_cors_spearman if spearman else o.pairwise_gene_cors
    cc, cg = 20000, genes
    sub_ip = ip[: cc + 1]
    c = timed(lambda: cfn(sub_ip, ix[: sub_ip[-1]], dt[: sub_ip[-1]], cc, cg, g1[:2000], g2[:2000]), 1)[0] if CPU else None
    emit(f"rs_pairwise_gene_cors_mc({'spearman' if spearman else 'pearson'}) 200k cells, 20k pairs", "r_mc_processing.rs:258-283",
         1e3 * t, c and 1e3 * c, "20 000 cells, 2 000 pairs")
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/api_timings_r01.json", "w") as f:
    for r in rows:
        f.write(json.dumps(r) + "\n")
