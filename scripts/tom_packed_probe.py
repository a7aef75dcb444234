"""cor_module_tom (R/methods_coexp_cor.R:118-164) at 20 000 genes: the fused bixcor_tom_packed call against the
three-call chain the R method makes (rs_upper_triangle_to_dense -> rs_tom -> rs_dense_to_upper_triangle)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bixverse_b200 import _lib, api
from bixverse_b200.synthetic import synthetic_bulk

lib = _lib.load()
api.init(1, 0)
n, p = 1000, 20000
P = p * (p - 1) // 2
x = synthetic_bulk(n, p, seed=2024)
packed = torch.from_numpy(api.rs_cor_upper_triangle(x, False, True)).pin_memory()
out = torch.empty(P, dtype=torch.float64).pin_memory()
rows = []
for name, prec in (("f16", _lib.PREC_F16X3), ("tf32", _lib.PREC_TF32X3), ("ozaki", _lib.PREC_OZAKI), ("f64", _lib.PREC_F64)):
    ts = []
    for rep in range(3):
        t0 = time.perf_counter()
        _lib.check(lib.bixcor_tom_packed(packed.data_ptr(), 1, p, 1, _lib.TOM_V2, prec, out.data_ptr()))
        ts.append(time.perf_counter() - t0)
    rows.append({"config": "cor_module_tom fused (packed in -> packed out), 20000 genes, signed v2", "precision": name,
                 "ms": 1e3 * float(np.median(ts[1:])), "checksum": float(out[:1000000].sum())})
    print(json.dumps(rows[-1]), flush=True)
fused = out.numpy().copy()
t0 = time.perf_counter()
dense = api.rs_upper_triangle_to_dense(packed.numpy(), True, p)
t1 = time.perf_counter()
tom = api.rs_tom(dense, "v2", True)
t2 = time.perf_counter()
back = api.rs_dense_to_upper_triangle(tom, True)
t3 = time.perf_counter()
rows.append({"config": "cor_module_tom as three calls (host unpack, rs_tom f64 dense -> dense, host repack)", "precision": "f64",
             "unpack_ms": 1e3 * (t1 - t0), "tom_ms": 1e3 * (t2 - t1), "repack_ms": 1e3 * (t3 - t2), "ms": 1e3 * (t3 - t0),
             "max_abs_diff_vs_fused": float(np.max(np.abs(back - fused)))})
print(json.dumps(rows[-1]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/tom_packed_r01.json", "w") as f:
    for r in rows:
        f.write(json.dumps(r) + "\n")
