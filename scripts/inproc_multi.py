"""Single-process multi-GPU mode (bixcor_init(n): what an R session linking libbixcor would use): host -> host
wall time of the packed correlation (config 2) and the differential correlation (config 3) on 1..N devices."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bixverse_b200 import _lib, api
from bixverse_b200.synthetic import synthetic_bulk, synthetic_pair

lib = _lib.load()
ndev_max = torch.cuda.device_count()
n, p = 1000, 20000
P = p * (p - 1) // 2
x = torch.from_numpy(np.ascontiguousarray(synthetic_bulk(n, p, seed=2024, tie_heavy=True).T)).pin_memory()
out = torch.empty(P, dtype=torch.float64).pin_memory()
xa, xb = synthetic_pair(500, 500, p, seed=2025)
xa = torch.from_numpy(np.ascontiguousarray(xa.T)).pin_memory()
xb = torch.from_numpy(np.ascontiguousarray(xb.T)).pin_memory()
outs = [torch.empty(P, dtype=torch.float64).pin_memory() for _ in range(4)]
rows = []
for nd in [d for d in (1, 2, 4, 8) if d <= ndev_max]:
    api.init(nd, 0)
    for name, prec in (("i8", _lib.PREC_I8EXACT), ("f64", _lib.PREC_F64)):
        ts = []
        for rep in range(4):
            t0 = time.perf_counter()
            _lib.check(lib.bixcor_cor_upper(x.data_ptr(), n, p, 1, 1, 6.0, prec, out.data_ptr()))
            ts.append(time.perf_counter() - t0)
        t = float(np.median(ts[1:]))
        rows.append({"config": "C2 host->host, one process", "devices": nd, "precision": name, "ms": 1e3 * t, "pairs_per_s": P / t,
                     "checksum": float(out[:1000000].sum())})
        print(json.dumps(rows[-1]), flush=True)
    ts = []
    for rep in range(3):
        t0 = time.perf_counter()
        _lib.check(lib.bixcor_diffcor(xa.data_ptr(), 500, xb.data_ptr(), 500, p, 0, _lib.PREC_F64, *[o_.data_ptr() for o_ in outs]))
        ts.append(time.perf_counter() - t0)
    t = float(np.median(ts[1:]))
    rows.append({"config": "C3 diffcor host->host, one process", "devices": nd, "precision": "f64", "ms": 1e3 * t, "pairs_per_s": P / t})
    print(json.dumps(rows[-1]), flush=True)
# multi-device TOM from a host matrix (packed result on the host): FP64 DMMA / Ozaki int8 / 3xTF32
tom_out = torch.empty(P, dtype=torch.float64).pin_memory()
for nd in [d for d in (1, 2, 4, 8) if d <= ndev_max]:
    api.init(nd, 0)
    for name, prec in (("f16", _lib.PREC_F16X3), ("tf32", _lib.PREC_TF32X3), ("ozaki", _lib.PREC_OZAKI), ("f64", _lib.PREC_F64)):
        ts = []
        for rep in range(3):
            t0 = time.perf_counter()
            _lib.check(lib.bixcor_tom_from_exp(x.data_ptr(), n, p, 0, 1, _lib.TOM_V1, 6.0, prec, 1, tom_out.data_ptr()))
            ts.append(time.perf_counter() - t0)
        t = float(np.median(ts[1:]))
        rows.append({"config": "C4 TOM-from-exp host->host, one process", "devices": nd, "precision": name, "ms": 1e3 * t,
                     "checksum": float(tom_out[:1000000].sum())})
        print(json.dumps(rows[-1]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/inproc_multi_r01.json", "w") as f:
    for r in rows:
        f.write(json.dumps(r) + "\n")
