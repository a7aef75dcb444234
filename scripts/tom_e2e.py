"""Host -> host wall time of bixcor_tom_from_exp (config 4, packed, pageable and page-locked buffers)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bixverse_b200 import _lib, api
from bixverse_b200.synthetic import synthetic_bulk
lib = _lib.load(); api.init(1, 0)
n, p = 1000, 20000
x = np.ascontiguousarray(synthetic_bulk(n, p, seed=2026).T)
out = np.zeros(p * (p - 1) // 2)
xp, op_ = torch.from_numpy(x).pin_memory(), torch.empty(p * (p - 1) // 2, dtype=torch.float64).pin_memory()
kinds = sys.argv[1:] or ["f16", "tf32", "ozaki", "f64"]
PREC = {"f64": _lib.PREC_F64, "tf32": _lib.PREC_TF32X3, "f16": _lib.PREC_F16X3, "ozaki": _lib.PREC_OZAKI}
for name in kinds:
    for label, xi, oi in (("pageable", x.ctypes.data, out.ctypes.data), ("pinned", xp.data_ptr(), op_.data_ptr())):
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); _lib.check(lib.bixcor_tom_from_exp(xi, n, p, 0, 1, 1, 6.0, PREC[name], 1, oi)); ts.append(time.perf_counter() - t0)
        print(name, label, "ms", [round(1e3 * t, 1) for t in ts], flush=True)
