#!/bin/bash
# ad-hoc round-2 check: TOM host->host, sparse tests, streamed C5
mkdir -p gpurun_out/r2i
python scripts/tom_e2e.py f16 tf32 2>&1 | tee gpurun_out/r2i/tom_e2e.log
python -m pytest tests -m gpu -q -k "sparse or config5" 2>&1 | tail -2
python - <<PY 2>&1 | tee gpurun_out/r2i/c5.log
import sys, time, numpy as np
sys.path.insert(0, ".")
from bixverse_b200 import _lib, api
from bixverse_b200.synthetic import synthetic_sparse_cells_fast
lib = _lib.load(); api.init(1, 0)
cells, genes = 1000000, 5000
ip, ix, dd = synthetic_sparse_cells_fast(cells, genes, 2027)
out = np.zeros(genes * (genes - 1) // 2)
for name, prec in (("f16", _lib.PREC_F16X3), ("tf32", _lib.PREC_TF32X3)):
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); _lib.check(lib.bixcor_sparse_gram_cor(ip.ctypes.data, ix.ctypes.data, dd.ctypes.data, cells, genes, prec, out.ctypes.data)); ts.append(time.perf_counter() - t0)
    print(name, "C5 1M x 5k host->host ms", [round(1e3 * t, 1) for t in ts], float(out[:100000].sum()))
PY
