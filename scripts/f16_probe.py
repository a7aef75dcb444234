"""TOM (20 000 genes, signed v1, beta 6 and beta 1) on the split-float kinds: time and deviation from the FP64 DMMA path."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bixverse_b200 import _lib, api
from bixverse_b200 import distributed as D
from bixverse_b200.synthetic import synthetic_bulk
api.init(1, 0); lib = _lib.load()
dev = torch.device("cuda", 0); be = D.DeviceBackend(dev)
n, p = 1000, 20000
x = synthetic_bulk(n, p, seed=2026)
xd = torch.from_numpy(np.ascontiguousarray(x.T)).to(dev)
kinds = [k for k in (sys.argv[1:] or ["tf32", "f16"])]
PREC = {"tf32": _lib.PREC_TF32X3, "f16": _lib.PREC_F16X3}
for beta in (6.0, 1.0):
    ref, _ = D.sharded_tom_from_exp(be, xd, n, p, False, True, "v1", beta=beta, precision=_lib.PREC_F64)
    for name in kinds:
        prec = PREC[name]
        D.sharded_tom_from_exp(be, xd, n, p, False, True, "v1", beta=beta, precision=prec)
        _lib.check(lib.bixcor_timing_enable(1))
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out, _ = D.sharded_tom_from_exp(be, xd, n, p, False, True, "v1", beta=beta, precision=prec)
        torch.cuda.synchronize(); t = time.perf_counter() - t0
        tm = _lib.last_timing(); _lib.check(lib.bixcor_timing_enable(0))
        print(json.dumps({"tom": name, "beta": beta, "lib": os.path.basename(_lib.LIB_PATH), "ctas": os.environ.get("BIXCOR_UMMA_CTAS", "2"),
                          "wall_ms": 1e3 * t, "gemm_ms": tm["gemm_ms"], "max_dev_vs_f64": (ref - out).abs().max().item()}), flush=True)
    del ref
