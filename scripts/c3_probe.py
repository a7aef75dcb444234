import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from bixverse_b200 import _lib, api
from bixverse_b200 import distributed as D
from bixverse_b200.synthetic import synthetic_pair, synthetic_bulk
api.init(1, 0); lib = _lib.load()
dev = torch.device("cuda", 0); be = D.DeviceBackend(dev)
p = 20000
xa, xb = synthetic_pair(500, 500, p, seed=2025)
xa = torch.from_numpy(np.ascontiguousarray(xa.T)).to(dev); xb = torch.from_numpy(np.ascontiguousarray(xb.T)).to(dev)
for name, prec, sp in (("f64-pearson", _lib.PREC_F64, False), ("i8-spearman", _lib.PREC_I8EXACT, True), ("tf32-pearson", _lib.PREC_TF32X3, False), ("ozaki-pearson", _lib.PREC_OZAKI, False)):
    D.sharded_diffcor(be, xa, 500, xb, 500, p, sp, prec)
    _lib.check(lib.bixcor_timing_enable(1)); torch.cuda.synchronize(); t0 = time.perf_counter()
    r = D.sharded_diffcor(be, xa, 500, xb, 500, p, sp, prec)
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    print("C3", name, "wall ms %.2f" % (1e3 * t), _lib.last_timing()); _lib.check(lib.bixcor_timing_enable(0)); del r
