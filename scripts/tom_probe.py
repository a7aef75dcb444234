import sys, os, json, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from bixverse_b200 import _lib, api
from bixverse_b200 import distributed as D
from bixverse_b200.synthetic import synthetic_bulk
api.init(1, 0); lib = _lib.load()
dev = torch.device("cuda", 0); be = D.DeviceBackend(dev)
n, p = 1000, 20000
x = synthetic_bulk(n, p, seed=2026)
xd = torch.from_numpy(np.ascontiguousarray(x.T)).to(dev)
for name, prec in (("tf32", _lib.PREC_TF32X3),):
    D.sharded_tom_from_exp(be, xd, n, p, False, True, "v1", beta=6.0, precision=prec)
    _lib.check(lib.bixcor_timing_enable(1))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out, _ = D.sharded_tom_from_exp(be, xd, n, p, False, True, "v1", beta=6.0, precision=prec)
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    tm = _lib.last_timing(); _lib.check(lib.bixcor_timing_enable(0))
    print("TOM", name, "wall ms", 1e3 * t, tm)
# accuracy of the tf32 Gram on config 2 rows against float64 numpy
from oracle import oracle as o
got = api.rs_cor_upper_triangle(x, False, True, precision=_lib.PREC_TF32X3)
z = x - x.mean(axis=0); z /= np.sqrt((z * z).sum(axis=0))
err = 0.0
for i in (0, 127, 5000, 19871):
    ref = z[:, i] @ z[:, i + 1:]
    off = int(o.packed_offset(i, p, 1))
    err = max(err, float(np.abs(got[off: off + p - 1 - i] - ref).max()))
print("tf32 pearson 20000x1000 max err sampled rows", err)
