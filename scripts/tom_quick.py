"""Quick 20k TOM timing (HBM resident, packed) + sampled parity for the split-float kinds; prints kernel ms from the library timers."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bixverse_b200 import _lib, api
from bixverse_b200 import distributed as D
from bixverse_b200.synthetic import synthetic_bulk
from oracle import sampled as sm

kinds = sys.argv[1:] or ["f16", "tf32"]
PREC = {"f64": _lib.PREC_F64, "tf32": _lib.PREC_TF32X3, "f16": _lib.PREC_F16X3, "ozaki": _lib.PREC_OZAKI}
api.init(1, 0)
lib = _lib.load()
dev = torch.device("cuda", 0)
be = D.DeviceBackend(dev)
n, p = 1000, 20000
x = synthetic_bulk(n, p, seed=2026)
xd = torch.from_numpy(np.ascontiguousarray(x.T)).to(dev)
so = sm.SampledOracle(x, False)
for k in kinds:
    out, _ = D.sharded_tom_from_exp(be, xd, n, p, False, True, "v1", beta=6.0, precision=PREC[k])
    torch.cuda.synchronize()
    _lib.check(lib.bixcor_timing_enable(1))
    t0 = time.perf_counter()
    out, _ = D.sharded_tom_from_exp(be, xd, n, p, False, True, "v1", beta=6.0, precision=PREC[k])
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    tm = _lib.last_timing()
    _lib.check(lib.bixcor_timing_enable(0))
    err = 0.0
    for i in sm.ROWS_20K:
        cols = sm.tom_sample_cols(i, p)
        a = D.packed_offset(i, p, 1)
        got = out[a:a + (p - i - 1)].cpu().numpy()[cols - i - 1]
        err = max(err, float(np.abs(got - so.tom_entries(i, cols, True, "v1", 6.0)).max()))
    print(k, "wall ms %.2f" % (1e3 * wall), tm, "max err %.3g" % err, flush=True)
