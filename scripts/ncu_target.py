"""One HBM-resident invocation of a named path, for ncu captures of its kernels (profiles/README_r02.md):

  python scripts/ncu_target.py c2 <auto|f64|tf32|f16|ozaki>     config 2: Spearman 20000 x 1000, |r|^6, packed
  python scripts/ncu_target.py tom <f16|tf32|ozaki|f64>         config 4: signed TOM v1 from expression, beta 6, packed
  python scripts/ncu_target.py c3 <auto|f64>                    config 3: differential correlation 20000 x (500 vs 500)
  python scripts/ncu_target.py c2host <auto|f16>                config 2 host -> host (page-locked buffers): the int32-transport
                                                                instantiation of the Gram kernel (auto), 20 band launches per call

Two warm-up invocations, then one measured; `-k regex:<kernel> -s <launches of the warm-ups>` selects the last one."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bixverse_b200 import _lib, api
from bixverse_b200 import distributed as D
from bixverse_b200.synthetic import synthetic_bulk, synthetic_pair

what, pname = sys.argv[1], sys.argv[2]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
PREC = {"auto": _lib.PREC_AUTO, "f64": _lib.PREC_F64, "i8": _lib.PREC_I8EXACT, "tf32": _lib.PREC_TF32X3, "f16": _lib.PREC_F16X3,
        "ozaki": _lib.PREC_OZAKI}[pname]
api.init(1, 0)
lib = _lib.load()
dev = torch.device("cuda", 0)
be = D.DeviceBackend(dev)
n, p = 1000, 20000
P = p * (p - 1) // 2
if what == "c2":
    xd = torch.from_numpy(np.ascontiguousarray(synthetic_bulk(n, p, seed=2024, tie_heavy=True).T)).to(dev)
    out = torch.empty(P, dtype=torch.float64, device=dev)
    for _ in range(reps):
        _lib.check(lib.bixcor_dev_cor_upper_rows(xd.data_ptr(), n, p, 1, 1, 6.0, PREC, 0, p, out.data_ptr()))
elif what == "c2host":
    xh = torch.from_numpy(np.ascontiguousarray(synthetic_bulk(n, p, seed=2024, tie_heavy=True).T)).pin_memory()
    out = torch.empty(P, dtype=torch.float64).pin_memory()
    for _ in range(reps):
        _lib.check(lib.bixcor_cor_upper(xh.data_ptr(), n, p, 1, 1, 6.0, PREC, out.data_ptr()))
elif what == "tom":
    xd = torch.from_numpy(np.ascontiguousarray(synthetic_bulk(n, p, seed=2026).T)).to(dev)
    for _ in range(reps):
        out, _ = D.sharded_tom_from_exp(be, xd, n, p, False, True, "v1", beta=6.0, precision=PREC)
elif what == "c3":
    xa, xb = synthetic_pair(500, 500, p, seed=2025)
    xa, xb = (torch.from_numpy(np.ascontiguousarray(v.T)).to(dev) for v in (xa, xb))
    outs = [torch.empty(P, dtype=torch.float64, device=dev) for _ in range(4)]
    for _ in range(reps):
        _lib.check(lib.bixcor_dev_diffcor_rows(xa.data_ptr(), 500, xb.data_ptr(), 500, p, 1, PREC, 0, p, *[o.data_ptr() for o in outs]))
    out = outs[2]
torch.cuda.synchronize()
print("ok", what, pname, float(out[:1000].sum()))
