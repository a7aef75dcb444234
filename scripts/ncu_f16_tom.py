"""One f16-kind TOM at 20 000 genes (for an ncu capture of gram_umma_kernel<KIND_F16, EPI_TOM>)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bixverse_b200 import _lib, api
from bixverse_b200 import distributed as D
from bixverse_b200.synthetic import synthetic_bulk
api.init(1, 0)
dev = torch.device("cuda", 0); be = D.DeviceBackend(dev)
n, p = 1000, 20000
xd = torch.from_numpy(np.ascontiguousarray(synthetic_bulk(n, p, seed=2026).T)).to(dev)
out, _ = D.sharded_tom_from_exp(be, xd, n, p, False, True, "v1", beta=6.0, precision=_lib.PREC_F16X3)
torch.cuda.synchronize()
print("ok", float(out[:1000].sum()))
