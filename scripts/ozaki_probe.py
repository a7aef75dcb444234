"""Ozaki int8 path at full size on one GPU: time and accuracy against the FP64 DMMA path (config 2 Pearson Gram, config 4 TOM)."""
import os, sys, time
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
res = {}
for name, prec in (("f64", _lib.PREC_F64), ("ozaki", _lib.PREC_OZAKI)):
    for spearman in (0, 1):
        out = be.cor_upper_rows(xd, n, p, spearman, 1, 1.0, prec, 0, p)
        _lib.check(lib.bixcor_timing_enable(1))
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = be.cor_upper_rows(xd, n, p, spearman, 1, 1.0, prec, 0, p)
        torch.cuda.synchronize(); t = time.perf_counter() - t0
        tm = _lib.last_timing(); _lib.check(lib.bixcor_timing_enable(0))
        print(f"cor_upper {name} spearman={spearman}: wall {1e3*t:.2f} ms  kernels {tm}", flush=True)
        res[(name, spearman)] = out
for spearman in (0, 1):
    d = (res[("f64", spearman)] - res[("ozaki", spearman)]).abs().max().item()
    print(f"max |f64 - ozaki| packed correlation, spearman={spearman}: {d:.3e}", flush=True)
del res
toms = {}
for name, prec in (("f64", _lib.PREC_F64), ("ozaki", _lib.PREC_OZAKI), ("ozaki4", _lib.PREC_OZAKI4), ("tf32", _lib.PREC_TF32X3), ("f16", _lib.PREC_F16X3)):
    D.sharded_tom_from_exp(be, xd, n, p, False, True, "v1", beta=6.0, precision=prec)
    _lib.check(lib.bixcor_timing_enable(1))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out, _ = D.sharded_tom_from_exp(be, xd, n, p, False, True, "v1", beta=6.0, precision=prec)
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    tm = _lib.last_timing(); _lib.check(lib.bixcor_timing_enable(0))
    print(f"TOM {name}: wall {1e3*t:.2f} ms  kernels {tm}", flush=True)
    toms[name] = out
print("max |f64 - ozaki| TOM:", (toms["f64"] - toms["ozaki"]).abs().max().item(), " max |f64 - ozaki4| TOM:", (toms["f64"] - toms["ozaki4"]).abs().max().item(), " max |f64 - tf32| TOM:", (toms["f64"] - toms["tf32"]).abs().max().item(), " max |f64 - f16| TOM:", (toms["f64"] - toms["f16"]).abs().max().item())
for name, prec in (("tf32", _lib.PREC_TF32X3), ("f16", _lib.PREC_F16X3)):  # beta = 1 and Pearson Gram accuracy of the split kinds
    out, _ = D.sharded_tom_from_exp(be, xd, n, p, False, True, "v1", beta=1.0, precision=prec)
    ref, _ = D.sharded_tom_from_exp(be, xd, n, p, False, True, "v1", beta=1.0, precision=_lib.PREC_F64)
    print(f"max |f64 - {name}| TOM beta=1:", (ref - out).abs().max().item(), flush=True)
    g = be.cor_upper_rows(xd, n, p, 0, 1, 1.0, prec, 0, p)
    gr = be.cor_upper_rows(xd, n, p, 0, 1, 1.0, _lib.PREC_F64, 0, p)
    print(f"max |f64 - {name}| Pearson packed:", (g - gr).abs().max().item(), flush=True)
# beta = 1 TOM (reference semantics, dense affinities: the hard case for fixed-point slices)
t1 = {}
for name, prec in (("f64", _lib.PREC_F64), ("ozaki", _lib.PREC_OZAKI)):
    out, _ = D.sharded_tom_from_exp(be, xd, n, p, False, True, "v1", beta=1.0, precision=prec)
    t1[name] = out
print("max |f64 - ozaki| TOM beta=1:", (t1["f64"] - t1["ozaki"]).abs().max().item())
