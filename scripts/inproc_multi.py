"""Single-process multi-GPU mode (bixcor_init(n): what an R session linking libbixcor uses; the library opens NCCL itself):
host -> host wall times of configs 2-5 on 1..N devices of one process.  Usage: python scripts/inproc_multi.py [devices ...]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bixverse_b200 import _lib, api
from bixverse_b200.synthetic import synthetic_bulk, synthetic_pair, synthetic_sparse_cells_fast

lib = _lib.load()
ndev_max = torch.cuda.device_count()
devs = [int(a) for a in sys.argv[1:]] or [d for d in (1, 2, 4, 8) if d <= ndev_max]
n, p = 1000, 20000
P = p * (p - 1) // 2
x_np = np.ascontiguousarray(synthetic_bulk(n, p, seed=2024, tie_heavy=True).T)
x = torch.from_numpy(x_np).pin_memory()
out = torch.empty(P, dtype=torch.float64).pin_memory()
out_page = np.zeros(P)
xa, xb = synthetic_pair(500, 500, p, seed=2025)
xa = torch.from_numpy(np.ascontiguousarray(xa.T)).pin_memory()
xb = torch.from_numpy(np.ascontiguousarray(xb.T)).pin_memory()
outs = [torch.empty(P, dtype=torch.float64).pin_memory() for _ in range(4)]
cells, genes = 1000000, 5000
ip, ix, dd = synthetic_sparse_cells_fast(cells, genes, 2027)
sp_out = np.zeros(genes * (genes - 1) // 2)
rows = []


def timed(fn, reps=3):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts[1:]))


def add(**kw):
    rows.append(kw)
    print(json.dumps(kw), flush=True)


for nd in devs:
    api.init(nd, 0)
    nccl = int(lib.bixcor_nccl_version())
    t = timed(lambda: _lib.check(lib.bixcor_cor_upper(x.data_ptr(), n, p, 1, 1, 6.0, _lib.PREC_AUTO, out.data_ptr())), 4)
    add(config="C2 host->host (pinned), one process", devices=nd, precision="auto(i8)", ms=1e3 * t, pairs_per_s=P / t, checksum=float(out[:1000000].sum()), nccl=nccl)
    t = timed(lambda: _lib.check(lib.bixcor_cor_upper(x_np.ctypes.data, n, p, 1, 1, 6.0, _lib.PREC_AUTO, out_page.ctypes.data)), 4)
    add(config="C2 host->host (pageable), one process", devices=nd, precision="auto(i8)", ms=1e3 * t, pairs_per_s=P / t, checksum=float(out_page[:1000000].sum()))
    t = timed(lambda: _lib.check(lib.bixcor_diffcor(xa.data_ptr(), 500, xb.data_ptr(), 500, p, 0, _lib.PREC_F64, *[o_.data_ptr() for o_ in outs])))
    add(config="C3 diffcor host->host (pinned), one process", devices=nd, precision="f64", ms=1e3 * t, pairs_per_s=P / t)
    for name, prec in (("f16", _lib.PREC_F16X3), ("tf32", _lib.PREC_TF32X3), ("ozaki", _lib.PREC_OZAKI), ("f64", _lib.PREC_F64)):
        t = timed(lambda: _lib.check(lib.bixcor_tom_from_exp(x.data_ptr(), n, p, 0, 1, _lib.TOM_V1, 6.0, prec, 1, out.data_ptr())))
        add(config="C4 TOM-from-exp host->host (pinned), one process", devices=nd, precision=name, ms=1e3 * t, checksum=float(out[:1000000].sum()))
    for name, prec in (("f16", _lib.PREC_F16X3), ("tf32", _lib.PREC_TF32X3)):
        t = timed(lambda: _lib.check(lib.bixcor_sparse_gram_cor(ip.ctypes.data, ix.ctypes.data, dd.ctypes.data, cells, genes, prec, sp_out.ctypes.data)))
        add(config="C5 sparse Gram 1M x 5k host->host (pageable CSR), one process", devices=nd, precision=name, ms=1e3 * t, checksum=float(sp_out[:100000].sum()))
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/inproc_multi_r02.json", "w") as f:
    for r in rows:
        f.write(json.dumps(r) + "\n")
