"""Every BASELINE.json shape at full size on N GPUs (torchrun, one process per GPU), inputs resident in HBM:
CUDA-event time of the whole sharded job, max over ranks.  Run it at N = 1, 2, 4, 8; one JSON object per config.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/scale_runs.py
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from bixverse_b200 import _lib, api
from bixverse_b200 import distributed as D
from bixverse_b200.synthetic import synthetic_bulk, synthetic_pair, synthetic_sparse_cells_fast

world = int(os.environ.get("WORLD_SIZE", 1))
rank = int(os.environ.get("RANK", 0))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
api.init(1, local)
be = D.DeviceBackend(dev)
OUT = []


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def timed(fn, reps=3):
    fn()  # warm-up (allocations, tile lists, NCCL channels)
    ts = []
    for _ in range(reps):
        barrier()
        t0 = time.perf_counter()
        r = fn()
        barrier()
        ts.append(time.perf_counter() - t0)
        del r
    t = torch.tensor([float(np.median(ts))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def emit(d):
    d["n_gpus"] = world
    if rank == 0:
        OUT.append(d)
        print(json.dumps(d), flush=True)


P = 199_990_000
n, p = 1000, 20000
x = torch.from_numpy(np.ascontiguousarray(synthetic_bulk(n, p, seed=2024, tie_heavy=True).T)).to(dev)
for name, prec in (("i8", _lib.PREC_I8EXACT), ("f16", _lib.PREC_F16X3), ("tf32", _lib.PREC_TF32X3), ("ozaki-int8", _lib.PREC_OZAKI), ("f64", _lib.PREC_F64)):
    if prec == _lib.PREC_OZAKI and world > 1:
        continue  # the operand-exchange API carries precisions 0..2; the Ozaki Gram is timed on one GPU
    t = timed(lambda: D.sharded_cor_upper(be, x, n, p, True, 1, 6.0, prec))
    emit({"config": "C2 Spearman 20000x1000 |r|^6 packed", "precision": name, "ms": 1e3 * t, "pairs_per_s": P / t})
for name, prec in (("f16", _lib.PREC_F16X3), ("tf32", _lib.PREC_TF32X3), ("ozaki-int8", _lib.PREC_OZAKI), ("f64", _lib.PREC_F64)):
    t = timed(lambda: D.sharded_tom_from_exp(be, x, n, p, False, True, "v1", beta=6.0, precision=prec), reps=2)
    emit({"config": "C4 TOM-from-exp signed v1 beta=6 20000x1000 packed", "precision": name, "ms": 1e3 * t, "alg_tflops": 8.4e12 / t / 1e12})
# north-star target: upper-triangle correlation + 20k-gene TOM from a HOST expression matrix, results back on the host
x_pin = x.cpu().pin_memory()
r0, r1 = D.pair_balanced_rows(p, 1, world, rank)
my_len = D.packed_offset(r1, p, 1) - D.packed_offset(r0, p, 1)
cor_pin = torch.empty(max(my_len, 1), dtype=torch.float64).pin_memory()
tom_pin = torch.empty(max(my_len, 1), dtype=torch.float64).pin_memory()
ws = {}
for name, prec, tprec in (("f64 (FP64 DMMA Gram + TOM)", _lib.PREC_F64, _lib.PREC_F64),
                          ("exact-class (FP64 DMMA Gram, Ozaki int8 TOM)", _lib.PREC_F64, _lib.PREC_OZAKI),
                          ("fast (exact int8 Gram, 3xTF32 TOM)", _lib.PREC_I8EXACT, _lib.PREC_TF32X3),
                          ("fast (exact int8 Gram, 3 x f16 TOM)", _lib.PREC_I8EXACT, _lib.PREC_F16X3)):
    def job():
        D.host_sharded_cor_upper(be, x_pin, n, p, prec == _lib.PREC_I8EXACT, cor_pin, 1, 1.0, prec, workspace=ws)
        xd = x_pin.to(dev, non_blocking=True)
        t_out, _ = D.sharded_tom_from_exp(be, xd, n, p, False, True, "v1", beta=6.0, precision=tprec)
        tom_pin[: t_out.numel()].copy_(t_out)
        torch.cuda.synchronize(dev)
    t = timed(job, reps=2)
    emit({"config": "north-star: packed correlation + signed TOM (beta 6) of 20000x1000, host -> host", "precision": name, "ms": 1e3 * t})
del x, x_pin, cor_pin, tom_pin
xa, xb = synthetic_pair(500, 500, p, seed=2025)
xa = torch.from_numpy(np.ascontiguousarray(xa.T)).to(dev)
xb = torch.from_numpy(np.ascontiguousarray(xb.T)).to(dev)
for name, prec, sp in (("f64-pearson", _lib.PREC_F64, False), ("i8-spearman", _lib.PREC_I8EXACT, True), ("tf32-pearson", _lib.PREC_TF32X3, False), ("f16-pearson", _lib.PREC_F16X3, False), ("ozaki-pearson", _lib.PREC_OZAKI, False)):
    t = timed(lambda: D.sharded_diffcor(be, xa, 500, xb, 500, p, sp, prec), reps=2)
    emit({"config": "C3 diffcor 20000 genes 500 vs 500", "precision": name, "ms": 1e3 * t, "pairs_per_s": P / t})
del xa, xb
cells, genes = 1_000_000, 5000
c0, c1 = cells * rank // world, cells * (rank + 1) // world
ip, ix, dd = synthetic_sparse_cells_fast(c1 - c0, genes, seed=2027 + rank)  # this rank's cell shard
lp, li, ld = torch.from_numpy(ip).to(dev), torch.from_numpy(ix).to(dev), torch.from_numpy(dd).to(dev)
for name, prec in (("f16", _lib.PREC_F16X3), ("tf32", _lib.PREC_TF32X3), ("f64", _lib.PREC_F64)):
    t = timed(lambda: D.sharded_sparse_gram_cor(be, lp, li, ld, c1 - c0, genes, prec), reps=2)
    emit({"config": "C5 sparse CSR 1M cells x 5000 genes, cell shards + Gram all-reduce", "precision": name, "ms": 1e3 * t,
          "pairs_per_s": genes * (genes - 1) // 2 / t})
# ---- CPU baseline of every shape on the box's host cores (N = 1 run only, bounded samples of the same workloads):
# the NumPy/OpenBLAS oracle port following the reference's algorithm shape (the Rust crate cannot be built here)
if world == 1 and os.environ.get("SCALE_CPU", "1") != "0":
    from oracle import oracle as o

    try:
        from threadpoolctl import threadpool_info

        cores = max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
    except Exception:
        cores = os.cpu_count() or 1
    gs = 6000
    xs = synthetic_bulk(1000, gs, seed=2024, tie_heavy=True)
    t0 = time.perf_counter(); o.soft_threshold(o.cor_upper_triangle(xs, True, True), 6.0); t = time.perf_counter() - t0
    emit({"config": "C2 Spearman 20000x1000 |r|^6 packed", "precision": "cpu oracle port", "cores": cores, "sample": f"first {gs} genes",
          "ms": 1e3 * t, "pairs_per_s": gs * (gs - 1) / 2 / t})
    xa_s, xb_s = synthetic_pair(500, 500, gs, seed=2025)
    t0 = time.perf_counter(); o.differential_cor(xa_s, xb_s, False); t = time.perf_counter() - t0
    emit({"config": "C3 diffcor 20000 genes 500 vs 500", "precision": "cpu oracle port", "cores": cores, "sample": f"first {gs} genes",
          "ms": 1e3 * t, "pairs_per_s": gs * (gs - 1) / 2 / t})
    t0 = time.perf_counter(); o.tom_from_exp(xs, True, "v1", False, 6.0); t = time.perf_counter() - t0
    emit({"config": "C4 TOM-from-exp signed v1 beta=6 20000x1000 packed", "precision": "cpu oracle port", "cores": cores,
          "sample": f"first {gs} genes (full square A*A as the reference does)", "ms": 1e3 * t,
          "alg_tflops": (2.0 * 1000 * gs * (gs - 1) / 2 + 2.0 * gs * gs * (gs - 1) / 2) / t / 1e12})
    cs, gsp = 100_000, 2000
    ip_s, ix_s, d_s = synthetic_sparse_cells_fast(cs, gsp, seed=2027)
    t0 = time.perf_counter(); o.sparse_gram_cor(ip_s, ix_s, d_s, cs, gsp); t = time.perf_counter() - t0
    emit({"config": "C5 sparse CSR 1M cells x 5000 genes, cell shards + Gram all-reduce", "precision": "cpu oracle port", "cores": cores,
          "sample": f"{cs} cells x {gsp} genes", "ms": 1e3 * t, "pairs_per_s": gsp * (gsp - 1) / 2 / t, "pair_cells_per_s": gsp * (gsp - 1) / 2 * cs / t})
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/scale_r01_n{world}.json", "w") as f:
        for d_ in OUT:
            f.write(json.dumps(d_) + "\n")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
