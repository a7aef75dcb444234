"""torchrun helper: host -> host packed correlation of config 2 across ranks (pageable and page-locked buffers), e2e only.
   python -m torch.distributed.run --nproc-per-node N scripts/e2e_dist.py   (staging knobs come from the environment)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from bixverse_b200 import _lib, api
from bixverse_b200.distributed import DeviceBackend, host_sharded_cor_upper, packed_offset, pair_balanced_rows

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
api.init(1, local)
n, p = 1000, 20000
x = torch.from_numpy(np.ascontiguousarray(np.random.default_rng(0).normal(size=(p, n))))
r0, r1 = pair_balanced_rows(p, 1, world, rank)
ln = packed_offset(r1, p, 1) - packed_offset(r0, p, 1)
be, ws = DeviceBackend(dev), {}
res = {}
for label, xh, oh in (("pageable", x, torch.zeros(ln, dtype=torch.float64)), ("pinned", x.pin_memory(), torch.zeros(ln, dtype=torch.float64).pin_memory())):
    step = lambda: host_sharded_cor_upper(be, xh, n, p, True, oh, 1, 6.0, _lib.PREC_I8EXACT, workspace=ws)
    for _ in range(2):
        step()
    ts = []
    for _ in range(5):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        step()
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t.item()))
    res[label] = {"ms_min": 1e3 * min(ts), "ms_med": 1e3 * sorted(ts)[2]}
if rank == 0:
    env = {k: v for k, v in os.environ.items() if k.startswith("BIXCOR_")}
    print(json.dumps({"world": world, "env": env, **res}), flush=True)
dist.destroy_process_group()
