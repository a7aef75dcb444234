"""HBM write-only bandwidth on B200: fill_ / zero_ / copy of 1.6 GB (the packed result size of config 2)."""
import torch
dev = torch.device("cuda", 0)
n = 199_990_000
a = torch.empty(n, dtype=torch.float64, device=dev)
b = torch.empty(n, dtype=torch.float64, device=dev)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: a.fill_(1.5)); print(f"fill_ 1.6 GB: {ms:.3f} ms -> {8*n/ms/1e6:.0f} GB/s write")
ms = t(lambda: a.zero_()); print(f"zero_ 1.6 GB: {ms:.3f} ms -> {8*n/ms/1e6:.0f} GB/s write")
ms = t(lambda: b.copy_(a)); print(f"copy 1.6 GB: {ms:.3f} ms -> {16*n/ms/1e6:.0f} GB/s read+write")
ms = t(lambda: torch.mul(a, 2.0, out=b)); print(f"mul 1.6 GB: {ms:.3f} ms -> {16*n/ms/1e6:.0f} GB/s read+write")
ms = t(lambda: a.sum()); print(f"sum 1.6 GB: {ms:.3f} ms -> {8*n/ms/1e6:.0f} GB/s read")
