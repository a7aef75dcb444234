"""Rank kernel variants (csrc/columns.cuh): times the column kernel and the whole config-2 step per library build.
Usage: python scripts/rank_probe.py [lib.so ...]   (child mode: --child, library from BIXCOR_LIB)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np
    import torch

    from bixverse_b200 import _lib, api
    from bixverse_b200.synthetic import synthetic_bulk

    lib = _lib.load()
    api.init(1, 0)
    dev = torch.device("cuda", 0)
    res = {}
    for n, p, tie in ((1000, 20000, True), (1000, 20000, False), (1000, 2500, True), (500, 20000, True)):
        x = synthetic_bulk(n, p, seed=2024, tie_heavy=tie)
        xd = torch.from_numpy(np.ascontiguousarray(x.T)).to(dev)
        out = torch.empty(p * (p - 1) // 2, dtype=torch.float64, device=dev)
        step = lambda: _lib.check(lib.bixcor_dev_cor_upper_rows(xd.data_ptr(), n, p, 1, 1, 6.0, _lib.PREC_AUTO, 0, p, out.data_ptr()))
        for _ in range(3):
            step()
        _lib.check(lib.bixcor_timing_enable(1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            step()
        e1.record()
        torch.cuda.synchronize()
        tm = _lib.last_timing()
        _lib.check(lib.bixcor_timing_enable(0))
        res[f"n{n}_p{p}_{'ties' if tie else 'noties'}"] = {"cols_ms": tm["cols_ms"] / 10, "gemm_ms": tm["gemm_ms"] / 10, "checksum": float(out[:100000].sum())}
    print(json.dumps(res))


def main():
    libs = sys.argv[1:] or [os.path.join(ROOT, "bixverse_b200", "libbixcor.so")]
    for lp in libs:
        env = dict(os.environ, BIXCOR_LIB=os.path.abspath(lp))
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True)
        try:
            print(os.path.basename(lp), r.stdout.strip().splitlines()[-1], flush=True)
        except Exception:
            print(os.path.basename(lp), "ERROR", (r.stderr or r.stdout)[-400:], flush=True)


if __name__ == "__main__":
    child() if "--child" in sys.argv else main()
