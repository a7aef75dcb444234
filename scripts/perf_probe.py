"""Timing probes on config 2 (device resident): DMMA shapes, UMMA debug modes."""
import os, sys, json, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    from bixverse_b200 import _lib, api
    from bixverse_b200.synthetic import synthetic_bulk
    prec = int(sys.argv[2]); spearman = int(sys.argv[3]); n = int(sys.argv[4]); p = int(sys.argv[5])
    api.init(1, 0); lib = _lib.load()
    x = synthetic_bulk(n, p, seed=1, tie_heavy=True)
    dev = torch.device("cuda", 0)
    xd = torch.from_numpy(np.ascontiguousarray(x.T)).to(dev)
    out = torch.empty(p * (p - 1) // 2, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    beta = float(os.environ.get("PROBE_BETA", "6.0"))
    for _ in range(3):
        _lib.check(lib.bixcor_dev_cor_upper_rows(xd.data_ptr(), n, p, spearman, 1, beta, prec, 0, p, out.data_ptr()))
    _lib.check(lib.bixcor_timing_enable(1))
    for _ in range(5):
        _lib.check(lib.bixcor_dev_cor_upper_rows(xd.data_ptr(), n, p, spearman, 1, beta, prec, 0, p, out.data_ptr()))
    t = _lib.last_timing()
    torch.cuda.synchronize()
    print(json.dumps({"gemm_ms": t["gemm_ms"] / 5, "cols_ms": t["cols_ms"] / 5,
                      "checksum": int(out.view(torch.int64).sum().item()), "lib": os.path.basename(_lib.LIB_PATH)}))
    sys.exit(0)

def run(tag, prec, spearman=1, n=1000, p=20000, **env):
    e = dict(os.environ); e.update({k: str(v) for k, v in env.items()})
    r = subprocess.run([sys.executable, __file__, "child", str(prec), str(spearman), str(n), str(p)], env=e, capture_output=True, text=True, timeout=600)
    print(tag, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ("ERR " + r.stderr[-300:]), flush=True)
    prof = [ln for ln in r.stderr.splitlines() if "[umma prof]" in ln]
    if prof:
        print("   ", prof[-1], flush=True)

probes = sys.argv[1:] or ["i8:0", "i8:1", "i8:2", "i8:3", "tf32:0", "tf32:1"]
for pr in probes:
    kind, d, *rest = pr.split(":")
    beta = rest[0] if rest else "6.0"
    nn = int(rest[1]) if len(rest) > 1 else 1000
    run(f"umma {kind} dbg {d} beta {beta} n {nn} ctas {os.environ.get('BIXCOR_UMMA_CTAS', '2')}", {"i8": 2, "tf32": 1, "f16": 5}[kind], n=nn,
        BIXCOR_UMMA_DEBUG=d, PROBE_BETA=beta)
