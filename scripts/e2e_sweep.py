"""Sweep of the pageable staging pipeline (HostStager in csrc/bixcor.cu): ring geometry, copier threads, streaming stores.
Each setting runs in its own process (the knobs are read at bixcor_init).  Usage: python scripts/e2e_sweep.py [--out f.json]
Child mode: python scripts/e2e_sweep.py --child  (prints one JSON line for the current environment)."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np

    from bixverse_b200 import _lib, api

    n, p = 1000, 20000
    x = np.ascontiguousarray(np.random.default_rng(0).normal(size=(p, n)))  # gene-major == R's column-major n x p
    lib = _lib.load()
    api.init(1, 0)
    out = np.zeros(p * (p - 1) // 2)
    step = lambda: _lib.check(lib.bixcor_cor_upper(x.ctypes.data, n, p, 1, 1, 6.0, _lib.PREC_AUTO, out.ctypes.data))
    for _ in range(2):
        step()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    print(json.dumps({"ms_min": 1e3 * min(ts), "ms_med": 1e3 * sorted(ts)[len(ts) // 2]}))


def main():
    settings = [{}]
    if "--round2" in sys.argv:
        for slot_mb, slots, piece in ((8, 8, 1024), (4, 16, 512), (2, 16, 256), (4, 8, 512), (16, 8, 1024), (8, 16, 512)):
            for spin in (0, 1):
                settings.append({"BIXCOR_STAGE_SLOT_MB": slot_mb, "BIXCOR_STAGE_SLOTS": slots, "BIXCOR_STAGE_PIECE_KB": piece, "BIXCOR_STAGE_SPIN": spin})
        settings.append({"BIXCOR_STAGE_SLOT_MB": 8, "BIXCOR_STAGE_SLOTS": 8, "BIXCOR_STAGE_PIECE_KB": 1024, "BIXCOR_COPY_THREADS": 10})
        settings.append({"BIXCOR_STAGE_SLOT_MB": 8, "BIXCOR_STAGE_SLOTS": 8, "BIXCOR_STAGE_PIECE_KB": 1024, "BIXCOR_COPY_THREADS": 16})
    elif "--round2b" in sys.argv:  # int32 transport (4 bytes per pair on the wire, the copier threads expand to f64)
        settings.append({"BIXCOR_S32_TRANSPORT": 0})
        for thr in (6, 8, 10, 12, 16):
            settings.append({"BIXCOR_COPY_THREADS": thr})
        for slot_mb, slots, piece in ((2, 32, 256), (2, 16, 256), (4, 16, 256), (4, 16, 1024), (8, 8, 512), (1, 32, 256), (4, 32, 512)):
            settings.append({"BIXCOR_STAGE_SLOT_MB": slot_mb, "BIXCOR_STAGE_SLOTS": slots, "BIXCOR_STAGE_PIECE_KB": piece})
        settings.append({"BIXCOR_STAGE_SPIN": 0})
    elif "--round2c" in sys.argv:  # small rings (the ring should stay in the cache level the DMA writes into)
        for slot_kb, slots, piece in ((1024, 32, 256), (1024, 16, 256), (1024, 64, 256), (512, 32, 256), (512, 64, 256), (512, 32, 128), (1024, 32, 128),
                                      (256, 64, 128), (2048, 8, 256)):
            settings.append({"BIXCOR_STAGE_SLOT_KB": slot_kb, "BIXCOR_STAGE_SLOTS": slots, "BIXCOR_STAGE_PIECE_KB": piece})
        settings.append({"BIXCOR_STAGE_SLOT_KB": 1024, "BIXCOR_STAGE_SLOTS": 32, "BIXCOR_STAGE_PIECE_KB": 256, "BIXCOR_COPY_THREADS": 16})
        settings.append({"BIXCOR_STAGE_SLOT_KB": 1024, "BIXCOR_STAGE_SLOTS": 32, "BIXCOR_STAGE_PIECE_KB": 256, "BIXCOR_COPY_THREADS": 10})
        settings.append({"BIXCOR_STAGE_SLOT_KB": 1024, "BIXCOR_STAGE_SLOTS": 32, "BIXCOR_STAGE_PIECE_KB": 256, "BIXCOR_STAGE_SPIN": 0})
        settings.append({"BIXCOR_STAGE_SLOT_KB": 1024, "BIXCOR_STAGE_SLOTS": 32, "BIXCOR_STAGE_PIECE_KB": 256, "BIXCOR_S32_TRANSPORT": 0})
    elif "--round2d" in sys.argv:  # ring geometry of the int32 transport alone (BIXCOR_S32_*)
        for slot_kb, slots, piece in ((4096, 16, 512), (1024, 32, 256), (1024, 24, 256), (1024, 48, 256), (2048, 16, 256), (2048, 32, 512), (1024, 32, 512), (1536, 32, 256)):
            settings.append({"BIXCOR_S32_SLOT_KB": slot_kb, "BIXCOR_S32_SLOTS": slots, "BIXCOR_S32_PIECE_KB": piece})
        settings.append({})
    elif "--round2e" in sys.argv:  # landing signal: stream-written sequence word vs CUDA events; repeated to see the run-to-run spread
        for rep in range(3):
            settings.append({"BIXCOR_STAGE_SEQ": 1})
            settings.append({"BIXCOR_STAGE_SEQ": 0})
        for slot_kb, slots, piece in ((4096, 16, 512), (2048, 16, 256), (1024, 16, 256), (512, 32, 128), (512, 64, 256), (1024, 32, 128)):
            settings.append({"BIXCOR_S32_SLOT_KB": slot_kb, "BIXCOR_S32_SLOTS": slots, "BIXCOR_S32_PIECE_KB": piece})
        for thr in (12, 16):
            settings.append({"BIXCOR_COPY_THREADS": thr})
    else:
        for slot_mb, slots in ((32, 8), (16, 8), (8, 8), (4, 8), (2, 16), (8, 16), (64, 4)):
            settings.append({"BIXCOR_STAGE_SLOT_MB": slot_mb, "BIXCOR_STAGE_SLOTS": slots})
        for thr in (4, 8, 12, 16):
            settings.append({"BIXCOR_COPY_THREADS": thr})
        settings.append({"BIXCOR_STAGE_NT": 0})
        settings.append({"BIXCOR_STAGE_NT": 0, "BIXCOR_STAGE_SLOT_MB": 4, "BIXCOR_STAGE_SLOTS": 8})
        settings.append({"BIXCOR_STAGE_PIECE_KB": 1024})
        settings.append({"BIXCOR_STAGE_PIECE_KB": 1024, "BIXCOR_STAGE_SLOT_MB": 8, "BIXCOR_STAGE_SLOTS": 8})
    rows = []
    for s in settings:
        env = dict(os.environ)
        env.update({k: str(v) for k, v in s.items()})
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True)
        try:
            res = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            res = {"error": (r.stderr or r.stdout)[-300:]}
        rows.append({"env": s, **res})
        print(json.dumps(rows[-1]), flush=True)
    if "--out" in sys.argv:
        json.dump(rows, open(sys.argv[sys.argv.index("--out") + 1], "w"), indent=1)


if __name__ == "__main__":
    child() if "--child" in sys.argv else main()
