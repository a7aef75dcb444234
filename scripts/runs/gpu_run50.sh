mkdir -p gpurun_out/r2q
python scripts/e2e_sweep.py --round2c --out gpurun_out/r2q/e2e_sweep_small.json > gpurun_out/r2q/sweep.log 2>&1
BIXCOR_STAGE_SLOT_KB=1024 BIXCOR_STAGE_SLOTS=32 BIXCOR_STAGE_PIECE_KB=256 python scripts/tom_e2e.py f16 tf32 f64 > gpurun_out/r2q/tom_e2e_small.log 2>&1
