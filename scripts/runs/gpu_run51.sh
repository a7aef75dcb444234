mkdir -p gpurun_out/r2r
python scripts/e2e_sweep.py --round2d --out gpurun_out/r2r/e2e_sweep_s32geo.json > gpurun_out/r2r/sweep.log 2>&1
python scripts/tom_e2e.py f16 > gpurun_out/r2r/tom_e2e.log 2>&1
python -m pytest tests -m gpu -x -q -k "transport or pageable or host" 2>&1 | tail -2
