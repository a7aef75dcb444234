mkdir -p gpurun_out/r2s
python -m pytest tests -m gpu -x -q -k "transport or pageable or host or tom" 2>&1 | tail -2
python scripts/e2e_sweep.py --round2e --out gpurun_out/r2s/e2e_sweep_seq.json > gpurun_out/r2s/sweep.log 2>&1
python scripts/tom_e2e.py f16 tf32 > gpurun_out/r2s/tom_e2e.log 2>&1
