mkdir -p gpurun_out/r2p
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2p/pytest.log 2>&1; tail -3 gpurun_out/r2p/pytest.log
python scripts/e2e_sweep.py --round2b --out gpurun_out/r2p/e2e_sweep_s32.json > gpurun_out/r2p/sweep.log 2>&1
python bench.py --no-alt > gpurun_out/r2p/bench.json 2> gpurun_out/r2p/bench.err; echo bench rc=$?
python scripts/tom_e2e.py f16 tf32 > gpurun_out/r2p/tom_e2e.log 2>&1; tail -4 gpurun_out/r2p/tom_e2e.log
