#!/bin/bash
# run 28 (1 GPU): final N=1 evidence after the diffcor / Ozaki changes
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_n1.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_default.csv \
  python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
timeout 1700 python scripts/config_runs.py > gpurun_out/config_runs.log 2>&1
echo "configs exit $?" >> gpurun_out/config_runs.log
timeout 1200 python scripts/scale_runs.py > gpurun_out/scale_n1.log 2>&1
echo "scale n1 exit $?" >> gpurun_out/scale_n1.log
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -2 gpurun_out/bench_n1.log | cut -c1-200; grep -E "exit" gpurun_out/scale_n1.log gpurun_out/config_runs.log; grep -E "C3|C4" gpurun_out/scale_n1.log | cut -c1-200
