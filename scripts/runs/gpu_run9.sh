#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python scripts/perf_probe.py i8:0 i8:2 i8:3 tf32:0 > gpurun_out/perf_probe.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_default.log
timeout 1500 python scripts/config_runs.py > gpurun_out/config_runs.log 2>&1
echo "configs exit $?" >> gpurun_out/config_runs.log
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/perf_probe.log; tail -3 gpurun_out/bench_default.log | cut -c1-3000; cat gpurun_out/config_runs.log | cut -c1-400
