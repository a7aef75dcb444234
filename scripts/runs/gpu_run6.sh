#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 1200 python scripts/perf_probe.py > gpurun_out/perf_probe.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/perf_probe.log
