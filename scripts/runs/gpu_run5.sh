#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "cor_upper or diffcor or tom or sparse" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 1200 python scripts/perf_probe.py > gpurun_out/perf_probe.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/perf_probe.log
