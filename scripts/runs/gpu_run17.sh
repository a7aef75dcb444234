#!/bin/bash
# run 17: GPU tests incl. sparse TF32 path + all five configs at full size
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 1700 python scripts/config_runs.py > gpurun_out/config_runs.log 2>&1
echo "configs exit $?" >> gpurun_out/config_runs.log
tail -4 gpurun_out/pytest_gpu.log; cut -c1-420 gpurun_out/config_runs.log
