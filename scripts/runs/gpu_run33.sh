#!/bin/bash
# run 33: full GPU suite (new: rbh, cells_to_keep, many-cell Spearman, fused cor_module_tom) + smoke
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
tail -15 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log
