#!/bin/bash
# run 43: full GPU suite with the f16 sparse path and overflow guards; N = 1 scale rows (C5 f16)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
SCALE_CPU=0 timeout 400 python scripts/scale_runs.py > gpurun_out/scale_n1.log 2>&1
echo "scale n1 exit $?" >> gpurun_out/scale_n1.log
tail -3 gpurun_out/pytest_gpu.log; grep "C5\|exit" gpurun_out/scale_n1.log | cut -c1-220
