#!/bin/bash
# run 36: transposed interior tiles (stores from registers): full GPU suite + timing / checksums
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/perf_probe.py i8:0 i8:0:1.0 i8:0:6.0:500 i8:0:6.0:1800 tf32:0 tf32:0:1.0 > gpurun_out/swap_probe.log 2>&1
BIXCOR_UMMA_CTAS=1 timeout 300 python scripts/perf_probe.py i8:0 tf32:0 >> gpurun_out/swap_probe.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; cut -c1-220 gpurun_out/swap_probe.log
