#!/bin/bash
# run 39: binary16 hi/lo split kind (BIXCOR_PREC_F16X3): parity tests, timings against 3xTF32
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/perf_probe.py tf32:0 f16:0 tf32:0:1.0 f16:0:1.0 > gpurun_out/f16_probe.log 2>&1
timeout 400 python scripts/ozaki_probe.py >> gpurun_out/f16_probe.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; grep -v "^cor_upper\|ozaki" gpurun_out/f16_probe.log | cut -c1-230
