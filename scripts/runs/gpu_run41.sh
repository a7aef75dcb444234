#!/bin/bash
# run 41: f16 kind with promotion every 3 K blocks + CTA pairs: parity tests and TOM / Gram timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "tf32x3 or fuzz or operand_exchange or tom or smoke" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/f16_probe.py > gpurun_out/f16_tom.log 2>&1
timeout 300 python scripts/perf_probe.py f16:0 f16:0:1.0 tf32:0 >> gpurun_out/f16_tom.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cut -c1-260 gpurun_out/f16_tom.log
