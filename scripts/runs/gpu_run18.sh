#!/bin/bash
# run 18: first run of the CTA-pair (cta_group::2) tcgen05 kernels
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
timeout 300 python scripts/umma_debug.py > gpurun_out/umma_debug.log 2>&1
echo "debug exit $?" >> gpurun_out/umma_debug.log
tail -18 gpurun_out/umma_debug.log
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x -k "int8 or tf32 or operand or fast or sparse or coremo or config2" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python scripts/perf_probe.py i8:0 i8:1 tf32:0 > gpurun_out/perf_probe.log 2>&1
BIXCOR_UMMA_CTAS=1 timeout 600 python scripts/perf_probe.py i8:0 tf32:0 >> gpurun_out/perf_probe.log 2>&1
cat gpurun_out/perf_probe.log
