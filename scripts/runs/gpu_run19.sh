#!/bin/bash
# run 19: full GPU suite + bench after the CTA-pair kernels became the default
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_n1.log
BIXCOR_UMMA_CTAS=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n1_ctas1.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench_n1.log | cut -c1-600; grep -o '"tom".*' gpurun_out/bench_n1.log | cut -c1-700;  grep -o '"tom".*' gpurun_out/bench_n1_ctas1.log | cut -c1-700; grep -o '"paths".*"tom"' gpurun_out/bench_n1.log | cut -c1-900
