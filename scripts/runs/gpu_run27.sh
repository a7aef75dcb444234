#!/bin/bash
# run 27 (2 GPUs): NCCL driver test incl. the Ozaki TOM, full suite, refreshed N=1 bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_n1.log
tail -2 gpurun_out/bench_n1.log | cut -c1-200
CUDA_VISIBLE_DEVICES=0 timeout 600 python scripts/scale_runs.py > gpurun_out/scale_n1.log 2>&1
echo "scale exit $?" >> gpurun_out/scale_n1.log
grep -E "ozaki|exit" gpurun_out/scale_n1.log | cut -c1-220
