#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -k "sibling or covariance" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-tom --no-cpu --no-alt > gpurun_out/plain_i8.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gram_umma|rank_columns" -s 6 -c 2 -o gpurun_out/prof_i8_r01b \
  python bench.py --steps 2 --warmup 3 --no-tom --no-cpu --no-alt > gpurun_out/ncu_i8.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/ncu_i8.log | cut -c1-300
