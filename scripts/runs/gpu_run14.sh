#!/bin/bash
# round-1 session 2, run 14: new rank kernel (FP64 compares, mirror-stage network) + lean packed epilogue
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python scripts/perf_probe.py i8:0 i8:1 i8:3 tf32:0 > gpurun_out/perf_probe.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_n1.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-tom --no-cpu --no-alt > gpurun_out/plain_i8.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gram_umma|rank_columns" -s 6 -c 2 -o gpurun_out/prof_i8_r01c \
  python bench.py --steps 2 --warmup 3 --no-tom --no-cpu --no-alt > gpurun_out/ncu_i8.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/perf_probe.log; tail -2 gpurun_out/bench_n1.log | cut -c1-1200
