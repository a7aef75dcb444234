#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -k "rank or cor_upper or diffcor or int8" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
for v in 1 3; do
  BIXCOR_DMMA_VARIANT=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-tom --no-cpu > gpurun_out/bench_f64_v$v.log 2>&1
done
timeout 600 python bench.py --steps 5 --warmup 3 --precision i8 --no-cpu --no-tom > gpurun_out/bench_i8.log 2>&1
# ncu: launch list of the default bench command, then full sets of the two Gram kernels
timeout 600 python bench.py --steps 2 --warmup 3 --no-tom --no-cpu > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01.csv \
  python bench.py --steps 2 --warmup 3 --no-tom --no-cpu > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_dmma -s 2 -c 1 -o gpurun_out/prof_dmma_r01 \
  python bench.py --steps 2 --warmup 3 --no-tom --no-cpu > gpurun_out/ncu_dmma.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 --no-tom --no-cpu --precision i8 > gpurun_out/plain_i8.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gram_umma|rank_columns" -s 4 -c 2 -o gpurun_out/prof_i8_r01 \
  python bench.py --steps 2 --warmup 3 --no-tom --no-cpu --precision i8 > gpurun_out/ncu_i8.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
