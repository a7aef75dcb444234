#!/bin/bash
# run 15: full GPU tests, N=1 bench, ncu launch list + full captures of the three Gram paths
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_n1.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_default.csv \
  python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
for prec in i8 tf32 f64; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gram_umma|gram_dmma|rank_columns" -s 6 -c 2 -o gpurun_out/prof_${prec}_r01d \
    python bench.py --steps 2 --warmup 3 --no-tom --no-cpu --no-alt --precision $prec > gpurun_out/ncu_${prec}.log 2>&1
done
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench_n1.log | cut -c1-3000
