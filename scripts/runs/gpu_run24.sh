#!/bin/bash
# run 24 (1 GPU): full suite, bench, reference arm, ncu launch list + full captures, all configs, scale_runs with CPU baselines
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_n1.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_default.csv \
  python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
for prec in i8 tf32 f64; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gram_umma|gram_dmma|rank_columns" -s 6 -c 2 -o gpurun_out/prof_${prec}_r01f \
    python bench.py --steps 2 --warmup 3 --no-tom --no-cpu --no-alt --precision $prec > gpurun_out/ncu_${prec}.log 2>&1
done
timeout 900 ncu --set full --clock-control none -k regex:"standardise" -c 1 -o gpurun_out/prof_std_r01f python scripts/tom_probe.py > gpurun_out/ncu_std.log 2>&1
timeout 1700 python scripts/config_runs.py > gpurun_out/config_runs.log 2>&1
echo "configs exit $?" >> gpurun_out/config_runs.log
timeout 1200 python scripts/scale_runs.py > gpurun_out/scale_n1.log 2>&1
echo "scale n1 exit $?" >> gpurun_out/scale_n1.log
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -2 gpurun_out/bench_n1.log | cut -c1-500; tail -1 gpurun_out/bench_ref_n1.log | cut -c1-300; grep -E "^\{|exit" gpurun_out/scale_n1.log | cut -c1-250; tail -3 gpurun_out/config_runs.log | cut -c1-200
