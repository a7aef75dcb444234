#!/bin/bash
# run 25 (1 GPU): final round-1 evidence: full suite, smoke, bench, reference arm, ncu launch list, ncu full captures
# (exported to text on the box; the .ncu-rep files exceed the 64 MiB return limit), all configs, scale_runs + CPU baselines
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
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gram_umma|gram_dmma|rank_columns" -s 6 -c 2 -o /tmp/prof_${prec} \
    python bench.py --steps 2 --warmup 3 --no-tom --no-cpu --no-alt --precision $prec > gpurun_out/ncu_${prec}.log 2>&1
  ncu -i /tmp/prof_${prec}.ncu-rep --page details > gpurun_out/ncu_${prec}_details.txt 2>/dev/null
  ncu -i /tmp/prof_${prec}.ncu-rep --page raw --csv > gpurun_out/ncu_${prec}_raw.csv 2>/dev/null
  rm -f /tmp/prof_${prec}.ncu-rep
done
timeout 900 ncu --set full --clock-control none -k regex:"standardise|ozaki_slice|gram_umma" -c 8 -o /tmp/prof_tom python scripts/ozaki_probe.py > gpurun_out/ncu_tom.log 2>&1
ncu -i /tmp/prof_tom.ncu-rep --page details > gpurun_out/ncu_pearson_ozaki_details.txt 2>/dev/null
ncu -i /tmp/prof_tom.ncu-rep --page raw --csv > gpurun_out/ncu_pearson_ozaki_raw.csv 2>/dev/null
rm -f /tmp/prof_tom.ncu-rep
timeout 1700 python scripts/config_runs.py > gpurun_out/config_runs.log 2>&1
echo "configs exit $?" >> gpurun_out/config_runs.log
timeout 1200 python scripts/scale_runs.py > gpurun_out/scale_n1.log 2>&1
echo "scale n1 exit $?" >> gpurun_out/scale_n1.log
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -2 gpurun_out/bench_n1.log | cut -c1-300; grep -E "exit" gpurun_out/scale_n1.log gpurun_out/config_runs.log; du -sh gpurun_out
