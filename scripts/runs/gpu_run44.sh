#!/bin/bash
# run 44: final round-1 evidence with the f16 kind: full suite, smoke, bench line, ncu captures of the f16 kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?" >> gpurun_out/bench_n1.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gram_umma" -s 1 -c 1 -o /tmp/prof_f16_tom \
  python scripts/ncu_f16_tom.py > gpurun_out/ncu_f16_tom.log 2>&1
ncu -i /tmp/prof_f16_tom.ncu-rep --page details > gpurun_out/ncu_f16_tom_details.txt 2>/dev/null
ncu -i /tmp/prof_f16_tom.ncu-rep --page raw --csv > gpurun_out/ncu_f16_tom_raw.csv 2>/dev/null
rm -f /tmp/prof_f16_tom.ncu-rep
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -1 gpurun_out/bench_n1.err; cut -c1-300 gpurun_out/bench_n1.json
grep -E "Duration|Compute \(SM\) Throughput|Memory Throughput|Registers Per|Tensor" gpurun_out/ncu_f16_tom_details.txt | head -12; du -sh gpurun_out
