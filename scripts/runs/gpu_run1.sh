#!/bin/bash
# First GPU pass: parity tests (FP64 path), bench, DMMA variant sweep, ncu launch list.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -k "not int8 and not tf32 and not fast" --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_f64.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_f64.log
for v in 0 2; do
  BIXCOR_DMMA_VARIANT=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-tom --no-cpu > gpurun_out/bench_f64_variant$v.log 2>&1
done
timeout 600 python bench.py --steps 2 --warmup 3 --no-tom --no-cpu > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv \
  python bench.py --steps 2 --warmup 3 --no-tom --no-cpu > gpurun_out/ncu_launches.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench_f64.log | cut -c1-1500
