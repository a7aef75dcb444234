#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -k "not int8 and not tf32 and not fast" --timeout 900 > gpurun_out/pytest_gpu_f64.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_f64.log
timeout 300 python scripts/umma_debug.py > gpurun_out/umma_debug.log 2>&1
echo "debug exit $?" >> gpurun_out/umma_debug.log
timeout 900 python -m pytest tests -m gpu -q -k "int8 or tf32 or fast" --timeout 600 > gpurun_out/pytest_gpu_umma.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_umma.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_f64.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_f64.log
timeout 600 python bench.py --steps 5 --warmup 3 --precision i8 --no-cpu --no-tom > gpurun_out/bench_i8.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_i8.log
timeout 600 python bench.py --steps 5 --warmup 3 --precision tf32 --no-cpu > gpurun_out/bench_tf32.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_tf32.log
tail -3 gpurun_out/pytest_gpu_f64.log; cat gpurun_out/umma_debug.log | tail -40; tail -15 gpurun_out/pytest_gpu_umma.log
