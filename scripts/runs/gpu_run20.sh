#!/bin/bash
# run 20 (2 GPUs): in-process multi-device test, memcheck on small shapes, N=2 bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "multi_device" > gpurun_out/pytest_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_multi.log
tail -3 gpurun_out/pytest_multi.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/umma_debug.py > gpurun_out/memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/memcheck.log
tail -6 gpurun_out/memcheck.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2>&1
echo "bench n2 exit $?" >> gpurun_out/bench_n2.log
tail -2 gpurun_out/bench_n2.log | cut -c1-400
