#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "rank or cor_upper or int8 or multi_device" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-tom --no-cpu > gpurun_out/bench_n1.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2>&1
echo "bench n2 exit $?" >> gpurun_out/bench_n2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.log 2>&1
echo "ref exit $?" >> gpurun_out/bench_ref_n2.log
tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench_n1.log | cut -c1-600;  tail -3 gpurun_out/bench_n2.log | cut -c1-3000; tail -2 gpurun_out/bench_ref_n2.log | cut -c1-400
