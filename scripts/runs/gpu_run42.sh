#!/bin/bash
# run 42 (2 GPUs): NCCL world-2 tests incl. the f16 kind, scale rows at N = 1 / 2, bench N = 1, fused cor_module_tom probe
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -k "nccl or multi_device or dist" > gpurun_out/pytest_dist.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_dist.log
SCALE_CPU=0 timeout 400 python scripts/scale_runs.py > gpurun_out/scale_n1.log 2>&1
echo "scale n1 exit $?" >> gpurun_out/scale_n1.log
SCALE_CPU=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 scripts/scale_runs.py > gpurun_out/scale_n2.log 2>&1
echo "scale n2 exit $?" >> gpurun_out/scale_n2.log
timeout 300 python scripts/tom_packed_probe.py > gpurun_out/tom_packed.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?" >> gpurun_out/bench_n1.err
tail -3 gpurun_out/pytest_dist.log; grep exit gpurun_out/scale_n[12].log gpurun_out/bench_n1.err; grep "f16\|3 x f16" gpurun_out/scale_n1.log gpurun_out/scale_n2.log | cut -c1-230; cut -c1-200 gpurun_out/tom_packed.log
