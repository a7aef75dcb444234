#!/bin/bash
# run 23 (8 GPUs): full-size scale runs at N = 8 and 4, bench at N = 8
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n scripts/scale_runs.py > gpurun_out/scale_n$n.log 2>&1
echo "scale n$n exit $?" >> gpurun_out/scale_n$n.log
grep -E "^\{|exit|Error|error" gpurun_out/scale_n$n.log | cut -c1-200
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29558 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.log 2>&1
echo "bench n8 exit $?" >> gpurun_out/bench_n8.log
tail -2 gpurun_out/bench_n8.log | cut -c1-250
