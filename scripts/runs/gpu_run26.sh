#!/bin/bash
# run 26 (8 GPUs): scale_runs + bench at N = 8, 4, 2; reference arm at 8; single-process multi-device mode
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n scripts/scale_runs.py > gpurun_out/scale_n$n.log 2>&1
echo "scale n$n exit $?" >> gpurun_out/scale_n$n.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.log 2>&1
echo "bench n$n exit $?" >> gpurun_out/bench_n$n.log
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29568 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/bench_ref_n8.log 2>&1
timeout 600 python scripts/inproc_multi.py > gpurun_out/inproc_multi.log 2>&1
echo "inproc exit $?" >> gpurun_out/inproc_multi.log
grep -E "exit" gpurun_out/scale_n*.log gpurun_out/bench_n[248].log gpurun_out/inproc_multi.log; grep -E "^\{" gpurun_out/inproc_multi.log | cut -c1-200
