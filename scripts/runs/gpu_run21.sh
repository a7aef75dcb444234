#!/bin/bash
# run 21 (8 GPUs): scaling bench at N = 8 and 4
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.log 2>&1
echo "bench n$n exit $?" >> gpurun_out/bench_n$n.log
tail -2 gpurun_out/bench_n$n.log | cut -c1-300
done
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
