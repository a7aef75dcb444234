#!/bin/bash
# run 16 (2 GPUs): scaling bench at N=2 with the host-sharded e2e driver, plus the reference arm under torchrun
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2>&1
echo "bench n2 exit $?" >> gpurun_out/bench_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.log 2>&1
echo "ref exit $?" >> gpurun_out/bench_ref_n2.log
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
tail -2 gpurun_out/bench_n2.log | cut -c1-1600; tail -2 gpurun_out/bench_ref_n2.log | cut -c1-900
