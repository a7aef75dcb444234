#!/bin/bash
# run 30 (8 GPUs): final scale runs at N = 8 and 4, single-process multi-device mode up to 8 devices
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
for n in 8 4; do
SCALE_CPU=0 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n scripts/scale_runs.py > gpurun_out/scale_n$n.log 2>&1
echo "scale n$n exit $?" >> gpurun_out/scale_n$n.log
done
timeout 500 python scripts/inproc_multi.py > gpurun_out/inproc_multi.log 2>&1
echo "inproc exit $?" >> gpurun_out/inproc_multi.log
grep -E "exit" gpurun_out/scale_n[48].log gpurun_out/inproc_multi.log; grep -E "TOM" gpurun_out/inproc_multi.log | cut -c1-160
