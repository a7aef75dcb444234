#!/bin/bash
# run 31 (2 GPUs): fused cor_module_tom test + probe, refreshed N = 1 / N = 2 scale rows
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
timeout 300 python -m pytest tests -q -m gpu -x -k "cor_module_tom or layout or tom" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/tom_packed_probe.py > gpurun_out/tom_packed.log 2>&1
echo "probe exit $?" >> gpurun_out/tom_packed.log
SCALE_CPU=0 timeout 400 python scripts/scale_runs.py > gpurun_out/scale_n1.log 2>&1
echo "scale n1 exit $?" >> gpurun_out/scale_n1.log
SCALE_CPU=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 scripts/scale_runs.py > gpurun_out/scale_n2.log 2>&1
echo "scale n2 exit $?" >> gpurun_out/scale_n2.log
tail -3 gpurun_out/pytest_gpu.log; cut -c1-330 gpurun_out/tom_packed.log; grep exit gpurun_out/scale_n[12].log
