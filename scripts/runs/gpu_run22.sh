#!/bin/bash
# run 22 (2 GPUs): NCCL world-2 test of all drivers, full-size scale runs at N = 2 and 1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_dist_nccl.py -m gpu -q --timeout 600 > gpurun_out/pytest_nccl.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_nccl.log
tail -15 gpurun_out/pytest_nccl.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/scale_runs.py > gpurun_out/scale_n2.log 2>&1
echo "scale n2 exit $?" >> gpurun_out/scale_n2.log
grep -E "^\{|exit|Error|error" gpurun_out/scale_n2.log | cut -c1-260
timeout 900 python scripts/scale_runs.py > gpurun_out/scale_n1.log 2>&1
echo "scale n1 exit $?" >> gpurun_out/scale_n1.log
grep -E "^\{|exit|Error|error" gpurun_out/scale_n1.log | cut -c1-260
