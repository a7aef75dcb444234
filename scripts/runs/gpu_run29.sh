#!/bin/bash
# run 29 (2 GPUs): in-process multi-device TOM test, NCCL drivers test, in-process timings on 1-2 devices
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "multi_device or nccl" > gpurun_out/pytest_multi.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_multi.log
tail -12 gpurun_out/pytest_multi.log
timeout 600 python scripts/inproc_multi.py > gpurun_out/inproc_multi.log 2>&1
echo "inproc exit $?" >> gpurun_out/inproc_multi.log
grep -E "TOM|exit|rror" gpurun_out/inproc_multi.log | cut -c1-200
