#!/bin/bash
# run 35: wide (N = 256) int8 MMA variant: full GPU suite on the variant library + timing against the default
mkdir -p gpurun_out
export WIDE=$PWD/bixverse_b200/libbixcor_wide.so
BIXCOR_LIB=$WIDE timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_wide.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_wide.log
for v in "" _wide; do
  echo "== lib$v" >> gpurun_out/wide_mma.log
  BIXCOR_LIB=$PWD/bixverse_b200/libbixcor$v.so timeout 300 python scripts/perf_probe.py i8:0 i8:0:1.0 i8:0:6.0:500 i8:0:6.0:1800 >> gpurun_out/wide_mma.log 2>&1
  BIXCOR_LIB=$PWD/bixverse_b200/libbixcor$v.so BIXCOR_UMMA_CTAS=1 timeout 300 python scripts/perf_probe.py i8:0 >> gpurun_out/wide_mma.log 2>&1
  BIXCOR_LIB=$PWD/bixverse_b200/libbixcor$v.so timeout 300 python scripts/ozaki_probe.py >> gpurun_out/wide_mma.log 2>&1
done
tail -4 gpurun_out/pytest_wide.log; cut -c1-220 gpurun_out/wide_mma.log
