#!/bin/bash
# run 32: bulk-async store variants of the lean epilogue against the default build (timing + bit-exact checksum)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
for v in "" _bulk16 _bulk32 _bulk16x2; do
  echo "== lib$v" >> gpurun_out/bulk_probe.log
  BIXCOR_LIB=$PWD/bixverse_b200/libbixcor$v.so timeout 300 python scripts/perf_probe.py i8:0 tf32:0 i8:0:1.0 i8:0:6.0:500 >> gpurun_out/bulk_probe.log 2>&1
done
cat gpurun_out/bulk_probe.log | cut -c1-250
