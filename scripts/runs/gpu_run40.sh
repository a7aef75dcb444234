#!/bin/bash
# run 40: f16 split kind: promotion chunk 2 vs 4 K blocks, single CTA vs CTA pairs on the TOM
mkdir -p gpurun_out
timeout 300 python scripts/f16_probe.py f16 > gpurun_out/f16_tom.log 2>&1
BIXCOR_UMMA_CTAS=3 timeout 300 python scripts/f16_probe.py f16 >> gpurun_out/f16_tom.log 2>&1
BIXCOR_LIB=$PWD/bixverse_b200/libbixcor_f16c4.so timeout 300 python scripts/f16_probe.py f16 >> gpurun_out/f16_tom.log 2>&1
BIXCOR_LIB=$PWD/bixverse_b200/libbixcor_f16c4.so timeout 300 python scripts/perf_probe.py f16:0 f16:0:1.0 >> gpurun_out/f16_tom.log 2>&1
cut -c1-260 gpurun_out/f16_tom.log
