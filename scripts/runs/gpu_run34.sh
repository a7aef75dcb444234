#!/bin/bash
# run 34: timing-only probe of a 128-bit phase B (mode 9) in the profiling build
mkdir -p gpurun_out
BIXCOR_LIB=$PWD/bixverse_b200/libbixcor_prof.so timeout 300 python scripts/perf_probe.py i8:0 i8:9 i8:4 i8:5 tf32:0 tf32:9 > gpurun_out/wide_probe.log 2>&1
cut -c1-200 gpurun_out/wide_probe.log
