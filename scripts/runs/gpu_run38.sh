#!/bin/bash
# run 38: profiling build, modes 0 / 4 / 5 / 7 / 10 with per-phase cycle counters, twice each (noise check)
mkdir -p gpurun_out
export BIXCOR_LIB=$PWD/bixverse_b200/libbixcor_prof.so BIXCOR_UMMA_PROF=1
timeout 400 python scripts/perf_probe.py i8:0 i8:7 i8:4 i8:10 i8:5 i8:0 i8:7 i8:0:1.0 i8:7:1.0 > gpurun_out/fp64_probe.log 2>&1
cut -c1-200 gpurun_out/fp64_probe.log
