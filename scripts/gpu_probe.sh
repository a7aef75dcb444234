#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python scripts/perf_probe.py > gpurun_out/perf_probe.log 2>&1
cat gpurun_out/perf_probe.log
