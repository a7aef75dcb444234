#!/bin/bash
# run 37: per-API host->host timings with CPU port legs; final N = 1 bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(only_if_missing=True)" > gpurun_out/build.log 2>&1
timeout 900 python scripts/api_timings.py > gpurun_out/api_timings.log 2>&1
echo "api exit $?" >> gpurun_out/api_timings.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?" >> gpurun_out/bench_n1.err
cut -c1-330 gpurun_out/api_timings.log; cut -c1-600 gpurun_out/bench_n1.json; tail -2 gpurun_out/bench_n1.err
