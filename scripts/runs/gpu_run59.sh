mkdir -p gpurun_out/r2x
python scripts/rank_probe.py bixverse_b200/libbixcor_head.so bixverse_b200/libbixcor_rankold.so bixverse_b200/libbixcor.so > gpurun_out/r2x/rank_probe2.log 2>&1; cat gpurun_out/r2x/rank_probe2.log
