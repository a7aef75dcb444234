mkdir -p gpurun_out/r2x
python -m pytest tests -m gpu -x -q -k "rank or spearman or cor_upper or exchange" > gpurun_out/r2x/pytest_rank.log 2>&1; tail -3 gpurun_out/r2x/pytest_rank.log
python scripts/rank_probe.py bixverse_b200/libbixcor.so bixverse_b200/libbixcor_rank3.so bixverse_b200/libbixcor_rankold.so > gpurun_out/r2x/rank_probe.log 2>&1; cat gpurun_out/r2x/rank_probe.log
