mkdir -p gpurun_out/r2z
timeout 200 python -m pytest tests/test_gpu_dist_nccl.py tests/test_gpu_parity.py -m gpu -x -q -k "world2 or exchange" > gpurun_out/r2z/pytest_xchg_final.log 2>&1; tail -2 gpurun_out/r2z/pytest_xchg_final.log
