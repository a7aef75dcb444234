mkdir -p gpurun_out/r2t
timeout 300 python -m pytest tests/test_gpu_dist_nccl.py tests/test_gpu_parity.py -m gpu -x -q -k "world2 or exchange" > gpurun_out/r2t/pytest.log 2>&1; tail -15 gpurun_out/r2t/pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-alt --no-tom --no-extras > gpurun_out/r2t/bench_n2.json 2> gpurun_out/r2t/bench_n2.err; echo bench rc=$?; tail -3 gpurun_out/r2t/bench_n2.err
