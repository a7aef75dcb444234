mkdir -p gpurun_out/r2v
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/r2v/pytest_2gpu.log 2>&1; tail -4 gpurun_out/r2v/pytest_2gpu.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2v/bench_n2.json 2> gpurun_out/r2v/bench_n2.err; echo bench rc=$?; tail -2 gpurun_out/r2v/bench_n2.err
