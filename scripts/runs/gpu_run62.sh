mkdir -p gpurun_out/r2v
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2v/bench_n2.json 2> gpurun_out/r2v/bench_n2.err; echo bench rc=$?; tail -2 gpurun_out/r2v/bench_n2.err
