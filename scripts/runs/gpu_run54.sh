mkdir -p gpurun_out/r2u
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2u/bench_n8.json 2> gpurun_out/r2u/bench_n8.err; echo bench rc=$?; tail -3 gpurun_out/r2u/bench_n8.err
