mkdir -p gpurun_out/r2z
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 --no-tom --no-extras --no-alt > gpurun_out/r2z/bench_n2_short.json 2> gpurun_out/r2z/bench_n2_short.err; echo bench rc=$?; tail -2 gpurun_out/r2z/bench_n2_short.err
