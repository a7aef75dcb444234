#!/bin/bash
# usage: bash scripts/e2e_dist_sweep.sh N  -> gpurun_out/e2e_dist_nN.log
N=${1:-8}
mkdir -p gpurun_out
run() { env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/e2e_dist.py 2>/dev/null | grep '^{' | tee -a gpurun_out/e2e_dist_n$N.log; }
run BIXCOR_X=default
run BIXCOR_STAGE_SPIN=0
run BIXCOR_COPY_THREADS=2
run BIXCOR_COPY_THREADS=2 BIXCOR_STAGE_SPIN=0
run BIXCOR_COPY_THREADS=4 BIXCOR_STAGE_SPIN=0
run BIXCOR_STAGE_SPIN=0 BIXCOR_STAGE_SLOT_MB=16 BIXCOR_STAGE_SLOTS=8 BIXCOR_STAGE_PIECE_KB=4096
