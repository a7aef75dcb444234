#!/bin/bash
# ncu captures of round 2 (one gpurun call; every target runs plain first, `&&` gates the profiled run).
# Usage: bash scripts/ncu_r02.sh <set>   with set = c2 | tom | list
set -u
O=gpurun_out/ncu_r02
mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
cap() {  # name, kernel regex, skip, count, script args...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  python scripts/ncu_target.py "$@" > $O/${name}_plain.log 2>&1 && \
  $NCU -k regex:$rx -s $skip -c $cnt -o $O/$name python scripts/ncu_target.py "$@" > $O/${name}_ncu.log 2>&1
  echo "$name rc=$?"
  # gpurun brings back at most 64 MiB: keep the text pages, drop the report itself
  if [ -f $O/$name.ncu-rep ]; then
    ncu -i $O/$name.ncu-rep --page details > $O/${name}_details.txt 2>/dev/null
    ncu -i $O/$name.ncu-rep --page raw --csv > $O/${name}_raw.csv 2>/dev/null
    rm -f $O/$name.ncu-rep
  fi
}
case "${1:-c2}" in
  c2)
    # per invocation of config 2: one column kernel + one Gram kernel (Ozaki: 6 passes + slicer + finisher)
    cap c2_i8_gram gram_umma 2 1 c2 auto
    cap c2_i8_rank rank_columns_warp 2 1 c2 auto
    cap c2_f64_gram gram_dmma 2 1 c2 f64
    cap c2_tf32_gram gram_umma 2 1 c2 tf32
    cap c2_f16_gram gram_umma 2 1 c2 f16
    ;;
  tom)
    # per TOM invocation the tcgen05 kinds launch two gram_umma kernels (dense affinity, then A*A): the A*A is the 2nd
    cap tom_f16 gram_umma 5 1 tom f16
    cap tom_tf32 gram_umma 5 1 tom tf32
    ;;
  s32)
    # host -> host call: 20 band launches per invocation, bottom-up; launch 59 (3 invocations) is the top band (8 block rows x 157)
    cap c2_s32_gram_topband gram_umma 59 1 c2host auto
    ;;
  list)
    python bench.py --steps 2 --warmup 3 --no-alt --no-tom --no-cpu --no-extras > $O/bench_plain.log 2>&1 && \
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default.csv \
        python bench.py --steps 2 --warmup 3 --no-alt --no-tom --no-cpu --no-extras > $O/bench_ncu.log 2>&1
    echo "list rc=$?"
    ;;
esac
ls -la $O | tail -20
