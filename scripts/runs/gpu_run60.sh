mkdir -p gpurun_out/r2x
for v in head rankold rank3 ""; do
  lib=bixverse_b200/libbixcor${v:+_$v}.so
  BIXCOR_LIB=$PWD/$lib ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__warps_active.avg.per_cycle_active,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_lsu.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,launch__occupancy_limit_registers,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct,smsp__warp_issue_stalled_wait_per_warp_active.pct,smsp__warp_issue_stalled_not_selected_per_warp_active.pct,smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct,smsp__warp_issue_stalled_no_instruction_per_warp_active.pct --clock-control none -k regex:rank_columns_warp -s 2 -c 1 --csv --log-file gpurun_out/r2x/ncu_rank_${v:-packed}.csv python scripts/ncu_target.py c2 auto > gpurun_out/r2x/ncu_rank_${v:-packed}.log 2>&1
  echo "$v rc=$?"
done
