set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,launch__grid_size,launch__registers_per_thread,launch__occupancy_limit_shared_mem,smsp__warp_issue_stalled_barrier_per_warp_active.pct,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct,smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct,smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct --clock-control none -k regex:d2_ -s 45 -c 45 --csv --log-file gpurun_out/dw_metrics.csv python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-profile > gpurun_out/ncu_dw.out 2>&1
tail -n 2 gpurun_out/ncu_dw.out | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:d2_bwd -s 30 -c 6 -o gpurun_out/prof_dwb python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-profile > gpurun_out/ncu_dwb.out 2>&1
tail -n 2 gpurun_out/ncu_dwb.out | cut -c1-300
du -sh gpurun_out
