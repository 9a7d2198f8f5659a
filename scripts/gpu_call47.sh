#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"shade_bwd_kernel|shade_fwd_kernel|project_bwd_kernel|project_fwd_kernel|build_sublists" -c 5 -f -o gpurun_out/c47_prof python scripts/bench_composite.py --iters 1 > /dev/null 2>&1
ls -la gpurun_out/c47_prof.ncu-rep
ncu -i gpurun_out/c47_prof.ncu-rep --page raw --csv > gpurun_out/c47_prof.raw.csv 2>/dev/null
python scripts/ncu_pick.py gpurun_out/c47_prof.raw.csv gpu__time_duration.sum smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active launch__registers_per_thread smsp__warps_active.avg.per_cycle_active dram__bytes_read.sum dram__bytes_write.sum smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_wait_per_issue_active.ratio smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed lts__t_sector_hit_rate.pct
