"""Print a few named metrics per kernel from an `ncu --page raw --csv` dump:  python scripts/ncu_pick.py raw.csv [substr ...]"""
import csv
import sys

DEFAULT = ["gpu__time_duration.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "smsp__warps_active.avg.per_cycle_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum",
           "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
want = sys.argv[2:] or DEFAULT
for r in rows[2:]:
    print(r[h.index("Kernel Name")][:70])
    for w in want:
        for j, c in enumerate(h):
            if c == w or c.endswith("." + w):
                print("   %-90s %s" % (w, r[j]))
                break
