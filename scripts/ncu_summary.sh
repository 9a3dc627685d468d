#!/bin/bash
# usage: ncu_summary.sh <report.ncu-rep> <out.txt> : the metrics quoted in DESIGN.md, one block per captured launch
ncu -i "$1" --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; units=rows[1]
want=['Kernel Name','launch__grid_size','launch__block_size','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','gpu__time_duration.sum','sm__cycles_elapsed.max','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__inst_executed_pipe_tensor.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']
idx={h:i for i,h in enumerate(hdr)}
for r in rows[2:]:
    print('-----')
    for w in want:
        if w in idx: print(w,'=',r[idx[w]],units[idx[w]])
" > "$2"
