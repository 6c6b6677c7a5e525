import csv, subprocess, sys
rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
want = ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__cycles_elapsed.avg','smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed','smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed','smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
lines = []
for w in want:
    idx = [i for i, c in enumerate(h) if c == w]
    if not idx:
        continue
    i = idx[0]
    lines.append(f"{w} [{rows[1][i]}]: " + " | ".join(r[i][:52] for r in rows[2:]))
open(out, "w").write(title + "\n\n" + "\n".join(lines) + "\n")
print("\n".join(lines))
