#!/usr/bin/env python
"""Compact per-kernel summary of an `ncu --set full` report (raw page -> a few dozen counters).
usage: python scripts/ncu_summary.py report.ncu-rep [out.csv]"""
import csv, subprocess, sys, io

WANT = [
    ("time_ms", "gpu__time_duration.sum"),
    ("regs", "launch__registers_per_thread"),
    ("smem_kb", "launch__shared_mem_per_block_allocated"),
    ("occ_lim_regs", "launch__occupancy_limit_registers"),
    ("occ_lim_smem", "launch__occupancy_limit_shared_mem"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue_active_pct", "smsp__issue_active.avg.pct"),
    ("inst_per_cycle", "sm__inst_executed.avg.per_cycle_active"),
    ("pipe_fp64_pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("pipe_fp64_cycles_pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    ("pipe_xu_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("pipe_alu_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("pipe_fma_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("pipe_lsu_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("pipe_cbu_pct", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active"),
    ("pipe_uniform_pct", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active"),
    ("inst_executed", "smsp__inst_executed.sum"),
    ("thread_inst_per_inst", "smsp__thread_inst_executed_per_inst_executed.ratio"),
    ("dfma_x2", "derived__smsp__sass_thread_inst_executed_op_dfma_pred_on_x2"),
    ("dadd", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"),
    ("dmul", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"),
    ("dfma", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"),
    ("dram_read_mb", "dram__bytes_read.sum"),
    ("dram_write_mb", "dram__bytes_write.sum"),
    ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ("local_ld_req", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum"),
    ("local_st_req", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum"),
    ("smem_bank_conf_ld", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"),
    ("stall_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall_short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall_math_throttle", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("stall_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall_branch", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"),
    ("stall_no_inst", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
    ("stall_dispatch", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"),
    ("stall_not_selected", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
    ("stall_imc", "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio"),
    ("stall_lg_throttle", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
    ("stall_mio_throttle", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
]

def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = [["kernel", "grid", "block"] + [w[0] + (f" [{units[idx[w[1]]]}]" if w[1] in idx and units[idx[w[1]]] else "") for w in WANT]]
    for r in rows[2:]:
        line = [r[idx["Kernel Name"]].split("(")[0], r[idx["Grid Size"]], r[idx["Block Size"]]]
        for _, m in WANT:
            line.append(r[idx[m]] if m in idx else "")
        out.append(line)
    # transposed print: one column per kernel
    w = csv.writer(open(sys.argv[2], "w", newline="") if len(sys.argv) > 2 else sys.stdout)
    for j in range(len(out[0])):
        w.writerow([row[j] for row in out])

main()
