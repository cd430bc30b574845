"""Key metrics of every kernel launch in an ncu report (read on the CPU box): python tools/ncu_table.py rep.ncu-rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
want = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn smem"),
        ("launch__occupancy_limit_registers", "occ limit regs (blocks)"), ("launch__occupancy_limit_shared_mem", "occ limit smem (blocks)"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
        ("smsp__inst_executed.sum", "warp instructions"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
        ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps/cycle"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("dram__bytes_read.sum.per_second", "dram read rate"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput of peak"),
        ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
        ("l1tex__t_bytes.sum", "L1 bytes"), ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput of peak"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected")]
ki = h.index("Kernel Name")
for r in rows[2:]:
    if len(r) <= ki:
        continue
    print("== " + r[ki][:110])
    for key, label in want:
        if key in h:
            i = h.index(key)
            print("   %-28s %14s %s" % (label, r[i][:14], units[i]))
