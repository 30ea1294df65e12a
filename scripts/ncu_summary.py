"""Summarise an .ncu-rep (raw page) into the handful of metrics the roofline discussion needs."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_hmma.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "lts__t_sector_hit_rate.pct", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "lts__t_bytes.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
idx = [(i, h) for i, h in enumerate(hdr) if h in want]
for r in rows[2:]:
    print("; ".join(f"{h.split('.')[0] if h!='Kernel Name' else 'kernel'}={r[i][:70]}{'' if h=='Kernel Name' else ' '+units[i]}" for i, h in idx))
