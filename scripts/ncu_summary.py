"""Print the headline metrics of an ncu report (first kernel matching the regex)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, u = rows[0], rows[1]
keep = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "lts__t_sectors.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_op_global_atom.sum",
        "smsp__inst_executed_op_global_red.sum", "smsp__inst_executed_op_global_ld.sum", "smsp__inst_executed_op_global_st.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "lts__d_sectors_fill_sysmem.sum",
        "lts__average_t_sector_hit_rate_realtime.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed")
for v in rows[2:3]:
    for i, x in enumerate(h):
        if x in keep:
            print(f"{x},{u[i]},{v[i]}")
    st = [(float(v[i].replace(",", "")), x) for i, x in enumerate(h)
          if x.startswith("smsp__average_warps_issue_stalled") and x.endswith("per_issue_active.ratio")]
    for val, x in sorted(st, reverse=True)[:6]:
        print(f"{x},ratio,{val:.2f}")
