"""Summarise one kernel of an .ncu-rep (ncu --set full) as text for profiles/: duration, DRAM bytes, cache and pipe
utilisation, occupancy, registers.      python tools/ncu_kernel_summary.py report.ncu-rep [label]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
label = sys.argv[2] if len(sys.argv) > 2 else rep
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"]
for r in rows[2:]:
    name = r[h.index("Kernel Name")] if "Kernel Name" in h else "?"
    print("# %s: %s" % (label, name))
    for k in WANT:
        if k in h:
            print("%-72s %-10s %s" % (k, u[h.index(k)], r[h.index(k)]))
